#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/oz_test.py c2_full_layout 3 > gpurun_out/r2o_oz.log 2>&1
tail -30 gpurun_out/r2o_oz.log
