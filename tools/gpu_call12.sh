#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_op_read_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_warps_issue_stalled_lg_throttle,smsp__pcsamp_sample_buffer_full,smsp__warps_active.avg.per_cycle_active,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__m_l1tex2xbar_write_sectors.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum
for spec in "" "B200_ACC_RSUB=8" "B200_ACC_RSUB=16,B200_ACC_CHUNK=5"; do
  tag=$(echo "$spec" | tr ',=' '__'); [ -z "$tag" ] && tag=default
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_accum_trie -s 4 -c 1 --csv --log-file gpurun_out/ncu_acc_$tag.csv python tools/qt_sweep.py "$spec" > gpurun_out/ncu_acc_$tag.log 2>&1
done
ls -la gpurun_out/ncu_acc_*
