// common.cuh -- shared device-side structures of the forward-simulation engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Device view of one layout atom after host-side expansion of the prefix table into independent
// circuits (see engine.cu: expand_table).  Circuits are stored longest-first.
struct AtomDev {
    int dim;            // d = 4^n_qubits
    int n_ops, n_rho, n_eff;
    int n_circ;         // = number of table rows
    int max_depth;
    int64_t n_elements;
    const uint32_t* circ_ptr;   // [n_circ+1] into circ_ops
    const int32_t*  circ_ops;   // expanded op sequences
    const int32_t*  circ_prep;  // [n_circ]
    const int32_t*  out_ptr;    // [n_circ+1] into out_eff/out_el
    const int32_t*  out_eff;
    const int32_t*  out_el;
};

// Device view of the model tensors ("W space" vector M = [G | rho | E]) plus transposed gates.
struct ModelDev {
    const double* M;     // [n_w]   G[n_ops][d][d] row-major, then rho[n_rho][d], then E[n_eff][d]
    const double* Gt;    // [n_ops][d][d]  Gt[g][j][i] = G[g][i][j]
    int64_t n_w;
    int64_t off_rho;     // n_ops*d*d
    int64_t off_eff;     // off_rho + n_rho*d
};

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    return __shfl_xor_sync(0xffffffffu, v, mask);
}
