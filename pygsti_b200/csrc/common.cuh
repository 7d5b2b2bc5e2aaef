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

// FP64 tensor-core MMA (there is no FP64 kind of tcgen05.mma: DMMA m8n8k4 is the FP64 tensor path on sm_100a)
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define B200_PEERS_MAX 7
// Where the d = 16 Jacobian kernel stores its accumulators
#define D16_SPAM_MAX 512  // SPAM/unmapped column list entries staged in shared memory
struct D16Args {
    const int32_t* colmap;    // [n_w]  J column of W index w (gate part used by the register epilogue), -1 = none
    const int32_t* spam_col;  // [n_spam] columns NOT fed by a gate element ...
    const int32_t* spam_w;    // [n_spam] ... and the rho/effect W index feeding each (or -1 -> zero)
    int n_spam;
    double* J;                // [n_elements][ld]
    int64_t ld;
    double* probs;            // [n_elements] or nullptr
    const double* row_scale;  // [n_elements] or nullptr: J row el is multiplied by row_scale[el] in the store epilogue
    // fused exchange (b200_fill_dprobs_bcast_dev): the same rows are also stored at the same offsets of the peers' arrays
    int n_peers;
    double* peerJ[B200_PEERS_MAX];
    double* peerP[B200_PEERS_MAX];
};
