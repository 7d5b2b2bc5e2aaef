// kernels_level.cuh -- level-batched dense propagation for large superoperators (d = 64, 256): the
// "dense contraction" configs of BASELINE.json (3-qubit d = 64, 4-qubit d = 256).
//
// A d x d mat-vec per circuit step is hopeless on a GPU at d = 256: every step re-reads a 512 KB gate matrix
// (measured: the per-circuit generic kernel moves 8 GB through L2 for a 500-circuit sample).  Instead all
// circuits advance together, one depth level per launch: at level k the circuits whose k-th layer is gate g
// form the columns of a dense product  S_{k+1}[:, cols_g] = G_g . S_k[:, cols_g]  -- a real GEMM, run on the FP64
// tensor cores (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind).  The reference does the same arithmetic one
// circuit and one row at a time (OpCRep_Dense::acton, pygsti/evotypes/densitymx/opcreps.cpp:40-54).
//
//   k_level_gemm<D> : one CTA per tile of <= 32 circuits that share the gate at this level.  The tile's states
//                     (32 x D doubles) are staged in shared memory (row stride D+4: conflict-free A fragments);
//                     each of the 4 warps owns D/4 output components for all 32 circuits:
//                     acc[4 m-tiles][D/32 n-tiles] in registers, B fragments (rows of G) streamed from L2 once
//                     per CTA.  States ping-pong between two buffers by level parity.
//   k_level_probs<D>: p = E . s_final (one warp per circuit), final state in buffer (L & 1).
#pragma once
#include "common.cuh"

struct LevelTile { uint32_t first; uint16_t count; uint16_t gate; };   // entries [first, first+count) of lvl_circ

// (Round 2 measured two variants of this loop on BASELINE config 5 and reverted both: a four-block register ring for the B
// fragments, 3.10 vs 2.93 ms, and split-K over two warp groups, 3.32 ms.  The level kernels are not bound by the L2 latency of
// one warp or by the 64-deep DMMA chain but by how little work one depth level holds: ~0.33 GFLOP = 9 us of the whole GPU.)
// Also measured and NOT kept (late round 2): ONE persistent launch for all levels -- (tile, output quarter) items from an atomic counter in
// level-major order, per-circuit arrival counters (release / acquire) instead of the barrier between levels: 3.39 vs 2.90 ms.  With random
// circuits the 32 circuits of a level-(k+1) tile come from ~14 different level-k gate groups, so an item waits for nearly the whole
// previous level anyway; the polling and the fences then cost more than the 128 launch gaps they remove.
// Inner product loop shared by k_level_gemm and k_level_gemm_rows (kernels_levelj.cuh): acc[mt][nt] += A[32 x D] . B[D x 8 NT].
// A fragments come from shared memory (row stride D + 4), B fragments (rows of G, 32 bytes per row and K step) from L2.  The B
// loads of the NEXT 16 K (the four sectors of one 128-byte line of each row) are issued before the 4 x 8 DMMA of the current
// 16 K: the first version loaded B inside the step that used it (unroll 2) and left the FP64 tensor pipe waiting on L2.
template <int D, int NT>
__device__ __forceinline__ void level_gemm_core(const double* __restrict__ ap, const double* __restrict__ bp, double (&acc)[4][NT][2])
{
    constexpr int LDS_ = D + 4;
    double bb[4][NT];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bb[s4][nt] = __ldg(bp + (size_t)nt * 8 * D + 4 * s4);
#pragma unroll 1
    for (int k0 = 0; k0 < D; k0 += 16) {
        double bn[4][NT];
        const int kn = (k0 + 16 < D) ? k0 + 16 : k0;          // (last block: reload the current one, unused)
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bn[s4][nt] = __ldg(bp + (size_t)nt * 8 * D + kn + 4 * s4);
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            double af[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = ap[mt * 8 * LDS_ + k0 + 4 * s4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bb[s4][nt]);
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bb[s4][nt] = bn[s4][nt];
    }
}

template <int D>
__global__ void __launch_bounds__(128)
k_level_init(AtomDev a, const double* __restrict__ rho, double* __restrict__ S0)
{
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)a.n_circ * D;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx / D), i = (int)(idx - (int64_t)c * D);
        S0[idx] = rho[(int64_t)a.circ_prep[c] * D + i];
    }
}

// dynamic smem: 32 * (D + 4) doubles
template <int D>
__global__ void __launch_bounds__(128)
k_level_gemm(const double* __restrict__ G, const LevelTile* __restrict__ tiles, const uint32_t* __restrict__ lvl_circ,
             const double* __restrict__ Sin, double* __restrict__ Sout)
{
    constexpr int LDS_ = D + 4;                 // row stride (doubles): (D+4) mod 16 = 4 -> conflict-free fragment loads
    constexpr int NT = 2;                       // n-tiles (of 8 outputs) per warp: a CTA owns 64 outputs (blockIdx.y)
    extern __shared__ __align__(16) double st[];   // [32][LDS_]
    const LevelTile tl = tiles[blockIdx.x];
    const uint32_t* circ = lvl_circ + tl.first;
    const int cnt = tl.count;
    const double* Gg = G + (size_t)tl.gate * D * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // stage the tile's states (zero rows for the padding circuits)
    for (int idx = threadIdx.x; idx < 32 * (D / 2); idx += blockDim.x) {
        const int r = idx / (D / 2), j2 = idx - r * (D / 2);
        double2 v = make_double2(0.0, 0.0);
        if (r < cnt) v = *reinterpret_cast<const double2*>(Sin + (size_t)circ[r] * D + 2 * j2);
        *reinterpret_cast<double2*>(st + r * LDS_ + 2 * j2) = v;
    }
    __syncthreads();
    const int mrow = lane >> 2, q = lane & 3;
    const int n0 = blockIdx.y * 64 + warp * 16; // this warp's output components [n0, n0 + 16)
    double acc[4][NT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    // S_next[c][i] = sum_j S[c][j] G[i][j] : A[m=c][k=j] from smem, B[k=j][n=i] = G[i][j] (row i contiguous in j)
    const double* bp = Gg + (size_t)(n0 + mrow) * D + q;      // + nt*8*D + k0
    const double* ap = st + mrow * LDS_ + q;                  // + mt*8*LDS_ + k0
    level_gemm_core<D, NT>(ap, bp, acc);
    // D fragment: row m = circuit, cols 2q, 2q+1 of the n-tile
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int r = mt * 8 + mrow;
        if (r < cnt) {
            double* op = Sout + (size_t)circ[r] * D + n0 + 2 * q;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(op + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128)
k_level_probs(AtomDev a, const double* __restrict__ E, const double* __restrict__ S0, const double* __restrict__ S1,
              double* __restrict__ out, int64_t el_stride)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 4 + warp, nw = gridDim.x * 4;
    for (int c = gw; c < a.n_circ; c += nw) {
        const int L = (int)(a.circ_ptr[c + 1] - a.circ_ptr[c]);
        const double* s = ((L & 1) ? S1 : S0) + (size_t)c * D;
        for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
            const double* e = E + (size_t)a.out_eff[qo] * D;
            double part = 0.0;
#pragma unroll
            for (int i = lane; i < D; i += 32) part += e[i] * s[i];
#pragma unroll
            for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
            if (lane == 0) out[(int64_t)a.out_el[qo] * el_stride] = part;
        }
    }
}
