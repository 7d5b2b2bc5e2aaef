// kernels_jtj.cuh -- C = A^T B over the ELEMENT axis on the FP64 tensor cores (DMMA), hand-written: the reductions that
// turn per-element derivative arrays into what the optimizer consumes, without the arrays leaving the device.
//
//   J^T J, J^T f   (A = B = J [n_elements x n_params], row-major)  -- DistributableCOPALayout.fill_jtj / fill_jtf
//                  (pygsti/layouts/distlayout.py:1220-1359; consumed by pygsti/optimize/simplerlm.py:677-678)
//   J1^T J2        (MLE Hessian block: sum_el w_d[el] dp[el,a] dp[el,b], objectivefns.py:4914-4990)
//
// Both operands are K-major in memory (K = element index, the row of J), which is exactly the "row.col" operand layout
// of mma.sync.m8n8k4.f64 when the tile is staged in shared memory as [k][column]: A fragment = As[k = lane&3][m = lane>>2],
// B fragment = Bs[k][n]; a padded row stride of 132 doubles (= 4 mod 16) makes both fragment loads bank-conflict free.
//
// Work decomposition: output tiles of 128 x 128 (lower triangle only for A == B) x K slices; one CTA (8 warps, warp tile
// 32 x 64 = 64 accumulator doubles per lane) per (slice, tile), 4-stage cp.async pipeline over 16 rows of J per stage.
// Slices write partial tiles to scratch and k_atb_reduce sums them in slice order: deterministic (no atomics).
// There is no FP64 kind of tcgen05.mma; the DMMA pipe (measured 37.2 TFLOP/s, tools/ubench_fp64.cu) is the roofline.
#pragma once
#include "common.cuh"

#define JT_T 128
#define JT_KC 16
#define JT_ST 4
#define JT_LDS 132
#define JT_STAGE_DOUBLES (2 * JT_KC * JT_LDS + JT_KC)      // A panel, B panel, f chunk

struct AtbArgs {
    const double* A; int64_t lda; int na;      // A [nE x na] (row stride lda, even; base 16-byte aligned)
    const double* B; int64_t ldb; int nb;      // B [nE x nb]; B == A && tri: symmetric product, lower-triangle tiles only
    int64_t nE;
    int tri;                                   // 1: tiles (bi >= bj) of the symmetric product A^T A
    int n_bi, n_bj, n_tiles, n_slices;
    int64_t rows_per_slice;                    // multiple of JT_KC
    const double* f;                           // [nE] or nullptr: also A^T f (accumulated by the bj == 0 / diagonal tiles)
    double* part;                              // [n_slices][n_tiles][128*128]
    double* part_f;                            // [n_slices][n_bi*128]
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void atb_tile_of(const AtbArgs& p, int tile, int& bi, int& bj) {
    if (p.tri) {
        bi = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= tile) ++bi;
        while (bi * (bi + 1) / 2 > tile) --bi;
        bj = tile - bi * (bi + 1) / 2;
    } else { bi = tile / p.n_bj; bj = tile - bi * p.n_bj; }
}

__global__ void __launch_bounds__(256, 1)
k_atb_dmma(AtbArgs p)
{
    extern __shared__ __align__(16) double sm_atb[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = blockIdx.x / p.n_tiles, tile = blockIdx.x - slice * p.n_tiles;
    int bi, bj;
    atb_tile_of(p, tile, bi, bj);
    const bool same = p.tri && bi == bj;                 // diagonal tile of the symmetric product: one panel serves both operands
    const bool do_f = p.f != nullptr && (p.tri ? bi == bj : bj == 0);
    const int64_t k_lo = (int64_t)slice * p.rows_per_slice;
    const int64_t k_hi = min(k_lo + p.rows_per_slice, p.nE);
    const int n_it = k_hi > k_lo ? (int)((k_hi - k_lo + JT_KC - 1) / JT_KC) : 0;
    const int colA = bi * JT_T, colB = bj * JT_T;

    auto load_stage = [&](int st, int it) {
        double* As = sm_atb + (size_t)st * JT_STAGE_DOUBLES;
        double* Bs = As + JT_KC * JT_LDS;
        double* fs = Bs + JT_KC * JT_LDS;
        const int64_t k0 = k_lo + (int64_t)it * JT_KC;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int chunk = tid + c * 256, row = chunk >> 6, c2 = (chunk & 63) * 2;
            const int64_t k = k0 + row;
            const bool kok = k < k_hi;
            {
                const int col = colA + c2;
                const int nbytes = kok ? max(0, min(16, (p.na - col) * 8)) : 0;
                const double* src = p.A + (kok ? k : 0) * p.lda + (nbytes > 0 ? col : 0);
                cp_async16_zfill(As + row * JT_LDS + c2, src, nbytes);
            }
            if (!same) {
                const int col = colB + c2;
                const int nbytes = kok ? max(0, min(16, (p.nb - col) * 8)) : 0;
                const double* src = p.B + (kok ? k : 0) * p.ldb + (nbytes > 0 ? col : 0);
                cp_async16_zfill(Bs + row * JT_LDS + c2, src, nbytes);
            }
        }
        if (do_f && tid < JT_KC) {
            const int64_t k = k0 + tid;
            cp_async8_zfill(fs + tid, p.f + (k < k_hi ? k : 0), k < k_hi ? 8 : 0);
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    double jf = 0.0;

    for (int s = 0; s < JT_ST - 1; ++s) {
        if (s < n_it) load_stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fk = lane & 3, fr = lane >> 2;
    for (int it = 0; it < n_it; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(JT_ST - 2) : "memory");
        __syncthreads();                                  // stage `it` has landed for every thread; stage it-1 is free
        {
            const int nx = it + JT_ST - 1;
            if (nx < n_it) load_stage(nx % JT_ST, nx);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const double* As = sm_atb + (size_t)(it % JT_ST) * JT_STAGE_DOUBLES;
        const double* Bs = same ? As : As + JT_KC * JT_LDS;
        const double* fs = As + 2 * JT_KC * JT_LDS;
#pragma unroll
        for (int kk = 0; kk < JT_KC / 4; ++kk) {
            const double* ap = As + (kk * 4 + fk) * JT_LDS + wm + fr;
            const double* bp = Bs + (kk * 4 + fk) * JT_LDS + wn + fr;
            double a[4], b[8];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = ap[8 * mt];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) b[nt] = bp[8 * nt];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        if (do_f && tid < JT_T) {
#pragma unroll
            for (int r = 0; r < JT_KC; ++r) jf = fma(As[r * JT_LDS + tid], fs[r], jf);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    double* pt = p.part + ((size_t)slice * p.n_tiles + tile) * (JT_T * JT_T);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<double2*>(pt + (wm + 8 * mt + fr) * JT_T + wn + 8 * nt + 2 * fk) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    if (do_f && tid < JT_T) p.part_f[(size_t)slice * p.n_bi * JT_T + colA + tid] = jf;
}

// C[i][j] = sum over slices (in slice order) of the partial tiles; tri: the upper triangle is the mirror image (bitwise symmetric).
// C is [na x nb] row-major with row stride ldc; beta = 1 adds to what C holds.
__global__ void __launch_bounds__(256)
k_atb_reduce(AtbArgs p, double* __restrict__ C, int64_t ldc, double* __restrict__ atf, int accumulate)
{
    const int tile = blockIdx.x;
    int bi, bj;
    atb_tile_of(p, tile, bi, bj);
    for (int e = threadIdx.x; e < JT_T * JT_T; e += blockDim.x) {
        const int m = e >> 7, n = e & 127;
        const int i = bi * JT_T + m, j = bj * JT_T + n;
        if (i >= p.na || j >= p.nb) continue;
        if (p.tri && j > i) continue;
        double s = 0.0;
        for (int sl = 0; sl < p.n_slices; ++sl) s += p.part[((size_t)sl * p.n_tiles + tile) * (JT_T * JT_T) + e];
        if (accumulate) s += C[(size_t)i * ldc + j];
        C[(size_t)i * ldc + j] = s;
        if (p.tri && j < i) C[(size_t)j * ldc + i] = s;
    }
    if (atf && p.f && (p.tri ? bi == bj : bj == 0)) {
        for (int m = threadIdx.x; m < JT_T; m += blockDim.x) {
            const int i = bi * JT_T + m;
            if (i >= p.na) continue;
            double s = 0.0;
            for (int sl = 0; sl < p.n_slices; ++sl) s += p.part_f[(size_t)sl * p.n_bi * JT_T + i];
            atf[i] = s;
        }
    }
}

// out[c] = sum_el w[el] X[el][c]   (X [nE x C] row-major; the `w_h . hprobs` half of an MLE Hessian block).  Two passes, both
// deterministic: row slices -> partial sums [n_slices][C], then k_wcolsum_reduce.
__global__ void __launch_bounds__(256)
k_wcolsum(const double* __restrict__ X, int64_t C, int64_t nE, const double* __restrict__ w, int64_t rows_per_slice,
          double* __restrict__ part)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t k_lo = blockIdx.y * rows_per_slice, k_hi = min(k_lo + rows_per_slice, nE);
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    int64_t k = k_lo;
    for (; k + 1 < k_hi; k += 2) { s0 = fma(w[k], X[k * C + c], s0); s1 = fma(w[k + 1], X[(k + 1) * C + c], s1); }
    if (k < k_hi) s0 = fma(w[k], X[k * C + c], s0);
    part[(size_t)blockIdx.y * C + c] = s0 + s1;
}
__global__ void __launch_bounds__(256)
k_wcolsum_reduce(const double* __restrict__ part, int64_t C, int n_slices, double* __restrict__ out)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int sl = 0; sl < n_slices; ++sl) s += part[(size_t)sl * C + c];
    out[c] = s;
}
