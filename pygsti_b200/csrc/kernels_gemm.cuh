// kernels_gemm.cuh -- C = A . B on the FP64 tensor cores (DMMA) with a block-sparse right-hand side: the contraction of a
// per-element derivative array with a member-derivative map.
//
//   J  = W  . D         general Jacobian path (D not a permutation: CPTPLND, H+S, TP-POVM, ... -- matrixforwardsim.py:126-170
//                       `_doperation` contracted as in `_dprobs_from_rho_e` :1059-1139)
//   H += Wp . D[:, p2]  first-order part of a Hessian rectangle, one tangent direction per batch entry
//   H  = W  . D2        second-derivative part (d2M/dtheta_a dtheta_b from member.hessian_wrt_params, matrixforwardsim.py:172-218)
//
// In round 1 these were one-thread-per-output scalar gather-dots over HBM-resident W rows (k_contract_csc, k_contract_hess,
// k_hess_d2): 60 ms for the BASELINE config 4 Jacobian whose W costs 0.07 ms to produce.  The maps are block-dense -- a
// Lindblad gate's 240 parameters touch all 256 elements of that gate and nothing else -- so the right-hand side is
// stored DENSE [K x N] and every 128-column tile of it carries the list of 16-row K chunks that contain a non-zero; a CTA
// only walks those chunks.  A is row-major (K contiguous), exactly the "row" operand of mma.sync.m8n8k4.f64; B is [k][n],
// the "col" operand.  CTA tile 128 x 128, 8 warps x (32 x 64), 3-stage cp.async pipeline, padded strides 20 / 132
// (= 4 mod 16: conflict-free fragment loads).  Optional per-row scale (objective-function Jacobian) and accumulate.
#pragma once
#include "common.cuh"
#include "kernels_jtj.cuh"     // cp_async16_zfill

#define GM_TM 128
#define GM_TN 128
#define GM_KC 16
#define GM_ST 3
#define GM_LDA 20
#define GM_LDB 132
#define GM_STAGE_DOUBLES (GM_TM * GM_LDA + GM_KC * GM_LDB)

struct AbArgs {
    const double* A; int64_t lda; int64_t strideA;   // [M x K] row-major; batch entry z at A + z * strideA
    const double* B; int64_t ldb;                    // dense [K x N] row-major (shared by the batch)
    double* C; int64_t ldc; int64_t strideC;         // [M x N]; batch entry z at C + z * strideC
    int64_t M; int N; int K;
    const int32_t* kt_ptr;                           // [n_tiles_n + 1] -> kt_idx: K chunks (of GM_KC rows) with non-zeros in that column tile
    const int32_t* kt_idx;
    const double* row_scale;                         // [M] or nullptr
    int accumulate;
};

__global__ void __launch_bounds__(256, 1)
k_ab_dmma(AbArgs p)
{
    extern __shared__ __align__(16) double sm_ab[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * GM_TM;
    const int n0 = blockIdx.y * GM_TN;
    const double* A = p.A + (size_t)blockIdx.z * p.strideA;
    double* C = p.C + (size_t)blockIdx.z * p.strideC;
    const int32_t* kl = p.kt_idx + p.kt_ptr[blockIdx.y];
    const int n_it = p.kt_ptr[blockIdx.y + 1] - p.kt_ptr[blockIdx.y];

    auto load_stage = [&](int st, int it) {
        double* As = sm_ab + (size_t)st * GM_STAGE_DOUBLES;
        double* Bs = As + GM_TM * GM_LDA;
        const int k0 = kl[it] * GM_KC;
        // A tile: 128 rows x 16 doubles = 1024 16-byte chunks; B tile: 16 rows x 128 doubles = 1024 chunks
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int chunk = tid + c * 256;
            {
                const int r = chunk >> 3, k2 = (chunk & 7) * 2;
                const int64_t m = m0 + r;
                const int nbytes = (m < p.M) ? max(0, min(16, (p.K - (k0 + k2)) * 8)) : 0;
                const double* src = A + (nbytes > 0 ? m * p.lda + k0 + k2 : 0);
                cp_async16_zfill(As + r * GM_LDA + k2, src, nbytes);
            }
            {
                const int r = chunk >> 6, c2 = (chunk & 63) * 2;
                const int k = k0 + r, col = n0 + c2;
                const int nbytes = (k < p.K) ? max(0, min(16, (p.N - col) * 8)) : 0;
                const double* src = p.B + (nbytes > 0 ? (int64_t)k * p.ldb + col : 0);
                cp_async16_zfill(Bs + r * GM_LDB + c2, src, nbytes);
            }
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    for (int s = 0; s < GM_ST - 1; ++s) {
        if (s < n_it) load_stage(s, s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fk = lane & 3, fr = lane >> 2;
    for (int it = 0; it < n_it; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(GM_ST - 2) : "memory");
        __syncthreads();
        {
            const int nx = it + GM_ST - 1;
            if (nx < n_it) load_stage(nx % GM_ST, nx);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const double* As = sm_ab + (size_t)(it % GM_ST) * GM_STAGE_DOUBLES;
        const double* Bs = As + GM_TM * GM_LDA;
#pragma unroll
        for (int kk = 0; kk < GM_KC / 4; ++kk) {
            const double* ap = As + (wm + fr) * GM_LDA + kk * 4 + fk;
            const double* bp = Bs + (kk * 4 + fk) * GM_LDB + wn + fr;
            double a[4], b[8];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = ap[8 * mt * GM_LDA];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) b[nt] = bp[8 * nt];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    const bool vec = ((p.ldc & 1) == 0) && ((p.strideC & 1) == 0) && (((uintptr_t)p.C & 15) == 0);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int64_t m = m0 + wm + 8 * mt + fr;
        if (m >= p.M) continue;
        const double sc = p.row_scale ? __ldg(p.row_scale + m) : 1.0;
        double* Cr = C + m * p.ldc;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int col = n0 + wn + 8 * nt + 2 * fk;
            double v0 = acc[mt][nt][0] * sc, v1 = acc[mt][nt][1] * sc;
            if (col + 1 < p.N && vec) {
                double2* dst = reinterpret_cast<double2*>(Cr + col);
                if (p.accumulate) { const double2 o = *dst; v0 += o.x; v1 += o.y; }
                *dst = make_double2(v0, v1);
            } else {
                if (col < p.N) Cr[col] = p.accumulate ? Cr[col] + v0 : v0;
                if (col + 1 < p.N) Cr[col + 1] = p.accumulate ? Cr[col + 1] + v1 : v1;
            }
        }
    }
}

// dense [K x ldb] from CSC columns `cols[0..n)` (or all columns when cols == nullptr): B[crow[t]][c] = cval[t].  B is zeroed first.
__global__ void k_csc_to_dense(const int32_t* __restrict__ cptr, const int32_t* __restrict__ crow, const double* __restrict__ cval,
                               const int32_t* __restrict__ cols, int n, double* __restrict__ B, int64_t ldb)
{
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const int pcol = cols ? cols[c] : c;
        for (int t = cptr[pcol] + threadIdx.x; t < cptr[pcol + 1]; t += blockDim.x) B[(int64_t)crow[t] * ldb + c] = cval[t];
    }
}
// dense from COO with distinct (row, col) pairs: B[rows[t]][cols[t]] = vals[t]
__global__ void k_coo_to_dense(int64_t nnz, const int32_t* __restrict__ rows, const int64_t* __restrict__ cols,
                               const double* __restrict__ vals, double* __restrict__ B, int64_t ldb)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nnz; t += (int64_t)gridDim.x * blockDim.x)
        B[(int64_t)rows[t] * ldb + cols[t]] = vals[t];
}
