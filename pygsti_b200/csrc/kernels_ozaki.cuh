// kernels_ozaki.cuh -- J^T J in FP64 accuracy on the 5th-generation tensor cores (tcgen05.mma, kind::i8, TMEM accumulators).
//
// tcgen05.mma has no FP64 kind.  The Ozaki splitting gets FP64 results out of exact integer products:
//   * every column p of J (= row of J^T) is scaled by 2^-e_p so that |x| < 1 (e_p from the column's largest magnitude) and cut
//     into T signed digits of 6 + 7 (T-1) bits:  x = sum_k d_k 2^(-6-7k),  d_k in [-64, 64]  (k_oz_slice; every step is exact
//     in FP64: multiplications by powers of two, round-to-nearest-integer, subtraction of that integer);
//   * (J^T J)[i][j] 2^-(e_i+e_j) = sum_{a,b} 2^(-12-7(a+b)) sum_el d_a[i][el] d_b[j][el]; the inner sums are int8 x int8 -> int32 GEMMs,
//     EXACT as long as (pairs per level) x (rows per K slice) x 64^2 < 2^31, and pairs with a + b >= T are below 2^-(6+7T) (T = 8:
//     2^-62; T = 7: 2^-55) relative to the product of the column maxima.
//
// Kernel k_oz_syrk<T>: one CTA per (output tile 128 x 64 of the lower triangle, K slice).  ALL T level accumulators (level t = a + b,
// 64 TMEM columns each: T x 64 <= 512 = the whole tensor memory of the SM) stay resident for the whole K slice, so a pipeline stage
// carries the T digit tiles of both operands ONCE (T x (128 + 64) rows x 32 K-bytes = 48 KB at T = 8) and feeds all T (T+1) / 2 = 36
// digit pairs from it: 196 MAC per operand byte.  (A first version kept one 128 x 256 accumulator per level and re-streamed both
// operands for every digit pair: 87 MAC/B = 136 GB through L2 per J^T J of config 2 -- L2-feed-bound at the DMMA kernel's speed.)
//   * digits live in global memory as the exact shared-memory image the tensor core wants (K-major, no swizzle: 8 x 16 B core matrices,
//     S[k stage][digit][row group of 8][K chunk of 16][row & 7][16 B]), so one digit tile of a stage is ONE contiguous run of 4 KB /
//     2 KB: the producer thread moves a stage with 2 T bulk copies of the TMA engine (cp.async.bulk ... mbarrier::complete_tx) -- no
//     tensor map, no address arithmetic in the other warps;
//   * warp 0 / lane 0 = producer, warp 1 / lane 0 = MMA issuer (tcgen05.commit frees the stage), 4 stages; after the last stage all four
//     warps drain the accumulators once: tcgen05.ld 32x32b, sum_t 2^(-12-7t) acc_t in FP64 from the smallest level up, one FP64 partial
//     tile per CTA;
//   * k_oz_reduce sums the partial tiles in slice order, scales by 2^(e_i + e_j) and mirrors: deterministic, bitwise symmetric.
#pragma once
#include "common.cuh"

#define OZ_TM 128
#define OZ_TN 64
#define OZ_KS 32                       // K elements (= bytes) per pipeline stage = one tcgen05.mma kind::i8
#define OZ_NST 4
#define OZ_LBO 128u                    // between the two 16-byte K chunks of a core-matrix row group
#define OZ_SBO 256u                    // between row groups of 8
#define OZ_MAX_STAGES_PER_SLICE 2047   // 8 pairs x 65 504 rows x 64^2 < 2^31
#define OZ_BAD_COLUMN 0x7fffffff       // exponent marker of a column with non-finite entries

// ---- one pass over J: per column the largest magnitude (-> exponent e_p) and, with f, the weighted sum (J^T f)[p] -----------------
// row slices -> partials [n_slices][Np], reduced in slice order by k_oz_colstats_reduce (deterministic)
__global__ void __launch_bounds__(256)
k_oz_colstats(const double* __restrict__ J, int64_t ld, int64_t nE, int Np, const double* __restrict__ f,
              double* __restrict__ part_sum, double* __restrict__ part_max)
{
    // block = 64 columns x 4 row lanes; row lane te of block row by takes rows by * 4 + te + i * (4 gridDim.y): many rows in flight per
    // column (the first version -- one thread per column walking a contiguous row slice, 4 loads in flight -- reached 4.8 TB/s)
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int64_t r0 = (int64_t)blockIdx.y * 4 + (threadIdx.x >> 6), rs = (int64_t)gridDim.y * 4;
    if (c >= Np) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, m0 = 0.0, m1 = 0.0;
    bool bad = false;
    int64_t k = r0;
    for (; k + 3 * rs < nE; k += 4 * rs) {
        const double v0 = J[k * ld + c], v1 = J[(k + rs) * ld + c], v2 = J[(k + 2 * rs) * ld + c], v3 = J[(k + 3 * rs) * ld + c];
        if (f) { s0 = fma(f[k], v0, s0); s1 = fma(f[k + rs], v1, s1); s2 = fma(f[k + 2 * rs], v2, s2); s3 = fma(f[k + 3 * rs], v3, s3); }
        m0 = fmax(m0, fmax(fabs(v0), fabs(v1))); m1 = fmax(m1, fmax(fabs(v2), fabs(v3)));
        bad |= !(fabs(v0) <= 1.7976931348623157e308) | !(fabs(v1) <= 1.7976931348623157e308) | !(fabs(v2) <= 1.7976931348623157e308) | !(fabs(v3) <= 1.7976931348623157e308);
    }
    for (; k < nE; k += rs) { const double v = J[k * ld + c]; if (f) s0 = fma(f[k], v, s0); m0 = fmax(m0, fabs(v)); bad |= !(fabs(v) <= 1.7976931348623157e308); }
    const size_t slot = (size_t)r0 * Np + c;                       // partial slot = by * 4 + te
    part_sum[slot] = (s0 + s1) + (s2 + s3);
    part_max[slot] = bad ? __longlong_as_double(0x7ff0000000000000LL) : fmax(m0, m1);     // (fmax drops NaN: Inf / NaN entries are flagged explicitly)
}
// e_p = smallest e with max_el |J[el][p]| < 2^e  (0 for an all-zero column)
__global__ void __launch_bounds__(256)
k_oz_colstats_reduce(const double* __restrict__ part_sum, const double* __restrict__ part_max, int Np, int n_slices,
                     double* __restrict__ jtf, int* __restrict__ expo)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= Np) return;
    double s = 0.0, m = 0.0;
    for (int sl = 0; sl < n_slices; ++sl) { s += part_sum[(size_t)sl * Np + c]; m = fmax(m, part_max[(size_t)sl * Np + c]); }
    if (jtf) jtf[c] = s;
    int e = 0;
    if (m > 0.0 && isfinite(m)) { frexp(m, &e); }                                  // m = f 2^e, f in [0.5, 1)  =>  |x| 2^-e < 1
    // a column holding Inf / NaN (flagged by k_oz_colstats as an infinite maximum) cannot be cut into digits: it is marked, sliced as zeros, and
    // its row and column of J^T J come out as NaN -- what the FP64 contraction would deliver for it
    expo[c] = isfinite(m) ? max(e, -900) : OZ_BAD_COLUMN;
}

// ---- digits, written as the operand image:  S[stage][t][p >> 3][chunk][p & 7][16],  zero padded to P_pad x n_stages ----------
// One thread = one column p x 16 consecutive elements = one 16-byte chunk of the image per digit: no shared memory.  The loads of a
// warp are 256 contiguous bytes of a Jacobian row; 8 neighbouring columns write 128 contiguous bytes.  Digit t is read off the low
// mantissa bits of y + 1.5 2^(52 - 7t) (round-to-nearest at 2^-7t in one DADD), the remainder y - d_t 2^-7t is exact.
template <int T>
__global__ void __launch_bounds__(256)
k_oz_slice(const double* __restrict__ J, int64_t ld, int64_t nE, int Np, const int* __restrict__ expo,
           int8_t* __restrict__ S, int64_t P_pad, int64_t n_stages)
{
    const int p = blockIdx.x * 64 + (threadIdx.x & 63);
    const int c4 = threadIdx.x >> 6;                          // which 16 of the block's 64 elements: stage 2 by + (c4 >> 1), chunk c4 & 1
    const int64_t el0 = (int64_t)blockIdx.y * 64 + c4 * 16;
    const int64_t stage = (int64_t)blockIdx.y * 2 + (c4 >> 1);
    if (stage >= n_stages) return;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (p < Np && el0 + i < nE) ? J[(el0 + i) * ld + p] : 0.0;
    const int ex = (p < Np) ? expo[p] : 0;
    const double sc = ex == OZ_BAD_COLUMN ? 0.0 : scalbn(1.0, 6 - ex);     // |x| sc < 64
    uint32_t w[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) { w[t][0] = 0u; w[t][1] = 0u; w[t][2] = 0u; w[t][3] = 0u; }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double y = ex == OZ_BAD_COLUMN ? 0.0 : x[i] * sc;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const double magic = 6755399441055744.0 / (double)(1ull << (7 * t));     // 1.5 2^(52 - 7t)
            const double m = y + magic;
            w[t][i >> 2] |= ((uint32_t)__double2loint(m) & 0xFFu) << (8 * (i & 3));
            y -= (m - magic);                                 // |remainder| <= 2^(-7t-1): the next digit is again in [-64, 64]
        }
    }
    int8_t* dst = S + (stage * T * P_pad + (p & ~7)) * OZ_KS + (c4 & 1) * 128 + (p & 7) * 16;
#pragma unroll
    for (int t = 0; t < T; ++t)
        *reinterpret_cast<int4*>(dst + (int64_t)t * P_pad * OZ_KS) = make_int4((int)w[t][0], (int)w[t][1], (int)w[t][2], (int)w[t][3]);
}

// ---- tcgen05 / mbarrier / bulk-copy helpers (forms as in CUTLASS cute/arch/{mma_sm100_umma,tmem_allocator_sm100,copy_sm90_tma}.hpp) --
__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = oz_smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(oz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {      // K-major, SWIZZLE_NONE, version 1 (Blackwell)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(OZ_LBO >> 4) << 16) | ((uint64_t)(OZ_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct OzArgs {
    const int8_t* S; int64_t P_pad;                  // digits [n_stages][T][P_pad x 32 bytes in core-matrix order]
    const int2* tiles; int n_tiles;                  // (row block of 128, column block of 64) with a lower-triangle part
    int n_kslices, stages_per_slice;                 // K slice = stages_per_slice stages of OZ_KS elements
    double* part;                                    // [n_kslices][n_tiles][128 * 64]
};

// instruction descriptor: D = S32 (bits 4-5 = 2), A = B = signed int8 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ constexpr uint32_t oz_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_TM >> 4) << 24);
}

template <int T>
__global__ void __launch_bounds__(128, 1)
k_oz_syrk(OzArgs p)
{
    constexpr uint32_t A_T = OZ_TM * OZ_KS, B_T = OZ_TN * OZ_KS, STAGE = T * (A_T + B_T);
    constexpr uint32_t TMEM_COLS = 512;                                  // T x 64 rounded up to a power of two (T = 7, 8)
    extern __shared__ __align__(1024) uint8_t oz_smem[];
    __shared__ uint64_t full_bar[OZ_NST], empty_bar[OZ_NST];
    __shared__ uint64_t acc_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x % p.n_tiles, ksl = blockIdx.x / p.n_tiles;
    const int2 tb = p.tiles[tile];
    const int64_t rowA0 = (int64_t)tb.x * OZ_TM, rowB0 = (int64_t)tb.y * OZ_TN;
    const int n_st = p.stages_per_slice;
    const int64_t st0 = (int64_t)ksl * n_st;

    if (tid == 0) {
        for (int s = 0; s < OZ_NST; ++s) { oz_mbar_init(&full_bar[s], 1); oz_mbar_init(&empty_bar[s], 1); }
        oz_mbar_init(&acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t smem0 = oz_smem_u32(oz_smem);

    if (warp == 0) {
        if (lane == 0) {                                   // ---- producer: 2 T bulk copies per stage ----
            for (int n = 0; n < n_st; ++n) {
                const uint32_t slot = (uint32_t)n % OZ_NST, use = (uint32_t)n / OZ_NST;
                oz_mbar_wait(&empty_bar[slot], (use & 1u) ^ 1u);
                oz_mbar_expect_tx(&full_bar[slot], STAGE);
                const int8_t* src = p.S + (st0 + n) * T * p.P_pad * OZ_KS;
                const uint32_t sbase = smem0 + slot * STAGE;
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    oz_bulk_g2s(sbase + t * A_T, src + ((int64_t)t * p.P_pad + rowA0) * OZ_KS, A_T, &full_bar[slot]);
                    oz_bulk_g2s(sbase + T * A_T + t * B_T, src + ((int64_t)t * p.P_pad + rowB0) * OZ_KS, B_T, &full_bar[slot]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer: level t gets the pairs (a, t - a) ----
            for (int n = 0; n < n_st; ++n) {
                const uint32_t slot = (uint32_t)n % OZ_NST, use = (uint32_t)n / OZ_NST;
                oz_mbar_wait(&full_bar[slot], use & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = oz_desc(smem0 + slot * STAGE), db = oz_desc(smem0 + slot * STAGE + T * A_T);
                // digit a of the rows against digits 0 .. T-1-a of the columns, STACKED along N (the column digit tiles are contiguous
                // row groups in shared memory and level a + b is accumulator column 64 (a + b)): one instruction of N = 64 (T - a)
                // (<= 256 per instruction) instead of T - a instructions of N = 64, which re-read the 4 KB row tile each time and
                // were shared-memory bound (ncu: tensor pipe 60 % busy, L1/shared 89 %).  T = 8: 12 instructions per stage.
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int off = 0; off < OZ_TN * (T - a); off += 256) {
                        const int nn = (OZ_TN * (T - a) - off) < 256 ? (OZ_TN * (T - a) - off) : 256;
                        oz_mma_i8(tmem + (uint32_t)(a * OZ_TN + off), da + (uint64_t)((a * A_T) >> 4), db + (uint64_t)((off * OZ_KS) >> 4),
                                  oz_idesc(nn), (n > 0 || a > 0) ? 1u : 0u);
                    }
                oz_commit(&empty_bar[slot]);
            }
            oz_commit(&acc_done);
        }
        __syncwarp();
    }

    // ---- drain: partial[row][col] = sum_t 2^(-12-7t) acc_t[row][col], smallest level first ----
    oz_mbar_wait(&acc_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        double* pr = p.part + ((size_t)ksl * p.n_tiles + tile) * (OZ_TM * OZ_TN) + (size_t)(warp * 32 + lane) * OZ_TN;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            double acc[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) acc[k] = 0.0;
#pragma unroll 1
            for (int t = T - 1; t >= 0; --t) {
                uint32_t v[32];
                oz_tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * OZ_TN + half * 32), v);
                const double w = scalbn(1.0, -12 - 7 * t);
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = fma((double)(int)v[k], w, acc[k]);
            }
#pragma unroll
            for (int k = 0; k < 32; k += 2) *reinterpret_cast<double2*>(pr + half * 32 + k) = make_double2(acc[k], acc[k + 1]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

// C[i][j] = 2^(e_i + e_j) sum over K slices (in order) of the partial tiles, lower triangle + mirror
__global__ void __launch_bounds__(256)
k_oz_reduce(OzArgs p, const int* __restrict__ expo, int Np, double* __restrict__ C, int64_t ldc)
{
    const int tile = blockIdx.x;
    const int2 tb = p.tiles[tile];
    for (int e = threadIdx.x; e < OZ_TM * OZ_TN; e += blockDim.x) {
        const int m = e / OZ_TN, n = e - m * OZ_TN;
        const int i = tb.x * OZ_TM + m, j = tb.y * OZ_TN + n;
        if (i >= Np || j > i) continue;
        double s = 0.0;
        for (int sl = 0; sl < p.n_kslices; ++sl) s += p.part[((size_t)sl * p.n_tiles + tile) * (OZ_TM * OZ_TN) + e];
        s = (expo[i] == OZ_BAD_COLUMN || expo[j] == OZ_BAD_COLUMN) ? __longlong_as_double(0x7ff8000000000000LL) : scalbn(s, expo[i] + expo[j]);
        C[(size_t)i * ldc + j] = s;
        if (j < i) C[(size_t)j * ldc + i] = s;
    }
}
