// kernels_factoredj.cuh -- analytic Jacobian for gates held as FACTOR PROGRAMS (products of small operations embedded on 1-2 qubits).
//
// The level-batched d >= 64 Jacobian (kernels_levelj.cuh) treats every layer operation as a dense d x d matrix: two sweeps of dense
// products that keep every level (the 26 GB adjoint table of BASELINE config 3) and d x d outer-product accumulators per gate.  But a
// layer of a local-noise / crosstalk-free model IS a product of 4 x 4 / 16 x 16 operations embedded on 1-2 qubits (the reference's
// OpCRep_Composed of OpCRep_Embedded, opcreps.cpp:242-276, 93-158), and its parameters live in those small matrices.  With
// F = I (x) g (x) I acting on the target digits `a` of the state index (rest digits `r` untouched):
//     d p / d g[a][b] = sum over the steps k that apply this factor  sum_r  e_k(a, r) s_{k-1}(b, r)
//     e_{k-1}(a', r)  = sum_a g[a][a'] e_k(a, r)                                   (backward vector through F^T)
// -- 4 or 16 multiply-adds per state component instead of d, small accumulators (16 / 256 doubles per factor) instead of d x d,
// and no adjoint table at all: the backward vector is consumed where it is produced.
//   k_fj_count     steps (factors) per circuit and the circuit of every outcome slot
//   k_fj_forward   one warp per circuit: state after every factor -> FS rows (D doubles, coalesced); probabilities
//   k_fj_backward  one warp per (circuit, outcome): walks the factors backwards.  Both products of a step run on the FP64 tensor
//                  cores (DMMA m8n8k4): accumulate  acc[a][b] += E(a, r) S(b, r)^T  (M = N = 4^nq, K = rest) and
//                  chain  e'(a', r) = F^T[a'][a] e(a, r)  (M = K = 4^nq, N = rest); the fragments are read straight from the
//                  swizzled shared-memory state (the same lane -> index map serves the A fragment of e and the B fragment of s),
//                  the F^T fragments from a per-CTA image built once, the accumulators live in shared memory in fragment order
//                  (one 16-byte read-modify-write per tile and lane).  The s row of the next step is requested a step ahead.
//                  Epilogue: J[el][p] = sum over the column's entries  D_f[w][p] W_f[w]  (CSC over the FACTOR-space elements:
//                  factor matrix entries, rho (= e_0), effect (= s_L)); one coalesced row store, row scale applied there.
#pragma once
#include "common.cuh"
#include "kernels_factored.cuh"

struct FjDev {
    const uint32_t* base;      // [n_circ + 1] first FS row of circuit c (steps_c + 1 rows: the state before every factor, and s_L)
    const int32_t* out_circ;   // [n_outcomes] circuit of outcome slot qo
    const uint16_t* step_fac;  // [rows] factor applied AFTER the state of FS row r (the last row of a circuit: unused)
    const int32_t* fao;        // [n_fac] offset (doubles) of factor f's accumulators in a warp's accumulator buffer
    const int32_t* ffo;        // [n_fac] offset (doubles) of factor f's F^T fragments in the CTA's fragment image
    int n_acc, n_frag;
    const int32_t* cptr;       // [n_params + 1]  CSC of the factor-space derivative map
    const uint32_t* ccode;     // [nnz]  kind << 30 | sign << 29 | ...: 0 = accumulator offset; 1 = rho (prep << 16 | index); 2 = effect (effect << 16 | index)
    const double* cval;        // [nnz]
    int n_params;
    int unit;                  // all values are +-1: the sign is bit 29 of the code and cval is not read
};

__global__ void k_fj_count(AtomDev a, const int32_t* __restrict__ fptr, uint32_t* __restrict__ n_rows, int32_t* __restrict__ out_circ)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.n_circ; c += gridDim.x * blockDim.x) {
        uint32_t n = 1;
        for (uint32_t k = a.circ_ptr[c]; k < a.circ_ptr[c + 1]; ++k) { const int g = a.circ_ops[k]; n += (uint32_t)(fptr[g + 1] - fptr[g]); }
        n_rows[c] = n;
        for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) out_circ[qo] = c;
    }
}

__global__ void k_fj_steps(AtomDev a, const int32_t* __restrict__ fptr, const uint32_t* __restrict__ base, uint16_t* __restrict__ step_fac)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.n_circ; c += gridDim.x * blockDim.x) {
        uint32_t r = base[c];
        for (uint32_t k = a.circ_ptr[c]; k < a.circ_ptr[c + 1]; ++k) {
            const int g = a.circ_ops[k];
            for (int f = fptr[g]; f < fptr[g + 1]; ++f) step_fac[r++] = (uint16_t)f;
        }
        step_fac[r] = 0;
    }
}

// One warp per circuit; shared memory as k_probs_factored.
template <int D>
__global__ void __launch_bounds__(FAC_WARPS * 32)
k_fj_forward(AtomDev a, FactoredDev fd, int n_mats, int n_fac, const uint32_t* __restrict__ base, const double* __restrict__ rho,
             const double* __restrict__ E, double* __restrict__ FS, double* __restrict__ probs)
{
    extern __shared__ __align__(16) double smf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* b0 = smf + (size_t)warp * 2 * D;
    double* b1 = b0 + D;
    double* ms = smf + (size_t)FAC_WARPS * 2 * D;
    FactorRec* recs = reinterpret_cast<FactorRec*>(ms + ((n_mats + 1) & ~1));
    int* fptr = reinterpret_cast<int*>(recs + n_fac);
    for (int i = threadIdx.x; i < n_mats; i += blockDim.x) ms[i] = fd.mats[i];
    for (int i = threadIdx.x; i < n_fac; i += blockDim.x) recs[i] = fd.fac[i];
    for (int i = threadIdx.x; i <= a.n_ops; i += blockDim.x) fptr[i] = fd.op_fptr[i];
    __syncthreads();
    const int gw = blockIdx.x * FAC_WARPS + warp, nw = gridDim.x * FAC_WARPS;
    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c], L = a.circ_ptr[c + 1] - p0;
        const double* r = rho + (size_t)a.circ_prep[c] * D;
        double* cur = b0; double* nxt = b1;
        double* row = FS + (size_t)base[c] * D;
        for (int i = lane; i < D; i += 32) { const double v = r[i]; cur[fac_sw((unsigned)i)] = v; row[i] = v; }
        __syncwarp();
        int gnext = L ? __ldg(a.circ_ops + p0) : 0;
        for (uint32_t k = 0; k < L; ++k) {
            const int g = gnext;
            if (k + 1 < L) gnext = __ldg(a.circ_ops + p0 + k + 1);
            const int f0 = fptr[g], f1 = fptr[g + 1];
            for (int f = f0; f < f1; ++f) {
                const FactorRec fr = recs[f];
                apply_factor<D>(fr, ms + fr.moff, cur, nxt, lane);
                __syncwarp();
                double* x = cur; cur = nxt; nxt = x;
                row += D;
                for (int i = lane; i < D; i += 32) row[i] = cur[fac_sw((unsigned)i)];
            }
        }
        if (probs) {
            for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
                const double* e = E + (size_t)a.out_eff[qo] * D;
                double part = 0.0;
                for (int i = lane; i < D; i += 32) part = fma(__ldg(e + i), cur[fac_sw((unsigned)i)], part);
#pragma unroll
                for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
                if (lane == 0) probs[a.out_el[qo]] = part;
            }
        }
        __syncwarp();
    }
}

// swizzle of the factored kernels' shared-memory vectors: a fragment load walks the state with the 16 lanes of a half-warp differing in
// TWO base-4 digits of the index (which two depends on the factor's target qubits).  With the bank (8-byte word of a 128-byte row) a
// GF(2)-linear function of the digits d0..d3 --  bank = (d0 ^ d2 ^ C d3) | (d1 ^ d2 ^ d3) << 2,  C = [[0,1],[1,1]]  -- every pair of digit
// columns forms an invertible 4 x 4 matrix, so ANY two digits map the 16 lanes onto 16 different banks (tests/test_factored_fragments_cpu.py
// checks every target combination of d = 64 and d = 256).  The first swizzle folded digits 2, 3 into digit 1 only and left 4-way conflicts
// for factors whose first target is the last qubit (ncu: 16 % of the backward kernel's shared-memory wavefronts were conflicts).
__device__ __forceinline__ unsigned fj_sw(unsigned i) {
    const unsigned d2 = (i >> 4) & 3u, d3 = (i >> 6) & 3u;
    const unsigned c3 = (d3 >> 1) | (((d3 ^ (d3 >> 1)) & 1u) << 1);
    return i ^ (d2 ^ c3) ^ ((d2 ^ d3) << 2);
}
__device__ __forceinline__ unsigned fj_idx1(unsigned a, unsigned r, int sh) { return fj_sw(fac_insert2(r, sh) | (a << sh)); }
__device__ __forceinline__ unsigned fj_idx2(unsigned a, unsigned r, int lo, int hi, int s0, int s1) {
    return fj_sw(fac_insert2(fac_insert2(r, lo), hi) | ((a >> 2) << s0) | ((a & 3u) << s1));
}

// Shared memory: F^T fragment image [n_frag] | FactorRec [n_fac] | fptr [n_ops + 1] | fao [n_fac] | ffo [n_fac] | per warp: e0 [D], e1 [D],
// s [D], accumulators [n_acc].
template <int D>
__global__ void __launch_bounds__(256)
k_fj_backward(AtomDev a, FactoredDev fd, FjDev fj, int n_fac, const double* __restrict__ E, const double* __restrict__ FS,
              double* __restrict__ J, int64_t ld, const double* __restrict__ row_scale, unsigned* __restrict__ counter, int n_items)
{
    extern __shared__ __align__(16) double smj[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    double* frag = smj;
    FactorRec* recs = reinterpret_cast<FactorRec*>(frag + ((fj.n_frag + 1) & ~1));
    int* fptr = reinterpret_cast<int*>(recs + n_fac);
    int* fao = fptr + (a.n_ops + 1);
    int* ffo = fao + n_fac;
    double* wbase = smj + (((fj.n_frag + 1) & ~1) + n_fac * 2 + ((a.n_ops + 1 + 2 * n_fac + 3) >> 2) * 2) + warp * (3 * D + fj.n_acc);   // (offsets from smj keep the shared address space)
    double* eb0 = wbase; double* eb1 = wbase + D; double* sb = wbase + 2 * D; double* acc = wbase + 3 * D;

    for (int i = threadIdx.x; i < n_fac; i += blockDim.x) { recs[i] = fd.fac[i]; fao[i] = fj.fao[i]; ffo[i] = fj.ffo[i]; }
    for (int i = threadIdx.x; i <= a.n_ops; i += blockDim.x) fptr[i] = fd.op_fptr[i];
    for (int f = warp; f < n_fac; f += n_warps) {               // F^T fragments: A[row = a'][col = a] = F[a][a']
        const FactorRec fr = fd.fac[f];
        const double* m = fd.mats + fr.moff;
        double* dst = frag + fj.ffo[f];
        if (fr.nq == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {                         // q = mt * 4 + kk
                const int mt = q >> 2, kk = q & 3;
                dst[q * 32 + lane] = m[(4 * kk + (int)lt) * 16 + 8 * mt + (int)lg];
            }
        } else {
            dst[lane] = m[(int)lt * 4 + (int)(lg & 3u)];
        }
    }
    __syncthreads();

    const int Np = fj.n_params;
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int c = fj.out_circ[item];
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const uint32_t row0 = fj.base[c];
        uint32_t t = fj.base[c + 1] - row0 - 1;                    // factor steps of the circuit
        const int eff = a.out_eff[item];
        const int64_t el = a.out_el[item];
        const double* srow = FS + (size_t)row0 * D;
        double* cur = eb0; double* nxt = eb1;
        for (int i = lane; i < fj.n_acc; i += 32) acc[i] = 0.0;
        for (int i = lane; i < D; i += 32) cur[fj_sw((unsigned)i)] = __ldg(E + (size_t)eff * D + i);
        double sreg[D / 32];
        if (t > 0) {
#pragma unroll
            for (int j = 0; j < D / 32; ++j) sreg[j] = __ldg(srow + (size_t)(t - 1) * D + lane + 32 * j);
        }
        __syncwarp();
        for (int k = L - 1; k >= 0; --k) {
            const int g = __ldg(a.circ_ops + p0 + k);
            const int f0 = fptr[g];
            for (int f = fptr[g + 1] - 1; f >= f0; --f) {
                --t;                                               // this step: factor f between s_t (before) and e (after)
#pragma unroll
                for (int j = 0; j < D / 32; ++j) sb[fj_sw((unsigned)(lane + 32 * j))] = sreg[j];
                __syncwarp();
                if (t > 0) {
#pragma unroll
                    for (int j = 0; j < D / 32; ++j) sreg[j] = __ldg(srow + (size_t)(t - 1) * D + lane + 32 * j);
                }
                const FactorRec fr = recs[f];
                if (fr.nq == 2) {
                    constexpr int R = D / 16;                      // rest values per target index
                    const int s0 = fr.shift[0], s1 = fr.shift[1];
                    const int lo = s0 < s1 ? s0 : s1, hi = s0 < s1 ? s1 : s0;
                    // accumulate: acc[a][b] += sum_r e(a, r) s(b, r)
                    double2* ap = reinterpret_cast<double2*>(acc + fao[f]) + lane;
                    double2 c00 = ap[0], c01 = ap[32], c10 = ap[64], c11 = ap[96];
#pragma unroll
                    for (int kk = 0; kk < R / 4; ++kk) {
                        const unsigned r = 4u * kk + lt;
                        const unsigned i0 = fj_idx2(lg, r, lo, hi, s0, s1), i1 = fj_idx2(lg + 8u, r, lo, hi, s0, s1);
                        const double a0 = cur[i0], a1 = cur[i1], q0 = sb[i0], q1 = sb[i1];
                        dmma884(c00.x, c00.y, a0, q0); dmma884(c01.x, c01.y, a0, q1);
                        dmma884(c10.x, c10.y, a1, q0); dmma884(c11.x, c11.y, a1, q1);
                    }
                    ap[0] = c00; ap[32] = c01; ap[64] = c10; ap[96] = c11;
                    // chain: e'(a', r) = sum_a F[a][a'] e(a, r)
                    const double* fg = frag + ffo[f] + lane;
                    double A0[4], A1[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { A0[kk] = fg[kk * 32]; A1[kk] = fg[(4 + kk) * 32]; }
#pragma unroll
                    for (int nt = 0; nt < (R + 7) / 8; ++nt) {
                        const unsigned rB = (8u * nt + lg) & (unsigned)(R - 1);      // lanes beyond R re-read a valid column
                        double2 o0 = make_double2(0.0, 0.0), o1 = make_double2(0.0, 0.0);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const double b = cur[fj_idx2(4u * kk + lt, rB, lo, hi, s0, s1)];
                            dmma884(o0.x, o0.y, A0[kk], b); dmma884(o1.x, o1.y, A1[kk], b);
                        }
                        const unsigned rO = 8u * nt + 2u * lt;
                        if (rO < (unsigned)R) {
                            nxt[fj_idx2(lg, rO, lo, hi, s0, s1)] = o0.x; nxt[fj_idx2(lg, rO + 1u, lo, hi, s0, s1)] = o0.y;
                            nxt[fj_idx2(lg + 8u, rO, lo, hi, s0, s1)] = o1.x; nxt[fj_idx2(lg + 8u, rO + 1u, lo, hi, s0, s1)] = o1.y;
                        }
                    }
                } else {
                    constexpr int R = D / 4;
                    const int sh = fr.shift[0];
                    double2 cc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int kk = 0; kk < R / 4; ++kk) {
                        const unsigned ix = fj_idx1(lg & 3u, 4u * kk + lt, sh);
                        dmma884(cc.x, cc.y, cur[ix], sb[ix]);
                    }
                    if (lg < 4u && lt < 2u) {
                        double2* ap = reinterpret_cast<double2*>(acc + fao[f] + lg * 4u + 2u * lt);
                        double2 v = *ap; v.x += cc.x; v.y += cc.y; *ap = v;
                    }
                    const double A = frag[ffo[f] + lane];
#pragma unroll
                    for (int nt = 0; nt < R / 8; ++nt) {
                        const double b = cur[fj_idx1(lt, 8u * nt + lg, sh)];
                        double2 o = make_double2(0.0, 0.0);
                        dmma884(o.x, o.y, A, b);
                        if (lg < 4u) {
                            const unsigned rO = 8u * nt + 2u * lt;
                            nxt[fj_idx1(lg, rO, sh)] = o.x; nxt[fj_idx1(lg, rO + 1u, sh)] = o.y;
                        }
                    }
                }
                __syncwarp();
                double* x = cur; cur = nxt; nxt = x;
            }
        }
        // ---- epilogue: cur = e_0 (-> rho block), s_L (-> effect block), accumulators (-> factor blocks) ----
        for (int i = lane; i < D; i += 32) sb[fj_sw((unsigned)i)] = __ldg(srow + (size_t)(fj.base[c + 1] - row0 - 1) * D + i);
        __syncwarp();
        const int prep = a.circ_prep[c];
        const double sc = row_scale ? __ldg(row_scale + el) : 1.0;
        double* Jrow = J + el * ld;
        for (int p = lane; p < Np; p += 32) {
            double v = 0.0;
            const int t1 = __ldg(fj.cptr + p + 1);
            for (int q = __ldg(fj.cptr + p); q < t1; ++q) {
                const uint32_t code = __ldg(fj.ccode + q);
                const double val = __ldg(fj.cval + q);
                const uint32_t kind = code >> 30;
                double x;
                if (kind == 0u) x = acc[code & 0x1FFFFFFFu];
                else {
                    const int i = (int)((code >> 16) & 0x1FFFu); const unsigned idx = fj_sw(code & 0xFFFFu);
                    x = (kind == 1u) ? (i == prep ? cur[idx] : 0.0) : (i == eff ? sb[idx] : 0.0);
                }
                v = fma(val, x, v);
            }
            Jrow[p] = v * sc;
        }
        __syncwarp();
    }
}

// =====================================================================================================================
// d = 64 (3 qubits) specialisation.  ncu on the generic backward kernel at BASELINE config 3: 265 issued instructions per factor
// step, 54 % issue-active, DMMA 27 % -- the lane -> state-index arithmetic (two digit insertions + swizzle per fragment element,
// ~14 elements per step) was the kernel.  Those indices depend only on (factor, lane), so each CTA builds them ONCE into a
// shared-memory table: one 16-byte entry per (factor, lane) = the 10 state indices of a step as bytes + the factor's metadata.
// The accumulators of up to 4 two-qubit factors live in REGISTERS (a warp-uniform switch selects the set), which removes the
// 16-byte read-modify-write per tile and step that made up half of the shared-memory wavefronts; a 1-qubit accumulate splits
// the 16 rest values over the two row halves of the 8 x 8 tile (2 DMMA instead of 4, the two diagonal 4 x 4 blocks are summed
// in the epilogue).  The forward kernel runs the same DMMA chain with the F fragments (the scalar factor application was
// shared-memory bound: 97 % L1 pipe).
// Measured and NOT kept: two outcomes of a circuit per warp (shared s / F^T fragments and table entry, both backward vectors in one set
// of 8 chain DMMA with N = (outcome, rest)): 29 % fewer instructions and 26 % fewer shared-memory wavefronts per outcome, but only 10 warps
// per SM fit (2 x accumulators in registers and shared memory) -- 15.9 vs 14.7 ms at config 3.  The kernel sits on a plateau of issue
// (44 %), L1 / shared (77 %) and DMMA (36 %) with every instruction of a warp waiting ~9 cycles at 4 warps per scheduler.
// =====================================================================================================================
#define FJ64_REG_SLOTS 4

struct Fj64Tab { uint4* tab; double* frag; };

// table entry of (factor f, lane): x, y, z = indices ix0..ix9 as bytes, z byte 2 = nq, z byte 3 = register slot (0xFF: shared memory),
// w = accumulator offset | fragment offset << 16
__device__ __forceinline__ uint4 fj64_entry(const FactorRec fr, int fao, int ffo, int slot, unsigned lg, unsigned lt)
{
    unsigned ix[10];
    if (fr.nq == 2) {
        const int s0 = fr.shift[0], s1 = fr.shift[1];
        const int lo = s0 < s1 ? s0 : s1, hi = s0 < s1 ? s1 : s0;
        ix[0] = fj_idx2(lg, lt, lo, hi, s0, s1); ix[1] = fj_idx2(lg + 8u, lt, lo, hi, s0, s1);
        for (unsigned kk = 0; kk < 4; ++kk) ix[2 + kk] = fj_idx2(4u * kk + lt, lg & 3u, lo, hi, s0, s1);
        const unsigned r = 2u * (lt & 1u);
        ix[6] = fj_idx2(lg, r, lo, hi, s0, s1); ix[7] = fj_idx2(lg, r + 1u, lo, hi, s0, s1);
        ix[8] = fj_idx2(lg + 8u, r, lo, hi, s0, s1); ix[9] = fj_idx2(lg + 8u, r + 1u, lo, hi, s0, s1);
    } else {
        const int sh = fr.shift[0];
        ix[0] = fj_idx1(lg & 3u, (lg >> 2) * 8u + lt, sh); ix[1] = fj_idx1(lg & 3u, (lg >> 2) * 8u + 4u + lt, sh);
        ix[2] = ix[3] = 0u;
        ix[4] = fj_idx1(lt, lg, sh); ix[5] = fj_idx1(lt, 8u + lg, sh);
        ix[6] = fj_idx1(lg & 3u, 2u * lt, sh); ix[7] = fj_idx1(lg & 3u, 2u * lt + 1u, sh);
        ix[8] = fj_idx1(lg & 3u, 8u + 2u * lt, sh); ix[9] = fj_idx1(lg & 3u, 9u + 2u * lt, sh);
    }
    uint4 e;
    e.x = ix[0] | (ix[1] << 8) | (ix[2] << 16) | (ix[3] << 24);
    e.y = ix[4] | (ix[5] << 8) | (ix[6] << 16) | (ix[7] << 24);
    e.z = ix[8] | (ix[9] << 8) | ((unsigned)fr.nq << 16) | ((unsigned)(slot & 0xFF) << 24);
    e.w = (unsigned)fao | ((unsigned)ffo << 16);
    return e;
}

// chain step through one factor: nxt = (fragment matrix) cur.  fg = this lane's fragments of the factor (F for the forward, F^T
// for the backward direction): 2 qubits: fg[q * 32], q = mt * 4 + kk; 1 qubit: fg[0].  cur / nxt are addressed as wb[offset + index]
// with integer offsets: swapping POINTERS made the compiler fall back to generic-space loads and stores (LD.E / ST.E with 64-bit
// address arithmetic) instead of LDS / STS.
// The update is IN PLACE: a lane's stores depend on its DMMA results, and a (warp-collective) DMMA cannot complete before every lane's
// operand loads have returned -- no lane can overwrite an element another lane still has to read.
__device__ __forceinline__ void fj64_chain(const uint4 tb, const double* fg, double* wb, unsigned lg, unsigned lt)
{
    const double* cur = wb; double* nxt = wb;
    if (((tb.z >> 16) & 0xFFu) == 2u) {
        double2 o0 = make_double2(0.0, 0.0), o1 = make_double2(0.0, 0.0);
        const double b0 = cur[(tb.x >> 16) & 0xFFu], b1 = cur[tb.x >> 24], b2 = cur[tb.y & 0xFFu], b3 = cur[(tb.y >> 8) & 0xFFu];
        dmma884(o0.x, o0.y, fg[0 * 32], b0); dmma884(o1.x, o1.y, fg[4 * 32], b0);
        dmma884(o0.x, o0.y, fg[1 * 32], b1); dmma884(o1.x, o1.y, fg[5 * 32], b1);
        dmma884(o0.x, o0.y, fg[2 * 32], b2); dmma884(o1.x, o1.y, fg[6 * 32], b2);
        dmma884(o0.x, o0.y, fg[3 * 32], b3); dmma884(o1.x, o1.y, fg[7 * 32], b3);
        if (lt < 2u) {
            nxt[(tb.y >> 16) & 0xFFu] = o0.x; nxt[tb.y >> 24] = o0.y;
            nxt[tb.z & 0xFFu] = o1.x; nxt[(tb.z >> 8) & 0xFFu] = o1.y;
        }
    } else {
        const double A = fg[0];
        double2 o0 = make_double2(0.0, 0.0), o1 = make_double2(0.0, 0.0);
        dmma884(o0.x, o0.y, A, cur[tb.y & 0xFFu]);
        dmma884(o1.x, o1.y, A, cur[(tb.y >> 8) & 0xFFu]);
        if (lg < 4u) {
            nxt[(tb.y >> 16) & 0xFFu] = o0.x; nxt[tb.y >> 24] = o0.y;
            nxt[tb.z & 0xFFu] = o1.x; nxt[(tb.z >> 8) & 0xFFu] = o1.y;
        }
    }
}

// shared-memory prologue common to both kernels: fragment image (transposed or not), index table, op -> factor ranges
template <bool TRANSPOSED>
__device__ __forceinline__ void fj64_stage(const AtomDev& a, const FactoredDev& fd, const FjDev& fj, int n_fac, const int32_t* __restrict__ slots,
                                           double* frag, uint4* tab, int* fptr)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    for (int i = threadIdx.x; i <= a.n_ops; i += blockDim.x) fptr[i] = fd.op_fptr[i];
    for (int f = warp; f < n_fac; f += n_warps) {
        const FactorRec fr = fd.fac[f];
        const double* m = fd.mats + fr.moff;
        double* dst = frag + fj.ffo[f];
        if (fr.nq == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {                         // q = mt * 4 + kk: A[row = 8 mt + lg][col = 4 kk + lt]
                const int r = 8 * (q >> 2) + (int)lg, cc = 4 * (q & 3) + (int)lt;
                dst[q * 32 + lane] = TRANSPOSED ? m[cc * 16 + r] : m[r * 16 + cc];
            }
        } else {
            dst[lane] = TRANSPOSED ? m[(int)lt * 4 + (int)(lg & 3u)] : m[(int)(lg & 3u) * 4 + (int)lt];
        }
        tab[f * 32 + lane] = fj64_entry(fr, fj.fao[f], fj.ffo[f], slots ? slots[f] : 0xFF, lg, lt);
    }
    __syncthreads();
}

// Jacobian row of one outcome: J[p] = scale * sum over column p of the factor-space map  val * W_f[code]  -- accumulators (acc), e_0
// (cur, rho block of the circuit's prep) and s_L (sb, block of the outcome's effect).  One lane per parameter, FOUR parameters per lane
// in flight: the dependent loads cptr -> code -> element are L2 round trips when the CTA's shared memory leaves little L1 (ncu: the
// one-parameter-at-a-time loop was 25-37 % of the backward kernels' stall samples for 14-20 % of their instructions).
struct FjPeers { int n; double* J[B200_PEERS_MAX]; };     // fused exchange: the same Jacobian rows also go into the peers' arrays (NVLink)
template <bool PEERS>
__device__ __forceinline__ void fj64_row(const FjDev& fj, const double* acc, const double* cur, const double* sb, int prep, int eff,
                                         double* __restrict__ Jrow, double sc, int lane, const FjPeers& peers, int64_t roff)
{
    const int Np = fj.n_params;
    for (int p0 = lane; p0 < Np; p0 += 128) {
        int b[4], n[4]; double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + 32 * u;
            b[u] = p < Np ? __ldg(fj.cptr + p) : 0;
            n[u] = p < Np ? __ldg(fj.cptr + p + 1) - b[u] : 0;
            v[u] = 0.0;
        }
        const int nmax = max(max(n[0], n[1]), max(n[2], n[3]));
        for (int q = 0; q < nmax; ++q) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (q < n[u]) {
                    const uint32_t code = __ldg(fj.ccode + b[u] + q);
                    const double val = fj.unit ? ((code & 0x20000000u) ? -1.0 : 1.0) : __ldg(fj.cval + b[u] + q);
                    const uint32_t kind = code >> 30;
                    double x;
                    if (kind == 0u) x = acc[code & 0x1FFFFFFFu];
                    else {
                        const int i = (int)((code >> 16) & 0x1FFFu); const unsigned idx = fj_sw(code & 0xFFFFu);
                        x = (kind == 1u) ? (i == prep ? cur[idx] : 0.0) : (i == eff ? sb[idx] : 0.0);
                    }
                    v[u] = fma(val, x, v[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (p0 + 32 * u < Np) Jrow[p0 + 32 * u] = v[u] * sc;
        if (PEERS) for (int r = 0; r < peers.n; ++r) {         // 256-byte coalesced runs per warp instruction and destination
            double* Pr = peers.J[r] + roff;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (p0 + 32 * u < Np) Pr[p0 + 32 * u] = v[u] * sc;
        }
    }
}

// One warp per circuit.  Shared memory: fragment image [n_frag] | table [n_fac][32] uint4 | fptr [n_ops + 1] | per warp: 2 x 64 doubles.
__global__ void __launch_bounds__(256)
k_fj64_forward(AtomDev a, FactoredDev fd, FjDev fj, int n_fac, const double* __restrict__ rho, const double* __restrict__ E,
               double* __restrict__ FS, double* __restrict__ probs)
{
    constexpr int D = 64;
    extern __shared__ __align__(16) double smj[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    double* frag = smj;
    uint4* tab = reinterpret_cast<uint4*>(frag + ((fj.n_frag + 1) & ~1));
    int* fptr = reinterpret_cast<int*>(tab + (size_t)n_fac * 32);
    double* wbase = smj + (((fj.n_frag + 1) & ~1) + n_fac * 64 + ((a.n_ops + 4) >> 2) * 2) + warp * 2 * D;     // (plain offsets from smj: keeps the shared address space)
    fj64_stage<false>(a, fd, fj, n_fac, nullptr, frag, tab, fptr);
    const unsigned i0 = fj_sw((unsigned)lane), i1 = fj_sw((unsigned)lane + 32u);
    const int gw = blockIdx.x * n_warps + warp, nw = gridDim.x * n_warps;
    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c], L = a.circ_ptr[c + 1] - p0;
        const double* r = rho + (size_t)a.circ_prep[c] * D;
        double* row = FS + (size_t)fj.base[c] * D;
        { const double v0 = __ldg(r + lane), v1 = __ldg(r + lane + 32); wbase[i0] = v0; wbase[i1] = v1; row[lane] = v0; row[lane + 32] = v1; }
        __syncwarp();
        int gch = 0;
        for (uint32_t k = 0; k < L; ++k) {
            if ((k & 31u) == 0u) gch = (k + lane < L) ? __ldg(a.circ_ops + p0 + k + lane) : 0;      // 32 gate indices per load
            const int g = __shfl_sync(0xffffffffu, gch, (int)(k & 31u));
            const int f1 = fptr[g + 1];
            for (int f = fptr[g]; f < f1; ++f) {
                const uint4 tb = tab[f * 32 + lane];
                fj64_chain(tb, frag + (tb.w >> 16) + lane, wbase, lg, lt);
                __syncwarp();
                row += D;
                row[lane] = wbase[i0]; row[lane + 32] = wbase[i1];
                __syncwarp();                                      // (the next factor overwrites the state in place)
            }
        }
        if (probs) {
            for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
                const double* e = E + (size_t)a.out_eff[qo] * D;
                double part = fma(__ldg(e + lane), wbase[i0], __ldg(e + lane + 32) * wbase[i1]);
#pragma unroll
                for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
                if (lane == 0) probs[a.out_el[qo]] = part;
            }
        }
        __syncwarp();
    }
}

#define FJ64_ACC4(C, a0, a1, q0, q1) { dmma884(C[0].x, C[0].y, a0, q0); dmma884(C[1].x, C[1].y, a0, q1); \
                                       dmma884(C[2].x, C[2].y, a1, q0); dmma884(C[3].x, C[3].y, a1, q1); }

// One warp per (circuit, outcome).  Shared memory: F^T fragment image | table | fptr | per warp: e0, e1, s (64 doubles each), accumulators.
// slots[f]: register slot of a 2-qubit factor (0 .. FJ64_REG_SLOTS-1) or 0xFF; slot_fao[s]: accumulator offset of slot s or -1.
struct Fj64Slots { int fao[FJ64_REG_SLOTS]; };
template <bool PEERS>
__global__ void __launch_bounds__(256)
k_fj64_backward(AtomDev a, FactoredDev fd, FjDev fj, int n_fac, const int32_t* __restrict__ slots, Fj64Slots sl, const double* __restrict__ E,
                const double* __restrict__ FS, double* __restrict__ J, int64_t ld, const double* __restrict__ row_scale,
                unsigned* __restrict__ counter, int n_items, FjPeers peers)
{
    constexpr int D = 64;
    extern __shared__ __align__(16) double smj[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    double* frag = smj;
    uint4* tab = reinterpret_cast<uint4*>(frag + ((fj.n_frag + 1) & ~1));
    int* fptr = reinterpret_cast<int*>(tab + (size_t)n_fac * 32);
    double* wbase = smj + (((fj.n_frag + 1) & ~1) + n_fac * 64 + ((a.n_ops + 4) >> 2) * 2) + warp * (3 * D + fj.n_acc);
    double* sb = wbase + 2 * D; double* acc = wbase + 3 * D;       // wbase: e (updated in place) at offset 0
    fj64_stage<true>(a, fd, fj, n_fac, slots, frag, tab, fptr);
    const unsigned i0 = fj_sw((unsigned)lane), i1 = fj_sw((unsigned)lane + 32u);
    for (;;) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int c = fj.out_circ[item];
        const uint32_t row0 = fj.base[c];
        const uint32_t nst = fj.base[c + 1] - row0 - 1;            // factor steps of the circuit
        uint32_t t = nst;
        const int eff = a.out_eff[item];
        const int64_t el = a.out_el[item];
        const double* srow = FS + (size_t)row0 * D;
        for (int i = lane; i < fj.n_acc; i += 32) acc[i] = 0.0;
        wbase[i0] = __ldg(E + (size_t)eff * D + lane); wbase[i1] = __ldg(E + (size_t)eff * D + lane + 32);
        double2 R0[4], R1[4], R2[4], R3[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) R0[q] = R1[q] = R2[q] = R3[q] = make_double2(0.0, 0.0);
        double sr0 = 0.0, sr1 = 0.0;
        const double* sp = srow + (size_t)t * D + lane;           // this lane's element of the row AFTER the one requested next
        if (t > 0) { sp -= D; sr0 = __ldg(sp); sr1 = __ldg(sp + 32); }
        const uint16_t* sf = fj.step_fac + row0;
        // factor ids 32 at a time, the NEXT chunk requested when a chunk is entered; the table entry of a step is read during the step before
        int fch = 0, fnx = 0;
        uint4 tb = make_uint4(0u, 0u, 0u, 0u);
        if (t > 0) {
            const uint32_t ch = (t - 1u) >> 5;
            fch = (ch * 32u + lane < nst) ? (int)__ldg(sf + ch * 32u + lane) : 0;
            if (ch > 0u) fnx = (int)__ldg(sf + (ch - 1u) * 32u + lane);
            tb = tab[__shfl_sync(0xffffffffu, fch, (int)((t - 1u) & 31u)) * 32 + lane];
        }
        __syncwarp();
        {
            {
                while (t > 0) {
                --t;                                               // this step: factor f between s_t (before) and e (after)
                sb[i0] = sr0; sb[i1] = sr1;
                __syncwarp();
                uint4 tbn = tb;
                if (t > 0) {
                    sp -= D; sr0 = __ldg(sp); sr1 = __ldg(sp + 32);
                    const uint32_t tn = t - 1u;
                    if ((tn & 31u) == 31u) { fch = fnx; if ((tn >> 5) > 0u) fnx = (int)__ldg(sf + ((tn >> 5) - 1u) * 32u + lane); }
                    tbn = tab[__shfl_sync(0xffffffffu, fch, (int)(tn & 31u)) * 32 + lane];
                }
                const unsigned j0 = tb.x & 0xFFu, j1 = (tb.x >> 8) & 0xFFu;
                const double a0 = wbase[j0], a1 = wbase[j1], q0 = sb[j0], q1 = sb[j1];
                if (((tb.z >> 16) & 0xFFu) == 2u) {
                    switch (tb.z >> 24) {
                    case 0: FJ64_ACC4(R0, a0, a1, q0, q1); break;
                    case 1: FJ64_ACC4(R1, a0, a1, q0, q1); break;
                    case 2: FJ64_ACC4(R2, a0, a1, q0, q1); break;
                    case 3: FJ64_ACC4(R3, a0, a1, q0, q1); break;
                    default: {
                        double2* ap = reinterpret_cast<double2*>(acc + (tb.w & 0xFFFFu)) + lane;
                        double2 C[4] = {ap[0], ap[32], ap[64], ap[96]};
                        FJ64_ACC4(C, a0, a1, q0, q1);
                        ap[0] = C[0]; ap[32] = C[1]; ap[64] = C[2]; ap[96] = C[3];
                    } }
                } else {
                    // rows = (rest half h, a), K = 4 rest values of the half per DMMA: the two diagonal 4 x 4 blocks hold the sums
                    double2 cc = make_double2(0.0, 0.0);
                    dmma884(cc.x, cc.y, a0, q0); dmma884(cc.x, cc.y, a1, q1);
                    if ((lg >> 2) == (lt >> 1)) {
                        double2* ap = reinterpret_cast<double2*>(acc + (tb.w & 0xFFFFu) + (lg >> 2) * 16u + (lg & 3u) * 4u + 2u * (lt & 1u));
                        double2 v = *ap; v.x += cc.x; v.y += cc.y; *ap = v;
                    }
                }
                fj64_chain(tb, frag + (tb.w >> 16) + lane, wbase, lg, lt);
                tb = tbn;
                __syncwarp();
                }
            }
        }
        const double* cur = wbase;
        // ---- epilogue: register slots -> accumulator buffer; cur = e_0 (rho block); s_L (effect block) ----
        if (sl.fao[0] >= 0) { double2* ap = reinterpret_cast<double2*>(acc + sl.fao[0]) + lane; ap[0] = R0[0]; ap[32] = R0[1]; ap[64] = R0[2]; ap[96] = R0[3]; }
        if (sl.fao[1] >= 0) { double2* ap = reinterpret_cast<double2*>(acc + sl.fao[1]) + lane; ap[0] = R1[0]; ap[32] = R1[1]; ap[64] = R1[2]; ap[96] = R1[3]; }
        if (sl.fao[2] >= 0) { double2* ap = reinterpret_cast<double2*>(acc + sl.fao[2]) + lane; ap[0] = R2[0]; ap[32] = R2[1]; ap[64] = R2[2]; ap[96] = R2[3]; }
        if (sl.fao[3] >= 0) { double2* ap = reinterpret_cast<double2*>(acc + sl.fao[3]) + lane; ap[0] = R3[0]; ap[32] = R3[1]; ap[64] = R3[2]; ap[96] = R3[3]; }
        sb[i0] = __ldg(srow + (size_t)nst * D + lane); sb[i1] = __ldg(srow + (size_t)nst * D + lane + 32);
        __syncwarp();
        fj64_row<PEERS>(fj, acc, cur, sb, a.circ_prep[c], eff, J + el * ld, row_scale ? __ldg(row_scale + el) : 1.0, lane, peers, el * ld);
        __syncwarp();
    }
}

// =====================================================================================================================
// Probabilities from factor programs on the FP64 tensor cores (d = 64 and d = 256): k_probs_factored (kernels_factored.cuh) applies a
// factor with scalar FMAs whose matrix operands arrive as shared-memory broadcasts -- one LDS per FMA, the L1 pipe at 97 %.  Here a factor
// is the same DMMA chain step as in the Jacobian kernels: out(a', r) = sum_a F[a'][a] s(a, r), M = K = 4^nq, N = rest in tiles of 8, the
// state updated in place (a tile's outputs and the next tile's inputs have disjoint rest indices), all state indices of a step from a
// per-(factor, lane) table of 32 bytes built once per CTA.
// =====================================================================================================================
template <int D> struct FpCfg {
    static constexpr int R2 = D / 16, R1 = D / 4;
    static constexpr int NT2 = (R2 + 7) / 8, NT1 = R1 / 8;            // N tiles of a 2-qubit / 1-qubit factor
};
struct FpEntry { uint32_t w[8]; };    // (stored as two uint4 halves [f][half][lane]: conflict-free 16-byte loads)  bytes: 2 qubits: B[nt * 4 + kk] (NT2 * 4), then stores S[nt * 4 + 0..3]; 1 qubit: B[nt] (NT1), then S[nt * 2 + 0..1]

template <int D>
__device__ __forceinline__ unsigned fp_byte(const FpEntry& e, int i) { return (e.w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

template <int D>
__device__ __forceinline__ FpEntry fp_entry(const FactorRec fr, unsigned lg, unsigned lt)
{
    using C = FpCfg<D>;
    unsigned char b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = 0;
    if (fr.nq == 2) {
        const int s0 = fr.shift[0], s1 = fr.shift[1];
        const int lo = s0 < s1 ? s0 : s1, hi = s0 < s1 ? s1 : s0;
        for (int nt = 0; nt < C::NT2; ++nt) {
            const unsigned rB = (8u * nt + lg) & (unsigned)(C::R2 - 1);
            for (unsigned kk = 0; kk < 4; ++kk) b[nt * 4 + kk] = (unsigned char)fj_idx2(4u * kk + lt, rB, lo, hi, s0, s1);
            const unsigned rO = (8u * nt + 2u * lt) & (unsigned)(C::R2 - 1);
            unsigned char* st = b + C::NT2 * 4 + nt * 4;
            st[0] = (unsigned char)fj_idx2(lg, rO, lo, hi, s0, s1); st[1] = (unsigned char)fj_idx2(lg, rO + 1u, lo, hi, s0, s1);
            st[2] = (unsigned char)fj_idx2(lg + 8u, rO, lo, hi, s0, s1); st[3] = (unsigned char)fj_idx2(lg + 8u, rO + 1u, lo, hi, s0, s1);
        }
    } else {
        const int sh = fr.shift[0];
        for (int nt = 0; nt < C::NT1; ++nt) {
            b[nt] = (unsigned char)fj_idx1(lt, 8u * nt + lg, sh);
            b[C::NT1 + nt * 2] = (unsigned char)fj_idx1(lg & 3u, 8u * nt + 2u * lt, sh);
            b[C::NT1 + nt * 2 + 1] = (unsigned char)fj_idx1(lg & 3u, 8u * nt + 2u * lt + 1u, sh);
        }
    }
    FpEntry e;
#pragma unroll
    for (int i = 0; i < 8; ++i) e.w[i] = (unsigned)b[4 * i] | ((unsigned)b[4 * i + 1] << 8) | ((unsigned)b[4 * i + 2] << 16) | ((unsigned)b[4 * i + 3] << 24);
    return e;
}

// Shared memory: F fragment image [n_frag] | table [n_fac][32] FpEntry | fptr [n_ops + 1] | nq, fragment offset per factor [2 n_fac] | per warp: D doubles
template <int D>
__global__ void __launch_bounds__(256)
k_probs_fdmma(AtomDev a, FactoredDev fd, const int32_t* __restrict__ ffo_g, int n_frag, int n_fac, const double* __restrict__ rho,
              const double* __restrict__ E, double* __restrict__ out, int64_t el_stride)
{
    using C = FpCfg<D>;
    extern __shared__ __align__(16) double smj[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    double* frag = smj;
    uint4* tab = reinterpret_cast<uint4*>(frag + ((n_frag + 1) & ~1));
    int* fptr = reinterpret_cast<int*>(tab + (size_t)n_fac * 64);
    int* fmeta = fptr + (a.n_ops + 1);                                 // [f] = nq | fragment offset << 2
    double* st = smj + (((n_frag + 1) & ~1) + n_fac * 128 + ((a.n_ops + 1 + n_fac + 3) >> 2) * 2) + warp * D;
    for (int i = threadIdx.x; i <= a.n_ops; i += blockDim.x) fptr[i] = fd.op_fptr[i];
    for (int f = warp; f < n_fac; f += n_warps) {
        const FactorRec fr = fd.fac[f];
        const double* m = fd.mats + fr.moff;
        double* dst = frag + ffo_g[f];
        if (fr.nq == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q * 32 + lane] = m[(8 * (q >> 2) + (int)lg) * 16 + 4 * (q & 3) + (int)lt];     // A[row = 8 mt + lg][col = 4 kk + lt]
        } else {
            dst[lane] = m[(int)(lg & 3u) * 4 + (int)lt];
        }
        const FpEntry en = fp_entry<D>(fr, lg, lt);
        tab[(f * 2) * 32 + lane] = make_uint4(en.w[0], en.w[1], en.w[2], en.w[3]);
        tab[(f * 2 + 1) * 32 + lane] = make_uint4(en.w[4], en.w[5], en.w[6], en.w[7]);
        if (lane == 0) fmeta[f] = fr.nq | (ffo_g[f] << 2);
    }
    __syncthreads();
    const int gw = blockIdx.x * n_warps + warp, nw = gridDim.x * n_warps;
    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c], L = a.circ_ptr[c + 1] - p0;
        const double* r = rho + (size_t)a.circ_prep[c] * D;
#pragma unroll
        for (int j = 0; j < D / 32; ++j) st[fj_sw((unsigned)(lane + 32 * j))] = __ldg(r + lane + 32 * j);
        __syncwarp();
        int gch = 0;
        for (uint32_t k = 0; k < L; ++k) {
            if ((k & 31u) == 0u) gch = (k + lane < L) ? __ldg(a.circ_ops + p0 + k + lane) : 0;
            const int g = __shfl_sync(0xffffffffu, gch, (int)(k & 31u));
            const int f1 = fptr[g + 1];
            for (int f = fptr[g]; f < f1; ++f) {
                FpEntry tb;
                { const uint4 h0 = tab[(f * 2) * 32 + lane], h1 = tab[(f * 2 + 1) * 32 + lane];
                  tb.w[0] = h0.x; tb.w[1] = h0.y; tb.w[2] = h0.z; tb.w[3] = h0.w; tb.w[4] = h1.x; tb.w[5] = h1.y; tb.w[6] = h1.z; tb.w[7] = h1.w; }
                const int meta = fmeta[f];
                const double* fg = frag + (meta >> 2) + lane;
                if ((meta & 3) == 2) {
                    double A0[4], A1[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { A0[kk] = fg[kk * 32]; A1[kk] = fg[(4 + kk) * 32]; }
#pragma unroll
                    for (int nt = 0; nt < C::NT2; ++nt) {
                        double2 o0 = make_double2(0.0, 0.0), o1 = make_double2(0.0, 0.0);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const double b = st[fp_byte<D>(tb, nt * 4 + kk)];
                            dmma884(o0.x, o0.y, A0[kk], b); dmma884(o1.x, o1.y, A1[kk], b);
                        }
                        if (C::R2 >= 8 || lt < 2u) {
                            const int s = C::NT2 * 4 + nt * 4;
                            st[fp_byte<D>(tb, s)] = o0.x; st[fp_byte<D>(tb, s + 1)] = o0.y;
                            st[fp_byte<D>(tb, s + 2)] = o1.x; st[fp_byte<D>(tb, s + 3)] = o1.y;
                        }
                    }
                } else {
                    const double A = fg[0];
#pragma unroll
                    for (int nt = 0; nt < C::NT1; ++nt) {
                        double2 o = make_double2(0.0, 0.0);
                        dmma884(o.x, o.y, A, st[fp_byte<D>(tb, nt)]);
                        if (lg < 4u) { st[fp_byte<D>(tb, C::NT1 + nt * 2)] = o.x; st[fp_byte<D>(tb, C::NT1 + nt * 2 + 1)] = o.y; }
                    }
                }
                __syncwarp();
            }
        }
        for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
            const double* e = E + (size_t)a.out_eff[qo] * D;
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < D / 32; ++j) part = fma(__ldg(e + lane + 32 * j), st[fj_sw((unsigned)(lane + 32 * j))], part);
#pragma unroll
            for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
            if (lane == 0) out[(int64_t)a.out_el[qo] * el_stride] = part;
        }
        __syncwarp();
    }
}
