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
    const int32_t* fao;        // [n_fac] offset (doubles) of factor f's accumulators in a warp's accumulator buffer
    const int32_t* ffo;        // [n_fac] offset (doubles) of factor f's F^T fragments in the CTA's fragment image
    int n_acc, n_frag;
    const int32_t* cptr;       // [n_params + 1]  CSC of the factor-space derivative map
    const uint32_t* ccode;     // [nnz]  kind << 30 | ...: 0 = accumulator offset; 1 = rho (prep << 16 | index); 2 = effect (effect << 16 | index)
    const double* cval;        // [nnz]
    int n_params;
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

// swizzle of the backward kernel's shared-memory vectors: the fragment loads walk the state with lanes differing in one or two
// base-4 digits; folding digits 2 and 3 into digit 1 keeps the 16 lanes of a half-warp on 16 different 8-byte banks for every
// target-qubit combination of d = 64 and d = 256 (fac_sw folds into digit 0, which collides with the K index of the fragments)
__device__ __forceinline__ unsigned fj_sw(unsigned i) { return i ^ ((((i >> 4) ^ (i >> 6)) & 3u) << 2); }
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
    double* wbase = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ffo + n_fac) + 15) & ~(uintptr_t)15) + (size_t)warp * (3 * D + fj.n_acc);
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
                if (kind == 0u) x = acc[code];
                else {
                    const int i = (int)((code >> 16) & 0x3FFFu); const unsigned idx = fj_sw(code & 0xFFFFu);
                    x = (kind == 1u) ? (i == prep ? cur[idx] : 0.0) : (i == eff ? sb[idx] : 0.0);
                }
                v = fma(val, x, v);
            }
            Jrow[p] = v * sc;
        }
        __syncwarp();
    }
}
