// kernels_factored.cuh -- gates that are PRODUCTS OF SMALL OPERATIONS EMBEDDED ON 1-2 QUBITS, applied without densifying.
//
// For dim > 64 the reference does not hold dense d x d gates: a layer operation of a local-noise / cloud-crosstalk model is an
// OpCRep_Composed of OpCRep_Embedded factors (pygsti/evotypes/densitymx/opcreps.cpp:242-276, 93-158), each a 4 x 4 or 16 x 16
// superoperator acting on the base-4 digits (one Pauli-basis digit per qubit) of the state index that belong to its target
// qubits.  Round 1 densified every layer label on the HOST (`to_dense`: 0.01-0.2 s per label at d = 256) and then ran dense
// 256 x 256 products.  Here the factor program itself is the device representation:
//   k_probs_factored<D>   one warp per circuit, the state (D doubles) ping-pongs between two shared-memory buffers; a factor
//                         costs 4 (1 qubit) or 16 (2 qubits) multiply-adds per state component instead of D -- 16-64x fewer
//                         flops than the dense product, and no 512 KB gate matrix is ever read.  p = E . s at the end.
//   k_factored_to_dense<D> builds the dense G / G^T the derivative paths consume by pushing the D basis vectors through the
//                         same programs (one warp per (gate, column)): the host never forms a d x d matrix.
// Index convention: state index i = sum_q digit_q 4^(n-1-q) (qubit 0 most significant, the Kronecker order of pyGSTi's
// Pauli-product basis); a factor with target qubits (q_0, q_1) uses small-matrix index t = 4 digit_{q_0} + digit_{q_1}.
#pragma once
#include "common.cuh"

struct FactorRec {
    int32_t nq;            // target qubits: 1 or 2
    int32_t shift[2];      // bit position of each target digit in the state index: 2 (n_qubits - 1 - q)
    int32_t moff;          // offset (doubles) of the (4^nq x 4^nq) row-major matrix in `mats`
};
struct FactoredDev {
    const int32_t* op_fptr;    // [n_ops + 1]
    const FactorRec* fac;
    const double* mats;
};

// s_out = F s_in for one embedded factor (warp-cooperative; s_in / s_out in shared memory, D doubles each).
// The D state components fall into D / 4^nq groups that share their non-target digits; a group's 4^nq inputs are loaded once
// into registers and all its outputs are formed from them, the matrix entries arriving as shared-memory BROADCASTS (all lanes of
// an instruction read the same word).  1 qubit: one lane per group (4 inputs -> 4 outputs, 16 FMA).  2 qubits: two lanes per
// group (16 inputs each, 8 outputs each, 128 FMA).
// The state is stored SWIZZLED, component i at fac_sw(i) = i ^ ((i >> 4) & 15): a group's members differ in one or two base-4
// digits, i.e. the lanes of a load walk the state with strides 4, 16 or 64 -- all multiples of the 16 eight-byte banks' period
// without the swizzle (ncu on the first version: 129 M of 246 M shared wavefronts were bank conflicts, L1 data pipe 98 % busy).
// History: one output per lane iteration with per-FMA index arithmetic and global matrix loads took 3.2 ms for BASELINE config
// 5's circuits (no faster than the dense tensor-core products); groups + broadcast matrices 1.05 ms.
__device__ __forceinline__ unsigned fac_insert2(unsigned v, int sh) {      // open a 2-bit hole at bit position sh
    return ((v >> sh) << (sh + 2)) | (v & ((1u << sh) - 1u));
}
__device__ __forceinline__ unsigned fac_sw(unsigned i) { return i ^ ((i >> 4) & 15u); }

template <int D>
__device__ __forceinline__ void apply_factor(const FactorRec f, const double* g, const double* s_in, double* s_out, int lane)
{
    if (f.nq == 1) {
        const int sh = f.shift[0];
#pragma unroll
        for (int r0 = 0; r0 < D / 4; r0 += 32) {
            const int r = r0 + lane;
            if (D / 4 < 32 && r >= D / 4) break;
            const unsigned rest = fac_insert2((unsigned)r, sh);
            double in[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) in[t] = s_in[fac_sw(rest | ((unsigned)t << sh))];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                double acc = g[t * 4] * in[0];
                acc = fma(g[t * 4 + 1], in[1], acc); acc = fma(g[t * 4 + 2], in[2], acc); acc = fma(g[t * 4 + 3], in[3], acc);
                s_out[fac_sw(rest | ((unsigned)t << sh))] = acc;
            }
        }
    } else {
        const int s0 = f.shift[0], s1 = f.shift[1];
        const int lo = s0 < s1 ? s0 : s1, hi = s0 < s1 ? s1 : s0;
        const int half = lane & 1;
#pragma unroll
        for (int r0 = 0; r0 < D / 16; r0 += 16) {
            const int r = r0 + (lane >> 1);
            if (D / 16 < 16 && r >= D / 16) break;
            const unsigned rest = fac_insert2(fac_insert2((unsigned)r, lo), hi);
            double in[16];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) in[a * 4 + b] = s_in[fac_sw(rest | ((unsigned)a << s0) | ((unsigned)b << s1))];
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const int t = half * 8 + tt;
                const double* gr = g + t * 16;
                double acc = gr[0] * in[0];
#pragma unroll
                for (int k = 1; k < 16; ++k) acc = fma(gr[k], in[k], acc);
                s_out[fac_sw(rest | ((unsigned)(t >> 2) << s0) | ((unsigned)(t & 3) << s1))] = acc;
            }
        }
    }
}

#define FAC_WARPS 8
#define FAC_MATS_MAX 6144      // doubles of factor matrices staged in shared memory (48 KB)
#define FAC_RECS_MAX 512       // factor records staged in shared memory
// One warp per circuit.  Shared memory: [FAC_WARPS][2][D] state buffers | n_mats factor matrices | n_fac FactorRec | n_ops + 1 ints
template <int D>
__global__ void __launch_bounds__(FAC_WARPS * 32)
k_probs_factored(AtomDev a, FactoredDev fd, int n_mats, int n_fac, const double* __restrict__ rho, const double* __restrict__ E,
                 double* __restrict__ out, int64_t el_stride)
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
        for (int i = lane; i < D; i += 32) cur[fac_sw((unsigned)i)] = r[i];
        __syncwarp();
        int gnext = L ? __ldg(a.circ_ops + p0) : 0;
        for (uint32_t k = 0; k < L; ++k) {
            const int g = gnext;
            if (k + 1 < L) gnext = __ldg(a.circ_ops + p0 + k + 1);          // next step's gate index requested a step ahead
            const int f0 = fptr[g], f1 = fptr[g + 1];
            for (int f = f0; f < f1; ++f) {
                const FactorRec fr = recs[f];
                apply_factor<D>(fr, ms + fr.moff, cur, nxt, lane);
                __syncwarp();
                double* x = cur; cur = nxt; nxt = x;
            }
        }
        for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
            const double* e = E + (size_t)a.out_eff[qo] * D;
            double part = 0.0;
            for (int i = lane; i < D; i += 32) part = fma(__ldg(e + i), cur[fac_sw((unsigned)i)], part);
#pragma unroll
            for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
            if (lane == 0) out[(int64_t)a.out_el[qo] * el_stride] = part;
        }
        __syncwarp();
    }
}

// dense G[g][i][j] and Gt[g][j][i]: column j of gate g = its factor program applied to the basis vector e_j
template <int D>
__global__ void __launch_bounds__(FAC_WARPS * 32)
k_factored_to_dense(int n_ops, FactoredDev fd, double* __restrict__ G, double* __restrict__ Gt)
{
    extern __shared__ __align__(16) double smf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* b0 = smf + (size_t)warp * 2 * D;
    double* b1 = b0 + D;
    const int64_t gw = (int64_t)blockIdx.x * FAC_WARPS + warp, nw = (int64_t)gridDim.x * FAC_WARPS;
    for (int64_t w = gw; w < (int64_t)n_ops * D; w += nw) {
        const int g = (int)(w / D), j = (int)(w - (int64_t)g * D);
        double* cur = b0; double* nxt = b1;
        for (int i = lane; i < D; i += 32) cur[fac_sw((unsigned)i)] = (i == j) ? 1.0 : 0.0;
        __syncwarp();
        const int f0 = __ldg(fd.op_fptr + g), f1 = __ldg(fd.op_fptr + g + 1);
        for (int f = f0; f < f1; ++f) {
            const FactorRec fr = fd.fac[f];
            apply_factor<D>(fr, fd.mats + fr.moff, cur, nxt, lane);
            __syncwarp();
            double* x = cur; cur = nxt; nxt = x;
        }
        double* Gg = G + (size_t)g * D * D; double* Gtg = Gt + (size_t)g * D * D;
        for (int i = lane; i < D; i += 32) { const double v = cur[fac_sw((unsigned)i)]; Gg[(size_t)i * D + j] = v; Gtg[(size_t)j * D + i] = v; }
        __syncwarp();
    }
}
