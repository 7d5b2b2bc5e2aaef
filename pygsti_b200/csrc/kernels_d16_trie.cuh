// kernels_d16_trie.cuh -- d = 16 Jacobian with prefix AND suffix sharing.
//
// The reference's Map simulator shares circuit PREFIXES through its prefix table / state cache
// (pygsti/layouts/prefixtable.py:26-101, mapforwardsim_calc_densitymx.pyx:224-283: 3 151 617 -> 189 293
// propagations on the BASELINE layout).  The analytic Jacobian additionally needs the backward vectors
// e_k = (G_{L-1} ... G_{k+1})^T E, which depend only on the circuit SUFFIX and the effect -- so they share
// exactly the same way through a trie of reversed circuits (151 802 nodes x 4 effects instead of 12.6 M
// chain steps).  Both tries are built once per layout atom on the host (engine.cu: build_trie).
//
//   phase A  k_trie_chains : each trie is cut into chains (maximal runs of nodes created together); warps pull
//            chains, ordered by start depth, from an atomic counter, wait for the parent node written by an
//            earlier chain (fence-free sentinel hand-off, see TRIE_SENT), then walk their chain: forward chains
//            s = G s (DFMA), backward chains E^T[8x16] <- E^T . G for all effects at once (DMMA, rows = effects).
//            Node values go to the tables S[n_fnodes][16], H[n_bnodes][n_eff][16].
//            Critical path = deepest circuit, ~340 k mat-vecs in total at BASELINE size: ~0.1 ms.
//   phase B  k_accum_trie_d16 : one warp per (circuit, outcome group, gate) unit:
//            W_g[i][j] = sum_{t: g_t = g} e_t[i] s_t[j]  as DMMA with K = 4 time steps; the rows are gathered from
//            the (~100 MB) tables through a host-built offset stream; a finished 16x16 block is a set of Jacobian
//            entries and is stored at once.  This kernel is the whole HBM-write-bound cost.
#pragma once
#include "common.cuh"
#include "kernels_d16.cuh"   // dmma884, D16Args, D16_SPAM_MAX

struct TrieDev {
    // forward trie (node value = state after the prefix; depth 0 = the prep itself)
    const int4* f_meta;        // [n_fchains] (parent node id or -(1+prep) for a root chain, first node id, length, 0):
                               //             the nodes of a chain are consecutive
    const uint8_t* f_op;       // [n_fnodes]  op applied to reach the node (255 = root copy)
    int n_fchains; uint32_t n_fnodes;
    // backward trie (node value = (suffix product)^T E_e for every effect e; depth 0 = E itself)
    const int4* b_meta; const uint8_t* b_op;
    int n_bchains; uint32_t n_bnodes;
    // per circuit step, in gate-bucket order (same order as TwoPhaseDev.bperm): node of s_k and node of e_k
    const uint32_t* fn_b; const uint32_t* bn_b;
    const uint32_t* f_end;     // [n_circ] node of s_L
    const uint32_t* b_end;     // [n_circ] node of e_0 (d p / d rho)
    const uint16_t* bcnt;      // [n_circ][n_ops] bucket sizes
    // value tables + synchronisation
    double* S;                 // [n_fnodes][16]
    double* H;                 // [n_bnodes][n_eff][16]
    unsigned* counters;        // [4] work counters: forward chains, backward chains, accumulate chunks (zeroed before launch)
    unsigned long long* prof;  // dev knob B200_CHAIN_PROF: [2 roles][4] warp-cycles in hand-out, parent wait, steps, total; [8..11] accumulate (or nullptr)
};

#define TRIE_WARPS 4
// Hand-off between chains without fences or flags: the tables are pre-filled with a NaN payload no computation can
// produce; a node value is complete when none of its words equals the sentinel (every 8-byte store is atomic, each
// word is written exactly once per call, readers poll through L2 with ld.cg).  A release/acquire flag per node was
// measured first: the MEMBAR of every release store put ~1.5 k cycles on each step of the critical path.
#define TRIE_SENT 0x7FF8DEADBEEF5EEDull
__device__ __forceinline__ bool is_sent(double v) { return (unsigned long long)__double_as_longlong(v) == TRIE_SENT; }

__global__ void k_fill_sentinel(double* __restrict__ p, size_t n) {
    const double s = __longlong_as_double((long long)TRIE_SENT);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = s;
}

// Work hand-out: a warp takes `kgrab` consecutive chains per atomicAdd (same-address L2 atomics serialise: ~136 k of them
// per Jacobian were a measurable part of the kernel) and reads one 16-byte record per chain; the ops of a chain are
// consecutive bytes and are fetched 32 at a time (one coalesced load, then a shuffle per step) so that no global load
// sits on the per-step critical path.  A warp processes the chains it holds in increasing index order and a chain only
// waits for a chain with a smaller index, handed out earlier: the spin-waits cannot deadlock.
// dynamic smem: n_ops*256 doubles (the CTA's role: backward B fragments or forward fragments)
//               + TRIE_WARPS*2*16 doubles (forward exchange)
__global__ void __launch_bounds__(TRIE_WARPS * 32)
k_trie_chains(AtomDev a, ModelDev m, TrieDev t, int kgrab, int fwd_only)
{
    extern __shared__ __align__(16) double smt[];
    double* frag = smt;
    double* fx_all = frag + a.n_ops * 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* G = m.M;
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;
    const int role = fwd_only ? 0 : (blockIdx.x & 1);      // 0: forward trie, 1: backward trie
    for (int idx = threadIdx.x; idx < a.n_ops * 256; idx += blockDim.x) {
        const int g = idx >> 8, r = (idx >> 5) & 7, l = idx & 31;
        if (role) {
            const int tt = r >> 1, u = r & 1, mr = l >> 2, q = l & 3;
            const int kk = (tt >> 1) * 8 + 2 * q + (tt & 1);
            frag[idx] = G[g * 256 + kk * 16 + 8 * u + mr];
        } else {
            frag[idx] = G[g * 256 + (l & 15) * 16 + (l >> 4) * 8 + r];
        }
    }
    __syncthreads();

    long long pr_grab = 0, pr_wait = 0, pr_step = 0; const long long pr_t0 = clock64();
    unsigned* ctr = t.counters + role;
    const int n_chains = role ? t.n_bchains : t.n_fchains;
    const int4* meta = role ? t.b_meta : t.f_meta;
    const uint8_t* ops = role ? t.b_op : t.f_op;

    if (role == 0) {
        const double* ffrag = frag;
        double* fx = fx_all + warp * 32;
        const int half = lane >> 4;
        for (;;) {
            const long long tq0 = clock64();
            int c0 = 0;
            if (lane == 0) c0 = (int)atomicAdd(ctr, (unsigned)kgrab);
            c0 = __shfl_sync(0xffffffffu, c0, 0);
            if (c0 >= n_chains) break;
            const int c1 = (c0 + kgrab < n_chains) ? c0 + kgrab : n_chains;
            int4 mt = __ldg(meta + c0);
            for (int ci = c0; ci < c1; ++ci) {
                const int parent = mt.x;
                const uint32_t first = (uint32_t)mt.y, len = (uint32_t)mt.z;
                if (ci + 1 < c1) mt = __ldg(meta + ci + 1);
                int opv = (lane < (int)len) ? (int)__ldg(ops + first + lane) : 0;
                const long long tq1 = clock64();
                if (ci == c0) pr_grab += tq1 - tq0;
                double v = 0.0;
                uint32_t i0 = 0;
                if (parent < 0) {                       // root chain: first node is the prep itself
                    if (lane < 16) { v = rho[(-1 - parent) * 16 + lane]; __stcg(t.S + (size_t)first * 16 + lane, v); }
                    i0 = 1;
                } else {
                    if (lane < 16) {
                        const double* pp = t.S + (size_t)parent * 16 + lane;
                        v = __ldcg(pp);
                        while (is_sent(v)) { __nanosleep(40); v = __ldcg(pp); }
                    }
                    __syncwarp();
                }
                const long long tq2 = clock64();
                pr_wait += tq2 - tq1;
                int cur = 0;
                if (lane < 16) fx[lane] = v;
                __syncwarp();
                for (uint32_t i = i0; i < len; ++i) {
                    if ((i & 31u) == 0u && i) opv = (i + lane < len) ? (int)__ldg(ops + first + i + lane) : 0;
                    const int g = __shfl_sync(0xffffffffu, opv, (int)(i & 31u));
                    const double2* s2 = reinterpret_cast<const double2*>(fx + cur * 16 + half * 8);
                    const double2 s0 = s2[0], s1 = s2[1], s2v = s2[2], s3 = s2[3];
                    const double* fp = ffrag + g * 256 + lane;
                    double f0 = fp[0] * s0.x, f1 = fp[32] * s0.y, f2 = fp[64] * s1.x, f3 = fp[96] * s1.y;
                    f0 = fma(fp[128], s2v.x, f0); f1 = fma(fp[160], s2v.y, f1);
                    f2 = fma(fp[192], s3.x, f2); f3 = fma(fp[224], s3.y, f3);
                    double w = (f0 + f1) + (f2 + f3);
                    w += shfl_xor_f64(w, 16);
                    cur ^= 1;
                    if (lane < 16) { fx[cur * 16 + lane] = w; __stcg(t.S + (size_t)(first + i) * 16 + lane, w); }
                    __syncwarp();
                }
                pr_step += clock64() - tq2;
            }
        }
    } else {
        const double* bfrag = frag;
        const int mrow = lane >> 2, q = lane & 3;
        const int ne = a.n_eff;
        const bool rowok = mrow < ne;
        for (;;) {
            const long long tq0 = clock64();
            int c0 = 0;
            if (lane == 0) c0 = (int)atomicAdd(ctr, (unsigned)kgrab);
            c0 = __shfl_sync(0xffffffffu, c0, 0);
            if (c0 >= n_chains) break;
            const int c1 = (c0 + kgrab < n_chains) ? c0 + kgrab : n_chains;
            int4 mt = __ldg(meta + c0);
            for (int ci = c0; ci < c1; ++ci) {
                const int parent = mt.x;
                const uint32_t first = (uint32_t)mt.y, len = (uint32_t)mt.z;
                if (ci + 1 < c1) mt = __ldg(meta + ci + 1);
                int opv = (lane < (int)len) ? (int)__ldg(ops + first + lane) : 0;
                const long long tq1 = clock64();
                if (ci == c0) pr_grab += tq1 - tq0;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                uint32_t i0 = 0;
                if (parent < 0) {                       // root: E itself
                    if (rowok) {
                        const double* Er = E + mrow * 16;
                        a0 = Er[2 * q]; a1 = Er[2 * q + 1]; a2 = Er[8 + 2 * q]; a3 = Er[9 + 2 * q];
                        double* hp = t.H + ((size_t)first * ne + mrow) * 16 + 2 * q;
                        __stcg(reinterpret_cast<double2*>(hp), make_double2(a0, a1));
                        __stcg(reinterpret_cast<double2*>(hp + 8), make_double2(a2, a3));
                    }
                    i0 = 1;
                } else {
                    if (rowok) {
                        const double* hp = t.H + ((size_t)parent * ne + mrow) * 16 + 2 * q;
                        double2 x = __ldcg(reinterpret_cast<const double2*>(hp));
                        double2 y = __ldcg(reinterpret_cast<const double2*>(hp + 8));
                        while (is_sent(x.x) || is_sent(x.y) || is_sent(y.x) || is_sent(y.y)) {
                            __nanosleep(40);
                            x = __ldcg(reinterpret_cast<const double2*>(hp)); y = __ldcg(reinterpret_cast<const double2*>(hp + 8));
                        }
                        a0 = x.x; a1 = x.y; a2 = y.x; a3 = y.y;
                    }
                    __syncwarp();
                }
                const long long tq2 = clock64();
                pr_wait += tq2 - tq1;
                for (uint32_t i = i0; i < len; ++i) {
                    if ((i & 31u) == 0u && i) opv = (i + lane < len) ? (int)__ldg(ops + first + i + lane) : 0;
                    const int g = __shfl_sync(0xffffffffu, opv, (int)(i & 31u));
                    const double* bp = bfrag + g * 256 + lane;
                    double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0, x00 = 0.0, x01 = 0.0, x10 = 0.0, x11 = 0.0;
                    dmma884(d00, d01, a0, bp[0]);   dmma884(d10, d11, a0, bp[32]);
                    dmma884(x00, x01, a1, bp[64]);  dmma884(x10, x11, a1, bp[96]);
                    dmma884(d00, d01, a2, bp[128]); dmma884(d10, d11, a2, bp[160]);
                    dmma884(x00, x01, a3, bp[192]); dmma884(x10, x11, a3, bp[224]);
                    a0 = d00 + x00; a1 = d01 + x01; a2 = d10 + x10; a3 = d11 + x11;
                    if (rowok) {
                        double* hp = t.H + ((size_t)(first + i) * ne + mrow) * 16 + 2 * q;
                        __stcg(reinterpret_cast<double2*>(hp), make_double2(a0, a1));
                        __stcg(reinterpret_cast<double2*>(hp + 8), make_double2(a2, a3));
                    }
                }
                pr_step += clock64() - tq2;
            }
        }
    }
    if (t.prof && lane == 0) {
        atomicAdd(t.prof + role * 4 + 0, (unsigned long long)pr_grab); atomicAdd(t.prof + role * 4 + 1, (unsigned long long)pr_wait);
        atomicAdd(t.prof + role * 4 + 2, (unsigned long long)pr_step); atomicAdd(t.prof + role * 4 + 3, (unsigned long long)(clock64() - pr_t0));
    }
}

// probabilities from the forward trie: p_el = E_e . s_L   (one half-warp per outcome)
__global__ void __launch_bounds__(256)
k_probs_trie_d16(AtomDev a, ModelDev m, const uint32_t* __restrict__ f_end, const double* __restrict__ S,
                 double* __restrict__ out, int64_t el_stride)
{
    const double* E = m.M + m.off_eff;
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;       // half-warp id = (circuit, outcome) slot
    const int64_t nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    for (int64_t c = hw; c < a.n_circ; c += nhw) {
        const double sv = S[(size_t)f_end[c] * 16 + sub];
        for (int q = a.out_ptr[c]; q < a.out_ptr[c + 1]; ++q) {
            double pr = E[a.out_eff[q] * 16 + sub] * sv;
#pragma unroll
            for (int mk = 8; mk > 0; mk >>= 1) pr += shfl_xor_f64(pr, mk);
            if (sub == 0) out[(int64_t)a.out_el[q] * el_stride] = pr;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase B: one warp per unit = (circuit, outcome group of <= 4 consecutive effects, gate); the group's outcomes are
// stacked along M:  W_g^{(o)}[i][j] = sum_t e_t^{(o)}[i] s_t[j]  ->  A = [e^{(0)}; ...; e^{(3)}] (64 x 4 per group of
// 4 time steps), B = s (4 x 16): 16 DMMA per 4 steps; the s rows are loaded once for the 4 outcomes and the 4 effect
// rows of a backward node are contiguous (512 B).
// All index work is done once on the host (engine.cu): `uidx` is ONE CONTIGUOUS STREAM, in unit order, of
// (S element offset, H element offset) pairs, every (unit) bucket padded to a multiple of 4 steps with the offsets of
// all-zero table rows.  Warps take chunks of consecutive units from an atomic counter and run a software pipeline
// over the stream that does not stop at unit boundaries (offsets two groups ahead, table rows one group ahead): the
// only exposed load latency is at the start of a chunk.  No predication in the loop: 1 + 2 + 8 eight-byte loads and
// 16 DMMA per 4 steps.  The unit of gate 0 also writes the SPAM / unmapped columns and the probabilities.
// dynamic smem: n_ops*4*32 int2 (column map fragments) + 2*SPAM_MAX ints
// ------------------------------------------------------------------------------------------------------------
#define AT_WARPS 8
#define AT_CHUNK 16

struct CGroup {            // 32 bytes: the outcomes of one circuit whose effect index lies in [e_base, e_base+4)
    int32_t el[4];         // element (Jacobian row) of the outcome with effect e_base + i, -1 = no such outcome
    uint32_t e_base;       // multiple of 4
    uint32_t prep;
    uint32_t f_end, b_end; // node of s_L, node of e_0
};
struct UnitRec {           // 32 bytes
    int32_t el[4];         // copy of the group's Jacobian rows
    uint32_t off;          // first stream entry of the unit
    uint32_t g_ng;         // gate | n_groups << 16
    uint32_t cgi;          // CGroup index (SPAM / probabilities, gate 0 only)
    uint32_t pad;
};

// 256-bit Jacobian stores (W256): the s-vector components are fed to the B fragments in the order
// jmap = {0,1,4,5,8,9,12,13 | 2,3,6,7,10,11,14,15}, so that a lane's four accumulators of one block row are the four
// CONSECUTIVE columns 4q..4q+3; one st.global.v4.f64 per lane then writes 8 full 128-byte lines per warp instruction
// (SASS STG.E.256) instead of two instructions that each touch half of 8 lines.
__device__ __forceinline__ void st256_cs(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int NO, bool W256>   // outcomes (consecutive effects) per unit: 4 (16 warps/SM) or 2 (24 warps/SM)
__global__ void __launch_bounds__(AT_WARPS * 32, (NO == 4 ? 2 : 3))
k_accum_trie_d16(AtomDev a, ModelDev m, TrieDev t, D16Args args, const UnitRec* __restrict__ units, int n_units,
                 const uint2* __restrict__ uidx, const CGroup* __restrict__ cgrp, unsigned* __restrict__ counter, int dbg)
{
    extern __shared__ __align__(16) unsigned char smb[];
    int2* cm_s = reinterpret_cast<int2*>(smb);                          // [n_ops*4][32]
    int* spamc_s = reinterpret_cast<int*>(cm_s + a.n_ops * 4 * 32);     // [SPAM_MAX]
    int* spamw_s = spamc_s + D16_SPAM_MAX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < a.n_ops * 4 * 32; idx += blockDim.x) {
        const int g = idx >> 7, tile = (idx >> 5) & 3, l = idx & 31;
        if (W256) {
            // tile = 2*h + z: z = 0 -> .x = first of the 4 consecutive columns of row 8h + mrow (or -1), .y unused
            const int i = 8 * (tile >> 1) + (l >> 2), j0 = 4 * (l & 3);
            const int* cp = args.colmap + g * 256 + i * 16 + j0;
            int2 cc = make_int2(-1, -1);
            if ((tile & 1) == 0 && cp[0] >= 0 && (cp[0] & 3) == 0 && (args.ld & 3) == 0 && cp[1] == cp[0] + 1 && cp[2] == cp[0] + 2 &&
                cp[3] == cp[0] + 3) cc.x = cp[0];
            cm_s[idx] = cc;
            continue;
        }
        const int i = 8 * (tile >> 1) + (l >> 2), jc = 8 * (tile & 1) + 2 * (l & 3);
        int2 cc = *reinterpret_cast<const int2*>(args.colmap + g * 256 + i * 16 + jc);
        if (cc.y == cc.x + 1 && cc.x >= 0 && ((cc.x | (int)(args.ld & 1)) & 1) == 0) cc.y = -2;
        cm_s[idx] = cc;
    }
    const int n_spam_s = args.n_spam < D16_SPAM_MAX ? args.n_spam : D16_SPAM_MAX;
    for (int tt = threadIdx.x; tt < n_spam_s; tt += blockDim.x) { spamc_s[tt] = args.spam_col[tt]; spamw_s[tt] = args.spam_w[tt]; }
    __syncthreads();

    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    auto ldk = [&](const double* p) -> double {
        double v; asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol_keep)); return v; };
    const unsigned mrow = lane >> 2, q = lane & 3;
    const unsigned ne16 = (unsigned)a.n_eff * 16u;
    const double* E = m.M + m.off_eff;
    const double* Sb = t.S + (W256 ? (4 * (mrow >> 1) + (mrow & 1)) : mrow);   // B-fragment column mrow <-> s component jmap0(mrow)
    constexpr int SB1 = W256 ? 2 : 8;                                         // second fragment: jmap1 = jmap0 + 2  |  mrow + 8
    const double* Hb = t.H + mrow;

    long long pr_pro = 0, pr_grp = 0, pr_epi = 0; const long long pr_t0 = clock64();   // dev knob B200_CHAIN_PROF
    for (;;) {
        const long long tq0 = clock64();
        int u0 = 0;
        if (lane == 0) u0 = (int)atomicAdd(counter, 1u) * AT_CHUNK;
        u0 = __shfl_sync(0xffffffffu, u0, 0);
        if (u0 >= n_units) break;
        const int u1 = (u0 + AT_CHUNK < n_units) ? u0 + AT_CHUNK : n_units;
        // ---- chunk prologue: the only exposed latency ----
        uint4 ra = __ldg(reinterpret_cast<const uint4*>(units + u0));        // el[4]
        uint4 rb = __ldg(reinterpret_cast<const uint4*>(units + u0) + 1);    // off, g_ng, cgi
        const uint2* ip = uidx + rb.x + q;
        uint2 nd1 = __ldg(ip + 4);
        double r[2 + 2 * NO];
        {
            const uint2 nd0 = __ldg(ip);
            const double* sp = Sb + nd0.x;
            const double* hp = Hb + nd0.y;
            r[0] = ldk(sp); r[1] = ldk(sp + SB1);
#pragma unroll
            for (int k = 0; k < 2 * NO; ++k) r[2 + k] = ldk(hp + 8 * k);
        }
        ip += 8;
        pr_pro += clock64() - tq0;
        for (int u = u0; u < u1; ++u) {
            const long long tq1 = clock64();
            // next unit's record (needed only after this unit's groups)
            const int un = (u + 1 < u1) ? u + 1 : u;
            const uint4 ran = __ldg(reinterpret_cast<const uint4*>(units + un));
            const uint4 rbn = __ldg(reinterpret_cast<const uint4*>(units + un) + 1);
            const int g = (int)(rb.y & 0xffffu), ngroups = (dbg == 1) ? 0 : (int)(rb.y >> 16);
            double acc[NO][8];
#pragma unroll
            for (int o = 0; o < NO; ++o)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[o][k] = 0.0;
#pragma unroll 1
            for (int gi = 0; gi < ngroups; ++gi) {
                const uint2 nd2 = __ldg(ip); ip += 4;
                const double* sp = Sb + nd1.x;
                const double* hp = Hb + nd1.y;
                double n[2 + 2 * NO];
                n[0] = ldk(sp); n[1] = ldk(sp + SB1);
#pragma unroll
                for (int k = 0; k < 2 * NO; ++k) n[2 + k] = ldk(hp + 8 * k);
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    dmma884(acc[o][0], acc[o][1], r[2 + 2 * o], r[0]); dmma884(acc[o][2], acc[o][3], r[2 + 2 * o], r[1]);
                    dmma884(acc[o][4], acc[o][5], r[3 + 2 * o], r[0]); dmma884(acc[o][6], acc[o][7], r[3 + 2 * o], r[1]);
                }
#pragma unroll
                for (int k = 0; k < 2 + 2 * NO; ++k) r[k] = n[k];
                nd1 = nd2;
            }
            const long long tq2 = clock64();
            pr_grp += tq2 - tq1;
            if (dbg == 2) { if (acc[0][0] + acc[NO - 1][1] == 1.2345e300) args.J[0] = 1.0; ra = ran; rb = rbn; continue; }
            // ---------------- epilogue: the finished 16x16 blocks are Jacobian entries ----------------
            const int2* cm = cm_s + g * 128 + lane;
            int2 cc[4];
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) cc[tile] = cm[tile * 32];
            const int els[4] = {(int)ra.x, (int)ra.y, (int)ra.z, (int)ra.w};
            const bool fast = W256 ? __all_sync(0xffffffffu, (cc[0].x >= 0) & (cc[2].x >= 0))
                                   : __all_sync(0xffffffffu, (cc[0].y == -2) & (cc[1].y == -2) & (cc[2].y == -2) & (cc[3].y == -2));
            if (args.row_scale) {          // objective-function row scaling fused into the epilogue
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const double sc = (els[o] >= 0) ? __ldg(args.row_scale + els[o]) : 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[o][k] *= sc;
                }
            }
            if (W256) {
                if (fast) {
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] >= 0) {
                            double* Jr = args.J + (int64_t)els[o] * args.ld;
                            st256_cs(Jr + cc[0].x, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                            st256_cs(Jr + cc[2].x, acc[o][4], acc[o][5], acc[o][6], acc[o][7]);
                        }
                    }
                } else {                       // arbitrary column map: scalar stores through the map itself
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] < 0) continue;
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int* cp = args.colmap + g * 256 + (8 * h + (int)mrow) * 16 + 4 * (int)q;
#pragma unroll
                            for (int k = 0; k < 4; ++k) { const int col = __ldg(cp + k); if (col >= 0) Jr[col] = acc[o][4 * h + k]; }
                        }
                    }
                }
            } else if (fast) {
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    if (els[o] >= 0) {
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int tile = 0; tile < 4; ++tile)
                            __stcs(reinterpret_cast<double2*>(Jr + cc[tile].x), make_double2(acc[o][tile * 2], acc[o][tile * 2 + 1]));
                    }
                }
            } else {
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    if (els[o] >= 0) {
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int tile = 0; tile < 4; ++tile) {
                            const double v0 = acc[o][tile * 2], v1 = acc[o][tile * 2 + 1];
                            if (cc[tile].y == -2) {
                                *reinterpret_cast<double2*>(Jr + cc[tile].x) = make_double2(v0, v1);
                            } else {
                                if (cc[tile].x >= 0) Jr[cc[tile].x] = v0;
                                if (cc[tile].y >= 0) Jr[cc[tile].y] = v1;
                            }
                        }
                    }
                }
            }
            if (g == 0) {
                // SPAM / unmapped columns and probabilities of the group's outcomes
                const uint4 cg1 = __ldg(reinterpret_cast<const uint4*>(cgrp + rb.z) + 1);   // e_base | prep | f_end | b_end
                const int prep = (int)cg1.y;
                const double* sL = t.S + (size_t)cg1.z * 16;
                for (int o = 0; o < NO; ++o) {
                    if (els[o] < 0) continue;
                    const int ei = (int)cg1.x + o;
                    double* Jr = args.J + (int64_t)els[o] * args.ld;
                    if (args.probs) {
                        double pr = (lane < 16) ? E[ei * 16 + lane] * sL[lane] : 0.0;
#pragma unroll
                        for (int mk = 8; mk > 0; mk >>= 1) pr += shfl_xor_f64(pr, mk);
                        if (lane == 0) args.probs[els[o]] = pr;
                    }
                    const int w_rho0 = (int)m.off_rho + prep * 16, w_eff0 = (int)m.off_eff + ei * 16;
                    const double* e0 = t.H + (size_t)cg1.w * ne16 + ei * 16;
                    const double sc = args.row_scale ? __ldg(args.row_scale + els[o]) : 1.0;
                    for (int tt = lane; tt < args.n_spam; tt += 32) {
                        const int w = tt < D16_SPAM_MAX ? spamw_s[tt] : args.spam_w[tt];
                        const int col = tt < D16_SPAM_MAX ? spamc_s[tt] : args.spam_col[tt];
                        double val = 0.0;
                        if (w >= w_rho0 && w < w_rho0 + 16) val = e0[w - w_rho0] * sc;
                        else if (w >= w_eff0 && w < w_eff0 + 16) val = sL[w - w_eff0] * sc;
                        Jr[col] = val;
                    }
                }
            }
            ra = ran; rb = rbn;
            pr_epi += clock64() - tq2;
        }
    }
    if (t.prof && lane == 0) {
        atomicAdd(t.prof + 8, (unsigned long long)pr_pro); atomicAdd(t.prof + 9, (unsigned long long)pr_grp);
        atomicAdd(t.prof + 10, (unsigned long long)pr_epi); atomicAdd(t.prof + 11, (unsigned long long)(clock64() - pr_t0));
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase B, version 7 (default): same units, same epilogue, but the table rows travel global -> shared with cp.async
// (LDGSTS, 16-byte chunks, L2 evict_last) into a per-warp ring of AT_RING groups, so that AT_RING-1 groups of gathers
// (2.5 KB each) are in flight per warp -- also while the warp is busy storing a finished block.  Version 6 kept one
// group in registers; its profile (B200_CHAIN_PROF) showed 57 % of the warp cycles in the gather loop at ~1.8 k cycles
// per group of 4 steps, i.e. one exposed L2 round trip per group: by Little's law ~20 KB in flight per SM was all it
// could sustain.  Ring slot layout (doubles): S rows at q*20 (+mrow, +8+mrow), H rows at 80 + q*(NO*16+4) + o*16 (+mrow,
// +8+mrow): both strides are 4 mod 16, which makes the 64-bit fragment loads conflict-free per half-warp.
// dynamic smem: AT_WARPS * AT_RING * slot doubles, then the column-map fragments and SPAM lists as in version 6.
// ------------------------------------------------------------------------------------------------------------
#define AT_RING 4
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, uint64_t pol) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NO>
__global__ void __launch_bounds__(AT_WARPS * 32, 2)
k_accum_trie_d16_cp(AtomDev a, ModelDev m, TrieDev t, D16Args args, const UnitRec* __restrict__ units, int n_units,
                    const uint2* __restrict__ uidx, const CGroup* __restrict__ cgrp, unsigned* __restrict__ counter, int dbg)
{
    constexpr bool W256 = false;
    constexpr int SS = 20, HS = NO * 16 + 4, SLOT = 4 * SS + 4 * HS;       // doubles
    extern __shared__ __align__(16) unsigned char smb[];
    double* ring_all = reinterpret_cast<double*>(smb);                   // [AT_WARPS][AT_RING][SLOT]
    int2* cm_s = reinterpret_cast<int2*>(ring_all + AT_WARPS * AT_RING * SLOT);   // [n_ops*4][32]
    int* spamc_s = reinterpret_cast<int*>(cm_s + a.n_ops * 4 * 32);     // [SPAM_MAX]
    int* spamw_s = spamc_s + D16_SPAM_MAX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ring = ring_all + warp * (AT_RING * SLOT);
    for (int idx = threadIdx.x; idx < a.n_ops * 4 * 32; idx += blockDim.x) {
        const int g = idx >> 7, tile = (idx >> 5) & 3, l = idx & 31;
        const int i = 8 * (tile >> 1) + (l >> 2), jc = 8 * (tile & 1) + 2 * (l & 3);
        int2 cc = *reinterpret_cast<const int2*>(args.colmap + g * 256 + i * 16 + jc);
        if (cc.y == cc.x + 1 && cc.x >= 0 && ((cc.x | (int)(args.ld & 1)) & 1) == 0) cc.y = -2;
        cm_s[idx] = cc;
    }
    const int n_spam_s = args.n_spam < D16_SPAM_MAX ? args.n_spam : D16_SPAM_MAX;
    for (int tt = threadIdx.x; tt < n_spam_s; tt += blockDim.x) { spamc_s[tt] = args.spam_col[tt]; spamw_s[tt] = args.spam_w[tt]; }
    __syncthreads();

    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    const unsigned mrow = lane >> 2, q = lane & 3;
    const unsigned ne16 = (unsigned)a.n_eff * 16u;
    const double* E = m.M + m.off_eff;
    // copy roles of this lane: S chunk (step lane>>3, 16-byte chunk lane&7); H chunks j = 0..NO-1 (step j' = (lane + 32 j) / (8 NO))
    const int s_step = lane >> 3, s_chunk = lane & 7;
    // issue the copies of one group (4 stream entries at ipg) into ring slot `slot`
    auto issue = [&](const uint2 nd, int slot) {                       // nd: stream entry q of the group (lane q holds entry q)
        double* sl = ring + slot * SLOT;
        const unsigned so = __shfl_sync(0xffffffffu, nd.x, s_step);
        cp_async16(sl + s_step * SS + 2 * s_chunk, t.S + so + 2 * s_chunk, pol_keep);
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            const int id = lane + 32 * j;                                // chunk id in [0, 4 * 8 NO)
            const int st = id / (8 * NO), ch = id - st * (8 * NO);
            const unsigned ho = __shfl_sync(0xffffffffu, nd.y, st);
            cp_async16(sl + 4 * SS + st * HS + 2 * ch, t.H + ho + 2 * ch, pol_keep);
        }
        cp_async_commit();
    };

    long long pr_pro = 0, pr_grp = 0, pr_epi = 0; const long long pr_t0 = clock64();   // dev knob B200_CHAIN_PROF
    for (;;) {
        const long long tq0 = clock64();
        int u0 = 0;
        if (lane == 0) u0 = (int)atomicAdd(counter, 1u) * AT_CHUNK;
        u0 = __shfl_sync(0xffffffffu, u0, 0);
        if (u0 >= n_units) break;
        const int u1 = (u0 + AT_CHUNK < n_units) ? u0 + AT_CHUNK : n_units;
        uint4 ra = __ldg(reinterpret_cast<const uint4*>(units + u0));        // el[4]
        uint4 rb = __ldg(reinterpret_cast<const uint4*>(units + u0) + 1);    // off, g_ng, cgi
        const uint2* ip = uidx + rb.x + q;                                    // this lane's entry of the next group to ISSUE
        __syncwarp();                                                         // all lanes are done with the ring of the last chunk
        uint2 ndq[AT_RING];
#pragma unroll
        for (int k = 0; k < AT_RING; ++k) ndq[k] = __ldg(ip + 4 * k);
#pragma unroll
        for (int k = 0; k < AT_RING - 1; ++k) issue(ndq[k], k);
        uint2 nd_next = ndq[AT_RING - 1];                                     // entries of the group issued in the first iteration
        ip += 4 * AT_RING;
        int gcur = 0;                                                         // ring slot of the next group to CONSUME
        pr_pro += clock64() - tq0;
        for (int u = u0; u < u1; ++u) {
            const long long tq1 = clock64();
            const int un = (u + 1 < u1) ? u + 1 : u;
            const uint4 ran = __ldg(reinterpret_cast<const uint4*>(units + un));
            const uint4 rbn = __ldg(reinterpret_cast<const uint4*>(units + un) + 1);
            const int g = (int)(rb.y & 0xffffu), ngroups = (dbg == 1) ? 0 : (int)(rb.y >> 16);
            double acc[NO][8];
#pragma unroll
            for (int o = 0; o < NO; ++o)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[o][k] = 0.0;
#pragma unroll 1
            for (int gi = 0; gi < ngroups; ++gi) {
                cp_async_wait<AT_RING - 2>();                                 // the oldest group in flight has landed (this lane's part)
                __syncwarp();                                                 // ... all lanes' parts; everyone left the slot refilled below
                issue(nd_next, (gcur + AT_RING - 1) & (AT_RING - 1));
                nd_next = __ldg(ip); ip += 4;                                 // (used one iteration later)
                const double* sl = ring + gcur * SLOT;
                const double b0 = sl[q * SS + mrow], b1 = sl[q * SS + 8 + mrow];
                const double* hp = sl + 4 * SS + q * HS + mrow;
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const double a0 = hp[o * 16], a1 = hp[o * 16 + 8];
                    dmma884(acc[o][0], acc[o][1], a0, b0); dmma884(acc[o][2], acc[o][3], a0, b1);
                    dmma884(acc[o][4], acc[o][5], a1, b0); dmma884(acc[o][6], acc[o][7], a1, b1);
                }
                gcur = (gcur + 1) & (AT_RING - 1);
            }
            const long long tq2 = clock64();
            pr_grp += tq2 - tq1;
            if (dbg == 2) { if (acc[0][0] + acc[NO - 1][1] == 1.2345e300) args.J[0] = 1.0; ra = ran; rb = rbn; continue; }
            // ---------------- epilogue: the finished 16x16 blocks are Jacobian entries ----------------
            const int2* cm = cm_s + g * 128 + lane;
            int2 cc[4];
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) cc[tile] = cm[tile * 32];
            const int els[4] = {(int)ra.x, (int)ra.y, (int)ra.z, (int)ra.w};
            const bool fast = W256 ? __all_sync(0xffffffffu, (cc[0].x >= 0) & (cc[2].x >= 0))
                                   : __all_sync(0xffffffffu, (cc[0].y == -2) & (cc[1].y == -2) & (cc[2].y == -2) & (cc[3].y == -2));
            if (args.row_scale) {          // objective-function row scaling fused into the epilogue
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const double sc = (els[o] >= 0) ? __ldg(args.row_scale + els[o]) : 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[o][k] *= sc;
                }
            }
            if (W256) {
                if (fast) {
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] >= 0) {
                            double* Jr = args.J + (int64_t)els[o] * args.ld;
                            st256_cs(Jr + cc[0].x, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                            st256_cs(Jr + cc[2].x, acc[o][4], acc[o][5], acc[o][6], acc[o][7]);
                        }
                    }
                } else {                       // arbitrary column map: scalar stores through the map itself
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] < 0) continue;
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int* cp = args.colmap + g * 256 + (8 * h + (int)mrow) * 16 + 4 * (int)q;
#pragma unroll
                            for (int k = 0; k < 4; ++k) { const int col = __ldg(cp + k); if (col >= 0) Jr[col] = acc[o][4 * h + k]; }
                        }
                    }
                }
            } else if (fast) {
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    if (els[o] >= 0) {
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int tile = 0; tile < 4; ++tile)
                            __stcs(reinterpret_cast<double2*>(Jr + cc[tile].x), make_double2(acc[o][tile * 2], acc[o][tile * 2 + 1]));
                    }
                }
            } else {
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    if (els[o] >= 0) {
                        double* Jr = args.J + (int64_t)els[o] * args.ld;
#pragma unroll
                        for (int tile = 0; tile < 4; ++tile) {
                            const double v0 = acc[o][tile * 2], v1 = acc[o][tile * 2 + 1];
                            if (cc[tile].y == -2) {
                                *reinterpret_cast<double2*>(Jr + cc[tile].x) = make_double2(v0, v1);
                            } else {
                                if (cc[tile].x >= 0) Jr[cc[tile].x] = v0;
                                if (cc[tile].y >= 0) Jr[cc[tile].y] = v1;
                            }
                        }
                    }
                }
            }
            if (g == 0) {
                // SPAM / unmapped columns and probabilities of the group's outcomes
                const uint4 cg1 = __ldg(reinterpret_cast<const uint4*>(cgrp + rb.z) + 1);   // e_base | prep | f_end | b_end
                const int prep = (int)cg1.y;
                const double* sL = t.S + (size_t)cg1.z * 16;
                for (int o = 0; o < NO; ++o) {
                    if (els[o] < 0) continue;
                    const int ei = (int)cg1.x + o;
                    double* Jr = args.J + (int64_t)els[o] * args.ld;
                    if (args.probs) {
                        double pr = (lane < 16) ? E[ei * 16 + lane] * sL[lane] : 0.0;
#pragma unroll
                        for (int mk = 8; mk > 0; mk >>= 1) pr += shfl_xor_f64(pr, mk);
                        if (lane == 0) args.probs[els[o]] = pr;
                    }
                    const int w_rho0 = (int)m.off_rho + prep * 16, w_eff0 = (int)m.off_eff + ei * 16;
                    const double* e0 = t.H + (size_t)cg1.w * ne16 + ei * 16;
                    const double sc = args.row_scale ? __ldg(args.row_scale + els[o]) : 1.0;
                    for (int tt = lane; tt < args.n_spam; tt += 32) {
                        const int w = tt < D16_SPAM_MAX ? spamw_s[tt] : args.spam_w[tt];
                        const int col = tt < D16_SPAM_MAX ? spamc_s[tt] : args.spam_col[tt];
                        double val = 0.0;
                        if (w >= w_rho0 && w < w_rho0 + 16) val = e0[w - w_rho0] * sc;
                        else if (w >= w_eff0 && w < w_eff0 + 16) val = sL[w - w_eff0] * sc;
                        Jr[col] = val;
                    }
                }
            }
            ra = ran; rb = rbn;
            pr_epi += clock64() - tq2;
        }
        cp_async_wait<0>();                                                   // drain the look-ahead copies before the ring is reused
    }
    if (t.prof && lane == 0) {
        atomicAdd(t.prof + 8, (unsigned long long)pr_pro); atomicAdd(t.prof + 9, (unsigned long long)pr_grp);
        atomicAdd(t.prof + 10, (unsigned long long)pr_epi); atomicAdd(t.prof + 11, (unsigned long long)(clock64() - pr_t0));
    }
}
