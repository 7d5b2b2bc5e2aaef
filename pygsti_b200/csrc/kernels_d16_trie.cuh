// kernels_d16_trie.cuh -- d = 16 Jacobian with prefix AND suffix sharing.
//
// The reference's Map simulator shares circuit PREFIXES through its prefix table / state cache
// (pygsti/layouts/prefixtable.py:26-101, mapforwardsim_calc_densitymx.pyx:224-283: 3 151 617 -> 189 293
// propagations on the BASELINE layout).  The analytic Jacobian additionally needs the backward vectors
// e_k = (G_{L-1} ... G_{k+1})^T E, which depend only on the circuit SUFFIX and the effect -- so they share
// exactly the same way through a trie of reversed circuits (151 802 nodes x 4 effects instead of 12.6 M
// chain steps).  Both tries are built once per layout atom on the host (trie_host.h: build_trie).
//
//   phase A  k_trie_prepare, k_trie_chains : each trie is cut into chains by a heavy-path decomposition (trie_host.h); warps
//            pull chains, ordered by start depth, one at a time from an atomic counter, wait for the parent node written by
//            an earlier chain (fence-free sentinel hand-off, see TRIE_SENT), then walk their chain: forward chains
//            s = G s (DFMA), backward chains E^T[8x16] <- E^T . G for all effects at once (DMMA, rows = effects).
//            Node values go to the tables S[n_fnodes][16], H[n_bnodes][n_eff][16].
//            ~340 k mat-vecs in total at BASELINE size: 0.074 ms.
//   phase B  k_accum_trie_d16 : one warp per (circuit, outcome group, gate) unit:
//            W_g[i][j] = sum_{t: g_t = g} e_t[i] s_t[j]  as DMMA with K = 4 time steps; the rows are gathered from
//            the (~100 MB) tables through a host-built offset stream; a finished 16x16 block is a set of Jacobian
//            entries and is stored at once.  This kernel is the whole HBM-write-bound cost (0.68 ms, 0.67 of HBM peak).
#pragma once
#include "common.cuh"

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
};

#define TRIE_WARPS 4
#ifndef TRIE_MIN_CTAS
#define TRIE_MIN_CTAS 8
#endif
// Hand-off between chains without fences or flags: the rows a chain can wait for are pre-filled with a NaN payload no
// computation can produce; a node value is complete when none of its words equals the sentinel (every 8-byte store is atomic, each
// word is written exactly once per call, readers poll through L2 with ld.cg).  A release/acquire flag per node was
// measured first: the MEMBAR of every release store put ~1.5 k cycles on each step of the critical path.
#define TRIE_SENT 0x7FF8DEADBEEF5EEDull
__device__ __forceinline__ bool is_sent(double v) { return (unsigned long long)__double_as_longlong(v) == TRIE_SENT; }

// Only the rows a chain waits for -- the parent nodes of the chain heads, 40.8 k of 189 k forward and 32.7 k of 152 k backward
// nodes on the BASELINE layout -- are ever polled, so only those are reset before the chains run (22 MB instead of 102 MB);
// the same launch zeroes the work counters.  One launch instead of a memset and two fills.
__global__ void k_trie_prepare(double* __restrict__ S, const uint32_t* __restrict__ f_par, uint32_t n_fpar,
                               double* __restrict__ H, const uint32_t* __restrict__ b_par, uint32_t n_bpar, uint32_t h_row /* n_eff*16 */,
                               unsigned* __restrict__ counters)
{
    const double s = __longlong_as_double((long long)TRIE_SENT);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (tid < 4) counters[tid] = 0u;
    const size_t nf = (size_t)n_fpar * 16;
    for (size_t i = tid; i < nf; i += nth) S[(size_t)f_par[i >> 4] * 16 + (i & 15)] = s;
    if (H) {
        const size_t nb = (size_t)n_bpar * h_row;
        for (size_t i = tid; i < nb; i += nth) { const size_t r = i / h_row; H[(size_t)b_par[r] * h_row + (i - r * h_row)] = s; }
    }
}

// Work hand-out: a warp takes ONE chain per atomicAdd and reads one 16-byte record for it; the ops of a chain are
// consecutive bytes and are fetched 32 at a time (one coalesced load, then a shuffle per step) so that no global load
// sits on the per-step critical path.  Chains are handed out by start depth and a chain only waits for a chain with a
// smaller index, handed out earlier: the spin-waits cannot deadlock.  The parent row and the ops of the chain are
// requested before the chain is walked, and the atomic for the next chain is issued before a SHORT chain (<= 4 nodes) is
// walked.  Claiming chains any earlier than that was measured to be much worse (a warp that holds several claimed
// chains makes every other warp wait for parents that sit in its queue): the hand-out stays just-in-time
// (profiles/README.md, round 1).  A prefetched parent row that is not complete yet (sentinel) falls back to polling.
// dynamic smem: n_ops*256 doubles (the CTA's role: backward B fragments or forward fragments)
//               + TRIE_WARPS*2*16 doubles (forward exchange)
__global__ void __launch_bounds__(TRIE_WARPS * 32, TRIE_MIN_CTAS)
k_trie_chains(AtomDev a, ModelDev m, TrieDev t, int fwd_only, unsigned sleep_ns)
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

    unsigned* ctr = t.counters + role;
    const int n_chains = role ? t.n_bchains : t.n_fchains;
    const int4* meta = role ? t.b_meta : t.f_meta;
    const uint8_t* ops = role ? t.b_op : t.f_op;
    constexpr unsigned FULL = 0xffffffffu;
    auto grab = [&]() -> int { int c = n_chains; if (lane == 0) c = (int)atomicAdd(ctr, 1u); return c; };   // valid in lane 0
    int c_next = grab();

    if (role == 0) {
        const double* ffrag = frag;
        double* fx = fx_all + warp * 32;
        const int half = lane >> 4;
        for (;;) {
            const int c0 = __shfl_sync(FULL, c_next, 0);
            if (c0 >= n_chains) break;
            int4 mt = make_int4(0, 0, 0, 0);                // lane 0: record of chain c0
            if (lane == 0) mt = __ldg(meta + c0);
            const int parent = __shfl_sync(FULL, mt.x, 0);
            const uint32_t first = (uint32_t)__shfl_sync(FULL, mt.y, 0), len = (uint32_t)__shfl_sync(FULL, mt.z, 0);
            int opv = (lane < (int)len) ? (int)__ldg(ops + first + lane) : 0;      // parent row + first 32 ops
            double v = 0.0;
            if (len && lane < 16) v = (parent < 0) ? rho[(-1 - parent) * 16 + lane] : __ldcg(t.S + (size_t)parent * 16 + lane);
            bool grabbed = false;
            if (len <= 4u) { c_next = grab(); grabbed = true; }                    // (never while a long chain is walked)
            uint32_t i0 = 0;
            if (len) {
                if (parent < 0) {                           // root chain: first node is the prep itself
                    if (lane < 16) __stcg(t.S + (size_t)first * 16 + lane, v);
                    i0 = 1;
                } else {
                    if (lane < 16) {
                        const double* pp = t.S + (size_t)parent * 16 + lane;
                        while (is_sent(v)) { __nanosleep(sleep_ns); v = __ldcg(pp); }
                    }
                    __syncwarp();
                }
            }
            int cur = 0;
            if (lane < 16) fx[lane] = v;
            __syncwarp();
            for (uint32_t i = i0; i < len; ++i) {
                if ((i & 31u) == 0u && i) opv = (i + lane < len) ? (int)__ldg(ops + first + i + lane) : 0;
                const int g = __shfl_sync(FULL, opv, (int)(i & 31u));
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
            if (!grabbed) c_next = grab();
        }
    } else {
        const double* bfrag = frag;
        const int mrow = lane >> 2, q = lane & 3;
        const int ne = a.n_eff;
        const bool rowok = mrow < ne;
        for (;;) {
            const int c0 = __shfl_sync(FULL, c_next, 0);
            if (c0 >= n_chains) break;
            int4 mt = make_int4(0, 0, 0, 0);
            if (lane == 0) mt = __ldg(meta + c0);
            const int parent = __shfl_sync(FULL, mt.x, 0);
            const uint32_t first = (uint32_t)__shfl_sync(FULL, mt.y, 0), len = (uint32_t)__shfl_sync(FULL, mt.z, 0);
            int opv = (lane < (int)len) ? (int)__ldg(ops + first + lane) : 0;
            double2 x = make_double2(0.0, 0.0), y = x;
            if (len && rowok) {
                if (parent < 0) {                           // root: E itself
                    const double* Er = E + mrow * 16;
                    x = make_double2(Er[2 * q], Er[2 * q + 1]); y = make_double2(Er[8 + 2 * q], Er[9 + 2 * q]);
                } else {
                    const double* hp = t.H + ((size_t)parent * ne + mrow) * 16 + 2 * q;
                    x = __ldcg(reinterpret_cast<const double2*>(hp)); y = __ldcg(reinterpret_cast<const double2*>(hp + 8));
                }
            }
            bool grabbed = false;
            if (len <= 4u) { c_next = grab(); grabbed = true; }
            uint32_t i0 = 0;
            if (len) {
                if (parent < 0) {
                    if (rowok) {
                        double* hp = t.H + ((size_t)first * ne + mrow) * 16 + 2 * q;
                        __stcg(reinterpret_cast<double2*>(hp), x); __stcg(reinterpret_cast<double2*>(hp + 8), y);
                    }
                    i0 = 1;
                } else {
                    if (rowok) {
                        const double* hp = t.H + ((size_t)parent * ne + mrow) * 16 + 2 * q;
                        while (is_sent(x.x) || is_sent(x.y) || is_sent(y.x) || is_sent(y.y)) {
                            __nanosleep(sleep_ns);
                            x = __ldcg(reinterpret_cast<const double2*>(hp)); y = __ldcg(reinterpret_cast<const double2*>(hp + 8));
                        }
                    }
                    __syncwarp();
                }
            }
            double a0 = x.x, a1 = x.y, a2 = y.x, a3 = y.y;
            for (uint32_t i = i0; i < len; ++i) {
                if ((i & 31u) == 0u && i) opv = (i + lane < len) ? (int)__ldg(ops + first + i + lane) : 0;
                const int g = __shfl_sync(FULL, opv, (int)(i & 31u));
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
            if (!grabbed) c_next = grab();
        }
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
#define AT_CHUNK 8      // units per atomic grab (measured on C2: 2: 0.843, 4: 0.773, 8: 0.771, 12: 0.788, 16: 0.801, 24: 0.843 ms)

struct UnitRec {           // 32 bytes
    int32_t el[4];         // Jacobian rows (elements) of the outcomes with effect e_base + i, -1 = no such outcome
    uint32_t off;          // first stream entry of the unit
    uint32_t g_ng;         // gate (8 bits) | n_groups << 8 (14 bits) | e_base << 22 (3 bits) | prep << 25 (7 bits)
    uint32_t f_end, b_end; // node of s_L, node of e_0 (SPAM columns / probabilities: written by the gate-0 unit)
};
#define UNIT_MAX_GROUPS 16383u
#define UNIT_MAX_PREP 127u

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// Variants of this kernel that were built, measured on the BASELINE layout and REMOVED because they lost (round 1,
// profiles/README.md): interleaved fragments with 128-bit gathers / 256-bit stores (+3.5 %), a cp.async.bulk (TMA)
// store epilogue through conflict-free shared staging (+5 %), CTA-level unit ranges for L1 reuse of the H gathers (+1 %),
// 2-outcome units at 24 warps/SM (+2 %), a cp.async gather ring (+13 %), next-chunk look-ahead (+6 %).
// dynamic smem: AT_WARPS * 5 * 16 doubles (SPAM slots), then the column-map fragments and the SPAM lists.
// PEERS (b200_fill_dprobs_bcast_dev): every store of the epilogue is repeated, at the same offset, into the arrays of the
// peer GPUs (NVLink peer memory): the exchange of the sharded Jacobian overlaps its production instead of following it.
constexpr int AT_NO = 4;   // outcomes (consecutive effects) per unit
template <bool PEERS>
__global__ void __launch_bounds__(AT_WARPS * 32, 2)
k_accum_trie_d16(AtomDev a, ModelDev m, TrieDev t, D16Args args, const UnitRec* __restrict__ units, int n_units,
                 const uint2* __restrict__ uidx, unsigned* __restrict__ counter, int chunk)
{
    constexpr int NO = AT_NO;
    extern __shared__ __align__(128) unsigned char smb[];
    constexpr int SPS = (1 + NO) * 16;                                  // per-warp slot: s_L row + the e_0 rows of the unit's outcomes
    double* spam_stage = reinterpret_cast<double*>(smb);                // [AT_WARPS][SPS]
    int2* cm_s = reinterpret_cast<int2*>(spam_stage + AT_WARPS * SPS);  // [n_ops*4][32]
    int* spamc_s = reinterpret_cast<int*>(cm_s + a.n_ops * 4 * 32);     // [SPAM_MAX]
    int* spamw_s = spamc_s + D16_SPAM_MAX;
    // PEERS: per-warp staging of a unit's NO finished 16 x 16 blocks (2 KB each, in Jacobian-row order) for the bulk stores, and per
    // gate the first Jacobian column of its block when the block is one contiguous, 16-byte aligned run of 256 columns (else -1)
    double* bulk_stage = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(spamw_s + D16_SPAM_MAX) + 127) & ~(uintptr_t)127);   // [AT_WARPS][NO * 256]
    int* gbase_s = reinterpret_cast<int*>(bulk_stage + (PEERS ? AT_WARPS * NO * 256 : 0));                                             // [n_ops]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (PEERS) {
        for (int g = threadIdx.x; g < a.n_ops; g += blockDim.x) {
            const int b0 = args.colmap[g * 256];
            bool ok = b0 >= 0 && (b0 & 1) == 0 && (args.ld & 1) == 0;
            for (int e = 1; e < 256 && ok; ++e) ok = args.colmap[g * 256 + e] == b0 + e;
            gbase_s[g] = ok ? b0 : -1;
        }
    }
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
    auto next_chunk = [&]() -> int {                                    // first unit of this warp's next chunk (all lanes)
        unsigned sc = 0;
        if (lane == 0) sc = atomicAdd(counter, 1u);
        sc = __shfl_sync(0xffffffffu, sc, 0);
        const long long u = (long long)sc * chunk;
        return u < (long long)n_units ? (int)u : n_units;
    };

    // L2 policies of the gathers: S rows (24 MB, reused during the whole kernel) evict_last; H rows (78 MB, consumed front to back
    // in unit order) no hint -- evict_last on both was measured 4 % slower (0.858 vs 0.825 ms), evict_first on H 10 % slower
    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    auto ldk = [&](const double* p) -> double {
        double v; asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol_keep)); return v; };
    auto ldn = [](const double* p) -> double { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; };
    // one group of 4 steps: fr[0..1] = B elements (N tiles 0, 1), fr[2+2o], fr[3+2o] = A elements of outcome o (M tiles 0, 1)
    auto fetch = [&](const double* sp, const double* hp, double* fr) {
        fr[0] = ldk(sp); fr[1] = ldk(sp + 8);
#pragma unroll
        for (int k = 0; k < 2 * NO; ++k) fr[2 + k] = ldn(hp + 8 * k);
    };
    const unsigned mrow = lane >> 2, q = lane & 3;
    const unsigned ne16 = (unsigned)a.n_eff * 16u;
    const double* E = m.M + m.off_eff;
    const double* Sb = t.S + mrow;
    const double* Hb = t.H + mrow;

    for (;;) {
        const int u0 = next_chunk();
        if (u0 >= n_units) break;
        const int u1 = (u0 + chunk < n_units) ? u0 + chunk : n_units;
        // ---- chunk prologue: the only exposed latency ----
        uint4 ra = __ldg(reinterpret_cast<const uint4*>(units + u0));        // el[4]
        uint4 rb = __ldg(reinterpret_cast<const uint4*>(units + u0) + 1);    // off, g_ng, f_end, b_end
        // Index stream: every lane loads its own entry (q) of the group after next, used one iteration later.
        const uint2* ip = uidx + rb.x + q;
        uint2 nd1 = __ldg(ip + 4);
        double r[2 + 2 * NO];
        {
            const uint2 nd0 = __ldg(ip);
            fetch(Sb + nd0.x, Hb + nd0.y, r);
        }
        ip += 8;
        for (int u = u0; u < u1; ++u) {
            // next unit's record (needed only after this unit's groups)
            const int un = (u + 1 < u1) ? u + 1 : u;
            const uint4 ran = __ldg(reinterpret_cast<const uint4*>(units + un));
            const uint4 rbn = __ldg(reinterpret_cast<const uint4*>(units + un) + 1);
            const int g = (int)(rb.y & 0xffu), ngroups = (int)((rb.y >> 8) & 0x3fffu);
            if (g == 0) {
                // gate-0 unit: it also writes the SPAM columns / probabilities, from s_L and the e_0 rows.  They are copied to a
                // shared slot NOW (cp.async, no registers) so that their latency hides behind the group loop.
                __syncwarp();
                double* slot = spam_stage + warp * SPS;
                const double* e0 = t.H + (size_t)rb.w * ne16 + ((rb.y >> 22) & 7u) * 16u;      // NO consecutive effect rows
                if (lane < NO * 8) cp_async16(slot + 16 + 2 * lane, e0 + 2 * lane);
                if (lane < 8) cp_async16(slot + 2 * lane, t.S + (size_t)rb.z * 16 + 2 * lane);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
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
                fetch(sp, hp, n);
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    dmma884(acc[o][0], acc[o][1], r[2 + 2 * o], r[0]); dmma884(acc[o][2], acc[o][3], r[2 + 2 * o], r[1]);
                    dmma884(acc[o][4], acc[o][5], r[3 + 2 * o], r[0]); dmma884(acc[o][6], acc[o][7], r[3 + 2 * o], r[1]);
                }
#pragma unroll
                for (int k = 0; k < 2 + 2 * NO; ++k) r[k] = n[k];
                nd1 = nd2;
            }
            // ---------------- epilogue: the finished 16x16 blocks are Jacobian entries ----------------
            const int2* cm = cm_s + g * 128 + lane;
            int2 cc[4];
#pragma unroll
            for (int tile = 0; tile < 4; ++tile) cc[tile] = cm[tile * 32];
            const int els[4] = {(int)ra.x, (int)ra.y, (int)ra.z, (int)ra.w};
            const bool fast = __all_sync(0xffffffffu, (cc[0].y == -2) & (cc[1].y == -2) & (cc[2].y == -2) & (cc[3].y == -2));
            if (args.row_scale) {          // objective-function row scaling fused into the epilogue
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const double sc = (els[o] >= 0) ? __ldg(args.row_scale + els[o]) : 0.0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[o][k] *= sc;
                }
            }
            const int gb = PEERS ? gbase_s[g] : -1;
            if (PEERS && gb >= 0) {
                // Fused exchange, bulk form: the unit's blocks are staged in shared memory in Jacobian-row order and leave as 2 KB bulk
                // stores of the TMA engine, ONE instruction for all (outcome, destination) pairs (lane = 8 outcome + destination;
                // destination 0 = this GPU's array, 1.. = the peers' arrays over NVLink).  The per-lane form below sends every 16-byte
                // store instruction once per destination: 128 store instructions per unit at 7 peers, and 64-byte NVLink writes.
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous unit's staging has been read
                __syncwarp();
                double* stg = bulk_stage + warp * (NO * 256);
#pragma unroll
                for (int o = 0; o < NO; ++o)
#pragma unroll
                    for (int tile = 0; tile < 4; ++tile)
                        *reinterpret_cast<double2*>(stg + o * 256 + (8 * (tile >> 1) + (int)mrow) * 16 + 8 * (tile & 1) + 2 * (int)q) =
                            make_double2(acc[o][tile * 2], acc[o][tile * 2 + 1]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                const int bo = lane >> 3, bd = lane & 7;
                const int elo = bo == 0 ? els[0] : bo == 1 ? els[1] : bo == 2 ? els[2] : els[3];
                if (elo >= 0 && bd <= args.n_peers) {
                    double* Jb = args.J;
#pragma unroll
                    for (int r = 0; r < B200_PEERS_MAX; ++r) if (bd == r + 1) Jb = args.peerJ[r];
                    double* dstp = Jb + (int64_t)elo * args.ld + gb;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(stg + bo * 256);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp), "r"(src), "r"(2048u) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else {
            const int n_dst = PEERS ? 1 + args.n_peers : 1;
#pragma unroll 1
            for (int dst = 0; dst < n_dst; ++dst) {
                double* Jb = (!PEERS || dst == 0) ? args.J : args.peerJ[dst - 1];
                if (fast) {
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] >= 0) {
                            double* Jr = Jb + (int64_t)els[o] * args.ld;
#pragma unroll
                            for (int tile = 0; tile < 4; ++tile)
                                __stcs(reinterpret_cast<double2*>(Jr + cc[tile].x), make_double2(acc[o][tile * 2], acc[o][tile * 2 + 1]));
                        }
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        if (els[o] >= 0) {
                            double* Jr = Jb + (int64_t)els[o] * args.ld;
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
            }
            }
            if (g == 0) {
                // SPAM / unmapped columns and probabilities of the group's outcomes.  The rows they come from (s_L and e_0 of every
                // outcome) were requested at the top of the unit -- the unit record carries their node ids -- and are handed around
                // with shuffles.
                const int l16 = lane & 15;
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                const double* slot = spam_stage + warp * SPS;
                const double sLv = slot[l16];
                double e0v[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o) e0v[o] = slot[16 + o * 16 + l16];
                const int prep = (int)(rb.y >> 25), e_base = (int)((rb.y >> 22) & 7u);
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    if (els[o] < 0) continue;                                   // (warp-uniform)
                    const int ei = e_base + o;
                    const int64_t roff = (int64_t)els[o] * args.ld;
                    double* Jr = args.J + roff;
                    if (args.probs) {
                        double pr = (lane < 16) ? E[ei * 16 + lane] * sLv : 0.0;
#pragma unroll
                        for (int mk = 8; mk > 0; mk >>= 1) pr += shfl_xor_f64(pr, mk);
                        if (lane == 0) {
                            args.probs[els[o]] = pr;
                            if (PEERS) for (int r = 0; r < args.n_peers; ++r) if (args.peerP[r]) args.peerP[r][els[o]] = pr;
                        }
                    }
                    const int w_rho0 = (int)m.off_rho + prep * 16, w_eff0 = (int)m.off_eff + ei * 16;
                    const double sc = args.row_scale ? __ldg(args.row_scale + els[o]) : 1.0;
                    for (int t0 = 0; t0 < args.n_spam; t0 += 32) {
                        const int tt = t0 + lane;
                        const bool ok = tt < args.n_spam;
                        const int w = !ok ? -1 : (tt < D16_SPAM_MAX ? spamw_s[tt] : args.spam_w[tt]);
                        const int col = !ok ? 0 : (tt < D16_SPAM_MAX ? spamc_s[tt] : args.spam_col[tt]);
                        const int ir = w - w_rho0, ie = w - w_eff0;
                        const double vr = __shfl_sync(0xffffffffu, e0v[o], ir & 15), ve = __shfl_sync(0xffffffffu, sLv, ie & 15);
                        double val = 0.0;
                        if (ir >= 0 && ir < 16) val = vr * sc;
                        else if (ie >= 0 && ie < 16) val = ve * sc;
                        if (ok) {
                            Jr[col] = val;
                            if (PEERS) for (int r = 0; r < args.n_peers; ++r) args.peerJ[r][roff + col] = val;
                        }
                    }
                }
            }
            ra = ran; rb = rbn;
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (PEERS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // every bulk store of this thread has completed before the CTA leaves
}
