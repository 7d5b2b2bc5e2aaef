// kernels_d16.cuh -- the 2-qubit (d = 16) Jacobian kernel: the BASELINE.json headline path.
//
// FP64 tensor-core formulation (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no f64 kind).  Measured in round 1:
// a 16x16 mat-vec per outcome out of shared memory is operand-bandwidth bound (27 shared-memory wavefronts per
// 16 DFMA) and the per-outcome epilogue's dependent global loads cost as much as the chain itself; both are
// designed out here.
//
// Persistent CTAs of 6 warps, a software pipeline over the CTA's circuits driven by named barriers:
//   warp 0  producer : forward chain s_k = G_k s_{k-1} of circuit it+1 (DFMA; gate fragments from shared memory,
//                      prefetched one step ahead), states (row stride 20 doubles: conflict-free DMMA fragment
//                      loads), op bytes, per-circuit metadata, and a counting sort of every 16-step chunk by gate
//                      (perm / cnt) into the other half of a double-buffered shared tile.
//   warp 1  chain    : backward chains of the circuit's (up to) 4 outcomes at once.  E^T[8 x 16] (rows 0-3 =
//                      outcomes) is the DMMA A operand, G the B operand (fragments from shared memory, loaded one
//                      step ahead -- no branch on the gate): E_new^T = E^T . G, 8 DMMA per step.  With the K-index
//                      relabelling sigma(t,q) = {2q, 2q+1, 8+2q, 9+2q}[t] the D fragment of one step IS the A
//                      fragment of the next: no shuffles and no shared-memory round trip on the dependent chain.
//                      Every e_k goes to a per-outcome history buffer (chunks of 16 steps, double buffered, the
//                      chunk that contains step 0 also carries e_0).
//   warps 2-5 accum  : outcome w-2.  W_g[i][j] += sum_t e_t[i] s_t[j] as DMMA with K = 4 time steps of the same
//                      gate (steps visited gate by gate through perm/cnt => static accumulator indexing),
//                      accumulators in registers: NG gates x (2x2 tiles x 2 doubles) = NG*8 doubles per lane.
//   epilogue         : the accumulators are the Jacobian row of a fully parameterised gate
//                      (dG[i,j]/dtheta_p = delta_{p,(i,j)}); stored straight to J through a per-lane column map
//                      held in shared memory, as 16-byte stores.
//   barriers         : FULL/EMPTY[chunk parity] between chain and accumulate warps (continuous chunk numbering
//                      across circuits, so the chain runs into the next circuit while the accumulate warps are
//                      still in the epilogue); PROD/CONS[buffer] between the producer and the five consumers.
// The only HBM traffic that scales with the problem is the Jacobian store: 8*(n_params+1) bytes per circuit
// outcome -- the kernel is HBM-write bound by construction.
//
// Arithmetic restated from the reference: dense acton (opcreps.cpp:40-54), adjoint acton
// (opcreps.cpp:56-68), effect dot (effectcreps.cpp:39-45); derivative structure
// matrixforwardsim.py:1059-1139 with dprod = sum_k (suffix) dG_k (prefix) (:729-792).
#pragma once
#include "common.cuh"

#define D16_ACC_WARPS 4
#define D16_THREADS ((D16_ACC_WARPS + 2) * 32)
#define D16_C 16          // chunk length (steps)
#define D16_NGP 8         // max gates (cnt stride)
#define D16_HS 20         // padded row stride (doubles) of states / history rows: 20 = 4 mod 16 -> conflict-free fragments
#define D16_HROWS (D16_C + 1)   // history rows per (buffer, outcome): C slots + the e_0 row
#define D16_SPAM_MAX 512  // SPAM/unmapped column list entries staged in shared memory
#define D16_META 16       // ints of per-circuit metadata: L, prep, o0, o1, eff[4], el[4]
// named barriers
#define D16_BAR_FULL 1    // +chunk parity : chain -> accum      (32 arrive + 128 sync)
#define D16_BAR_EMPTY 3   // +chunk parity : accum -> chain      (128 arrive + 32 sync)
#define D16_BAR_PROD 5    // +buffer       : producer -> consumers (32 arrive + 160 sync)
#define D16_BAR_CONS 7    // +buffer       : consumers -> producer (160 arrive + 32 sync)
#define D16_N_CA ((D16_ACC_WARPS + 1) * 32)
#define D16_N_ALL D16_THREADS

struct D16Args {
    const int32_t* colmap;    // [n_w]  J column of W index w (gate part used by the register epilogue), -1 = none
    const int32_t* spam_col;  // [n_spam] columns NOT fed by a gate element ...
    const int32_t* spam_w;    // [n_spam] ... and the rho/effect W index feeding each (or -1 -> zero)
    int n_spam;
    double* J;                // [n_elements][ld]
    int64_t ld;
    double* probs;            // [n_elements] or nullptr
    const double* row_scale;  // [n_elements] or nullptr: J row el is multiplied by row_scale[el] (trie path: fused in the
                              // epilogue; other d = 16 kernels: applied by k_scale_rows afterwards)
};

__host__ __device__ inline int d16_lpad(int max_depth) { return (max_depth + 2 * D16_C) & ~(D16_C - 1); }
__host__ __device__ inline int d16_nchmax(int max_depth) { return (max_depth + D16_C - 1) / D16_C + 1; }
// shared memory (doubles first, then ints, then bytes):
//   states[2][(Lmax+1)*HS] | hist[2][4][HROWS*HS] | bfrag[NG*8*32] | ffrag[NG*8*32]
//   cm[NG*4*32] int2 | spam_col, spam_w [SPAM_MAX] | meta[2][META]
//   ops[2][lpad] perm[2][lpad] cnt[2][nchmax*NGP]
__host__ __device__ inline size_t d16_smem_bytes(int ng, int max_depth) {
    size_t b = (size_t)2 * (max_depth + 1) * D16_HS * 8;
    b += (size_t)2 * 4 * D16_HROWS * D16_HS * 8;
    b += (size_t)2 * ng * 8 * 32 * 8;
    b += (size_t)ng * 4 * 32 * 8;
    b += (size_t)D16_SPAM_MAX * 8;
    b += (size_t)2 * D16_META * 4;
    b += (size_t)4 * d16_lpad(max_depth);
    b += (size_t)2 * d16_nchmax(max_depth) * D16_NGP;
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NG>
__global__ void __launch_bounds__(D16_THREADS, 2)
k_dprobs_d16(AtomDev a, ModelDev m, D16Args args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* states = reinterpret_cast<double*>(smem_raw);         // [2][(max_depth+1)][HS]
    const int st_stride = (a.max_depth + 1) * D16_HS;
    double* hist = states + 2 * st_stride;                        // [2][4][HROWS][HS]
    double* bfrag = hist + 2 * 4 * D16_HROWS * D16_HS;            // [NG][8][32]  chain B fragments
    double* ffrag = bfrag + NG * 8 * 32;                          // [NG][8][32]  forward fragments
    int2* cm_s = reinterpret_cast<int2*>(ffrag + NG * 8 * 32);    // [NG*4][32]  (c0, c1) per lane per tile; c1 = -2: 16-byte store at c0
    int* spamc_s = reinterpret_cast<int*>(cm_s + NG * 4 * 32);    // [SPAM_MAX]
    int* spamw_s = spamc_s + D16_SPAM_MAX;                        // [SPAM_MAX]
    int* meta_s = spamw_s + D16_SPAM_MAX;                         // [2][META]
    const int lpad = d16_lpad(a.max_depth);
    const int nchmax = d16_nchmax(a.max_depth);
    unsigned char* ops_s = reinterpret_cast<unsigned char*>(meta_s + 2 * D16_META);  // [2][lpad]
    unsigned char* perm_s = ops_s + 2 * lpad;                     // [2][lpad]   chunk-local slot, gate-sorted
    unsigned char* cnt_s = perm_s + 2 * lpad;                     // [2][nchmax][NGP]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const double* G = m.M;
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;

    // ---- one-time per CTA: fragment tables, per-lane column map, SPAM column lists -------------------------
    for (int idx = threadIdx.x; idx < NG * 8 * 32; idx += blockDim.x) {
        const int g = idx >> 8, r = (idx >> 5) & 7, l = idx & 31;
        double bv = 0.0, fv = 0.0;
        if (g < a.n_ops) {
            // chain B fragment r = t*2+u of lane l=(mrow,q): G_g[sigma(t,q)][8u + mrow], sigma = {2q, 2q+1, 8+2q, 9+2q}
            const int t = r >> 1, u = r & 1, mr = l >> 2, q = l & 3;
            const int kk = (t >> 1) * 8 + 2 * q + (t & 1);
            bv = G[g * 256 + kk * 16 + 8 * u + mr];
            // forward fragment r of lane l=(lo,half): G_g[lo][half*8 + r]
            fv = G[g * 256 + (l & 15) * 16 + (l >> 4) * 8 + r];
        }
        bfrag[idx] = bv; ffrag[idx] = fv;
    }
    for (int idx = threadIdx.x; idx < NG * 4 * 32; idx += blockDim.x) {
        const int g = idx >> 7, tile = (idx >> 5) & 3, l = idx & 31;
        int2 cc = make_int2(-1, -1);
        if (g < a.n_ops) {
            const int i = 8 * (tile >> 1) + (l >> 2), jc = 8 * (tile & 1) + 2 * (l & 3);
            cc = *reinterpret_cast<const int2*>(args.colmap + g * 256 + i * 16 + jc);
            if (cc.y == cc.x + 1 && cc.x >= 0 && ((cc.x | (int)(args.ld & 1)) & 1) == 0) cc.y = -2;
        }
        cm_s[idx] = cc;
    }
    const int n_spam_s = args.n_spam < D16_SPAM_MAX ? args.n_spam : D16_SPAM_MAX;
    for (int t = threadIdx.x; t < n_spam_s; t += blockDim.x) { spamc_s[t] = args.spam_col[t]; spamw_s[t] = args.spam_w[t]; }
    __syncthreads();

    // =====================================================================================
    // producer (warp 0): forward chain + op bytes + metadata + per-chunk gate buckets of circuit c -> buffer b
    //   chunk j covers steps k in [lo_j, hi_j), hi_j = L - j*C, lo_j = max(0, hi_j - C)   (top aligned)
    // =====================================================================================
    if (warp == 0) {
        const int lo = lane & 15, half = lane >> 4;
        int it = 0;
        for (int c = blockIdx.x; c < a.n_circ; c += gridDim.x, ++it) {
            const int b = it & 1;
            if (it >= 2) bar_sync_n(D16_BAR_CONS + b, D16_N_ALL);      // consumers are done with circuit it-2
            const uint32_t p0 = a.circ_ptr[c];
            const int L = (int)(a.circ_ptr[c + 1] - p0);
            const int32_t* ops = a.circ_ops + p0;
            double* st = states + b * st_stride;
            unsigned char* os = ops_s + b * lpad;
            unsigned char* pm = perm_s + b * lpad;
            unsigned char* cn = cnt_s + b * nchmax * D16_NGP;
            for (int k = lane; k < L; k += 32) os[k] = (unsigned char)ops[k];
            const int prep_c = a.circ_prep[c];
            if (lane < 16) st[lane] = rho[prep_c * 16 + lane];
            {
                int* mt = meta_s + b * D16_META;
                const int o0 = a.out_ptr[c], o1 = a.out_ptr[c + 1];
                if (lane == 0) { mt[0] = L; mt[1] = prep_c; mt[2] = o0; mt[3] = o1; }
                if (lane < 4) {
                    const bool ok = o0 + lane < o1;
                    mt[4 + lane] = ok ? a.out_eff[o0 + lane] : 0;
                    mt[8 + lane] = ok ? a.out_el[o0 + lane] : 0;
                }
            }
            __syncwarp();
            // gate buckets, two chunks per pass (one per half warp)
            const int nch = (L + D16_C - 1) / D16_C;
            for (int jb = 0; jb < nch; jb += 2) {
                const int j = jb + half;
                const int hi = L - j * D16_C;
                const int lo_j = hi - D16_C > 0 ? hi - D16_C : 0;
                const bool valid = (j < nch) && (lo < hi - lo_j);
                const int g = valid ? os[lo_j + lo] : 255;
                int off = 0, rank = 0, mycnt = 0;
#pragma unroll
                for (int gv = 0; gv < NG; ++gv) {
                    const unsigned mball = __ballot_sync(0xffffffffu, g == gv);
                    const unsigned mh = (mball >> (16 * half)) & 0xffffu;
                    const int cgv = __popc(mh);
                    if (g == gv) rank = off + __popc(mh & ((1u << lo) - 1u));
                    if (lo == gv) mycnt = cgv;
                    off += cgv;
                }
                if (valid) pm[j * D16_C + rank] = (unsigned char)lo;
                if (j < nch && lo < D16_NGP) cn[j * D16_NGP + lo] = (unsigned char)mycnt;
            }
            // forward chain: lane (lo, half): partial of s_new[lo] over inputs half*8..half*8+7, pair-summed
            const double2* s2 = reinterpret_cast<const double2*>(st + half * 8);
            double* sw = st + D16_HS + lane;
            double f[8];
            {
                const double* fp = ffrag + ((L > 0) ? os[0] : 0) * 256 + lane;
#pragma unroll
                for (int r = 0; r < 8; ++r) f[r] = fp[r * 32];
            }
            for (int k = 0; k < L; ++k) {
                const double2 s0 = s2[0], s1 = s2[1], s2v = s2[2], s3 = s2[3];
                double a0 = f[0] * s0.x, a1 = f[1] * s0.y, a2 = f[2] * s1.x, a3 = f[3] * s1.y;
                a0 = fma(f[4], s2v.x, a0); a1 = fma(f[5], s2v.y, a1);
                a2 = fma(f[6], s3.x, a2); a3 = fma(f[7], s3.y, a3);
                // prefetch the next step's gate fragment while the sums settle (os is padded; value unused at k = L-1)
                const double* fp = ffrag + (os[k + 1] & 7) * 256 + lane;
#pragma unroll
                for (int r = 0; r < 8; ++r) f[r] = fp[r * 32];
                double v = (a0 + a1) + (a2 + a3);
                v += shfl_xor_f64(v, 16);
                if (lane < 16) *sw = v;
                __syncwarp();
                s2 += D16_HS / 2; sw += D16_HS;
            }
            __threadfence_block();
            bar_arrive_n(D16_BAR_PROD + b, D16_N_ALL);
        }
        // drain the consumers' last arrivals so every barrier phase is balanced at exit
        if (it >= 2) { bar_sync_n(D16_BAR_CONS + (it & 1), D16_N_ALL); bar_sync_n(D16_BAR_CONS + ((it + 1) & 1), D16_N_ALL); }
        else if (it == 1) bar_sync_n(D16_BAR_CONS + 0, D16_N_ALL);
        return;
    }

    // =====================================================================================
    // chain warp (warp 1)
    // =====================================================================================
    if (warp == 1) {
        const int mrow = lane >> 2, q = lane & 3;
        int it = 0, gch = 0;
        for (int c = blockIdx.x; c < a.n_circ; c += gridDim.x, ++it) {
            const int buf = it & 1;
            bar_sync_n(D16_BAR_PROD + buf, D16_N_ALL);
            const int* mt = meta_s + buf * D16_META;
            const int L = mt[0];
            const double* st = states + buf * st_stride;
            const unsigned char* os = ops_s + buf * lpad;
            const int nch = (L + D16_C - 1) / D16_C;
            const int nch1 = nch > 0 ? nch : 1;
            const int o0 = mt[2], o1 = mt[3];
            for (int og = o0; og < o1; og += 4) {
                const int nin = (o1 - og) < 4 ? (o1 - og) : 4;
                const bool rowok = mrow < nin;
                const int ei = rowok ? (og == o0 ? mt[4 + (mrow & 3)] : a.out_eff[og + mrow]) : 0;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     // E^T[mrow][2q], [2q+1], [8+2q], [9+2q]
                if (rowok) {
                    const double* Er = E + ei * 16;
                    a0 = Er[2 * q]; a1 = Er[2 * q + 1]; a2 = Er[8 + 2 * q]; a3 = Er[9 + 2 * q];
                }
                if (args.probs) {
                    const double* sL = st + L * D16_HS;
                    double pr = a0 * sL[2 * q] + a1 * sL[2 * q + 1] + a2 * sL[8 + 2 * q] + a3 * sL[9 + 2 * q];
                    pr += shfl_xor_f64(pr, 1);
                    pr += shfl_xor_f64(pr, 2);
                    if (rowok && q == 0) args.probs[og == o0 ? mt[8 + (mrow & 3)] : a.out_el[og + mrow]] = pr;
                }
                // B fragments of the first step
                double bf[8];
                {
                    const double* bp = bfrag + ((L > 0) ? os[L - 1] : 0) * 256 + lane;
#pragma unroll
                    for (int r = 0; r < 8; ++r) bf[r] = bp[r * 32];
                }
                for (int j = 0; j < nch1; ++j, ++gch) {
                    const int b2 = gch & 1;
                    if (gch >= 2) bar_sync_n(D16_BAR_EMPTY + b2, D16_N_CA);
                    const int hi = L - j * D16_C;
                    const int lo_c = hi - D16_C > 0 ? hi - D16_C : 0;
                    double* hb = hist + (b2 * 4 + (mrow & 3)) * (D16_HROWS * D16_HS) + 2 * q;
                    for (int k = hi - 1; k >= lo_c; --k) {
                        if (mrow < 4) {
                            double* hp = hb + (k - lo_c) * D16_HS;
                            *reinterpret_cast<double2*>(hp) = make_double2(a0, a1);
                            *reinterpret_cast<double2*>(hp + 8) = make_double2(a2, a3);
                        }
                        double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
                        dmma884(d00, d01, a0, bf[0]); dmma884(d10, d11, a0, bf[1]);
                        dmma884(d00, d01, a1, bf[2]); dmma884(d10, d11, a1, bf[3]);
                        dmma884(d00, d01, a2, bf[4]); dmma884(d10, d11, a2, bf[5]);
                        dmma884(d00, d01, a3, bf[6]); dmma884(d10, d11, a3, bf[7]);
                        // next step's fragments (k-1; harmless garbage index masked at k = 0)
                        const double* bp = bfrag + (os[k > 0 ? k - 1 : 0] & 7) * 256 + lane;
#pragma unroll
                        for (int r = 0; r < 8; ++r) bf[r] = bp[r * 32];
                        a0 = d00; a1 = d01; a2 = d10; a3 = d11;
                    }
                    if (lo_c == 0 && mrow < 4) {       // e_0 travels in the extra row of this chunk's buffer
                        double* hp = hb + D16_C * D16_HS;
                        *reinterpret_cast<double2*>(hp) = make_double2(a0, a1);
                        *reinterpret_cast<double2*>(hp + 8) = make_double2(a2, a3);
                    }
                    __threadfence_block();
                    bar_arrive_n(D16_BAR_FULL + b2, D16_N_CA);
                }
            }
            bar_arrive_n(D16_BAR_CONS + buf, D16_N_ALL);
        }
        // drain the accumulate warps' last EMPTY arrivals
        if (gch >= 2) { bar_sync_n(D16_BAR_EMPTY + (gch & 1), D16_N_CA); bar_sync_n(D16_BAR_EMPTY + ((gch + 1) & 1), D16_N_CA); }
        else if (gch == 1) bar_sync_n(D16_BAR_EMPTY + 0, D16_N_CA);
        return;
    }

    // =====================================================================================
    // accumulate warps (warps 2..5): outcome ow of each group
    // =====================================================================================
    {
        const int ow = warp - 2;
        const int mrow = lane >> 2, q = lane & 3;
        int it = 0, gch = 0;
        for (int c = blockIdx.x; c < a.n_circ; c += gridDim.x, ++it) {
            const int buf = it & 1;
            bar_sync_n(D16_BAR_PROD + buf, D16_N_ALL);
            const int* mt = meta_s + buf * D16_META;
            const int L = mt[0];
            const int prep = mt[1];
            const double* st = states + buf * st_stride;
            const unsigned char* pm = perm_s + buf * lpad;
            const unsigned char* cn = cnt_s + buf * nchmax * D16_NGP;
            const int nch = (L + D16_C - 1) / D16_C;
            const int nch1 = nch > 0 ? nch : 1;
            const int o0 = mt[2], o1 = mt[3];
            for (int og = o0; og < o1; og += 4) {
                const bool active = (og + ow) < o1;
                double acc[NG][8];      // [g][(mt*2+nt)*2 + e] = W_g[8mt + mrow][8nt + 2q + e]
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[g][r] = 0.0;

                for (int j = 0; j < nch1; ++j, ++gch) {
                    const int b2 = gch & 1;
                    bar_sync_n(D16_BAR_FULL + b2, D16_N_CA);
                    const double* hbase = hist + (b2 * 4 + ow) * (D16_HROWS * D16_HS);
                    if (active && nch > 0) {
                        const int hi = L - j * D16_C;
                        const int lo_a = hi - D16_C > 0 ? hi - D16_C : 0;
                        const double* hb = hbase + mrow;
                        const double* sb = st + lo_a * D16_HS + mrow;
                        const unsigned char* pa = pm + j * D16_C;
                        const unsigned char* ca = cn + j * D16_NGP;
                        int tb = 0;
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            const int cg = ca[g];
                            for (int t0 = 0; t0 < cg; t0 += 4) {
                                const bool v = (t0 + q) < cg;
                                const int s = v ? pa[tb + t0 + q] : 0;
                                const double* hp = hb + s * D16_HS;
                                const double* sp = sb + s * D16_HS;
                                const double ea = v ? hp[0] : 0.0, eb = v ? hp[8] : 0.0;
                                const double sa = v ? sp[0] : 0.0, sbv = v ? sp[8] : 0.0;
                                dmma884(acc[g][0], acc[g][1], ea, sa);
                                dmma884(acc[g][2], acc[g][3], ea, sbv);
                                dmma884(acc[g][4], acc[g][5], eb, sa);
                                dmma884(acc[g][6], acc[g][7], eb, sbv);
                            }
                            tb += cg;
                        }
                    }
                    if (j == nch1 - 1 && active) {
                        // ---------------- epilogue (before releasing the chunk that carries e_0) ----------------
                        const int qo = og + ow;
                        const int ei = (og == o0) ? mt[4 + ow] : a.out_eff[qo];
                        const int64_t el = (og == o0) ? mt[8 + ow] : a.out_el[qo];
                        double* Jr = args.J + el * args.ld;
                        const int2* cm = cm_s + lane;
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
#pragma unroll
                            for (int tile = 0; tile < 4; ++tile) {
                                const int2 cc = cm[(g * 4 + tile) * 32];
                                const double v0 = acc[g][tile * 2], v1 = acc[g][tile * 2 + 1];
                                if (cc.y == -2) {
                                    *reinterpret_cast<double2*>(Jr + cc.x) = make_double2(v0, v1);
                                } else {
                                    if (cc.x >= 0) Jr[cc.x] = v0;
                                    if (cc.y >= 0) Jr[cc.y] = v1;
                                }
                            }
                        }
                        // SPAM / unmapped columns: d p/d rho = e_0, d p/d E_j = s_L for the outcome's own effect, else 0
                        const int w_rho0 = (int)m.off_rho + prep * 16, w_eff0 = (int)m.off_eff + ei * 16;
                        const double* e0 = hbase + D16_C * D16_HS;
                        const double* sL = st + L * D16_HS;
                        for (int t = lane; t < args.n_spam; t += 32) {
                            const int w = t < D16_SPAM_MAX ? spamw_s[t] : args.spam_w[t];
                            const int col = t < D16_SPAM_MAX ? spamc_s[t] : args.spam_col[t];
                            double val = 0.0;
                            if (w >= w_rho0 && w < w_rho0 + 16) val = e0[w - w_rho0];
                            else if (w >= w_eff0 && w < w_eff0 + 16) val = sL[w - w_eff0];
                            Jr[col] = val;
                        }
                    }
                    bar_arrive_n(D16_BAR_EMPTY + b2, D16_N_CA);
                }
            }
            bar_arrive_n(D16_BAR_CONS + buf, D16_N_ALL);
        }
    }
}
