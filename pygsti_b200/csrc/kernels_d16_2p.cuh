// kernels_d16_2p.cuh -- two-phase d = 16 Jacobian: latency-bound chains and throughput-bound accumulation
// are separate kernels so each can run at the occupancy it needs.
//
// Why (measured, round 1): in the fused kernel the 1280 register accumulators per outcome cap the number of
// circuits in flight at 2 per SM, while one dependent chain step costs ~300 cycles (4 dependent DMMA at ~64
// cycles latency).  The chains need MANY circuits in flight and almost no registers; the accumulation needs the
// registers but no dependent chain.  Splitting them costs one round trip of the chain vectors through a scratch
// buffer (640 B per circuit step) but lets both phases saturate.
//
//   phase A  k_chain_d16 : one warp per circuit.  Forward chain s_k = G_k s_{k-1} (DFMA, pair-summed dot
//            products) and the backward chains of up to 4 outcomes at once (DMMA: E^T[8x16] . G, D fragment of
//            one step = A fragment of the next through the K relabelling sigma(t,q) = {2q,2q+1,8+2q,9+2q}[t]),
//            interleaved in one loop (independent dependency chains => ILP).  Writes every s_k and e_k row
//            (128 B each) to the scratch, plus e_0 and the probabilities.
//   phase B  k_accum_d16 : one warp per (circuit, outcome), gate by gate:
//            W_g[i][j] = sum_{t: g_t = g} e_t[i] s_t[j]  as DMMA with K = 4 time steps (host-built per-circuit
//            bucket lists), only ONE gate's 16x16 accumulator live at a time (8 doubles per lane), stored to
//            the Jacobian row through the column map as soon as the gate is finished.
// Arithmetic restated from the reference: opcreps.cpp:40-68, effectcreps.cpp:39-45, matrixforwardsim.py:1059-1139.
#pragma once
#include "common.cuh"
#include "kernels_d16.cuh"   // dmma884, D16Args

// scratch rows (16 doubles each) of circuit c, starting at srow[c]:
//   [0 .. L]                         states s_0 .. s_L
//   then per outcome o (in out_ptr order): [0 .. L-1] e_k (vector used at step k), [L] e_0
struct TwoPhaseDev {
    const uint32_t* srow;      // [n_circ+1] first scratch row of each circuit (prefix sum)
    const uint16_t* bperm;     // [n_prop_expanded] steps of each circuit sorted by gate (offsets by circ_ptr)
    const uint16_t* bcnt;      // [n_circ][n_ops] bucket sizes
    double* scratch;
};

#define C2P_WARPS 8

// ------------------------------------------------------------------------------------------------------------
// phase A.  dynamic smem: n_ops*256 doubles (chain B fragments) + n_ops*256 doubles (forward fragments)
//           + C2P_WARPS*2*16 doubles (forward state exchange)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(C2P_WARPS * 32)
k_chain_d16(AtomDev a, ModelDev m, TwoPhaseDev tp, double* __restrict__ probs, int c_begin, int c_end)
{
    extern __shared__ __align__(16) double sm2p[];
    double* bfrag = sm2p;                        // [n_ops][8][32]
    double* ffrag = bfrag + a.n_ops * 256;       // [n_ops][8][32]
    double* fx_all = ffrag + a.n_ops * 256;      // [warps][2][16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* G = m.M;
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;

    for (int idx = threadIdx.x; idx < a.n_ops * 256; idx += blockDim.x) {
        const int g = idx >> 8, r = (idx >> 5) & 7, l = idx & 31;
        const int t = r >> 1, u = r & 1, mr = l >> 2, q = l & 3;
        const int kk = (t >> 1) * 8 + 2 * q + (t & 1);
        bfrag[idx] = G[g * 256 + kk * 16 + 8 * u + mr];                 // G_g[sigma(t,q)][8u + mrow]
        ffrag[idx] = G[g * 256 + (l & 15) * 16 + (l >> 4) * 8 + r];     // G_g[lo][half*8 + r]
    }
    __syncthreads();

    double* fx = fx_all + warp * 32;
    const int lo = lane & 15, half = lane >> 4;
    const int mrow = lane >> 2, q = lane & 3;
    const int gw = blockIdx.x * C2P_WARPS + warp;
    const int nw = gridDim.x * C2P_WARPS;

    for (int c = c_begin + gw; c < c_end; c += nw) {
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const int32_t* ops = a.circ_ops + p0;
        const int o0 = a.out_ptr[c], o1 = a.out_ptr[c + 1];
        double* S = tp.scratch + (size_t)tp.srow[c] * 16;
        const int prep = a.circ_prep[c];

        for (int og = o0; og < o1 || og == o0; og += 4) {
            const bool first = (og == o0);
            const int nin = (o1 - og) < 4 ? (o1 - og) : 4;
            const bool rowok = mrow < nin;
            const int ei = rowok ? a.out_eff[og + mrow] : 0;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (rowok) {
                const double* Er = E + ei * 16;
                a0 = Er[2 * q]; a1 = Er[2 * q + 1]; a2 = Er[8 + 2 * q]; a3 = Er[9 + 2 * q];
            }
            // history base of outcome row mrow (clamped for the padding rows)
            double* H = S + (size_t)(L + 1) * 16 * (1 + (og - o0) + (mrow & 3)) + 2 * q;

            // forward chain only in the first outcome group
            int cur = 0;
            if (first) {
                const double r0 = (lane < 16) ? rho[prep * 16 + lane] : 0.0;
                if (lane < 16) { fx[lane] = r0; S[lane] = r0; }
                __syncwarp();
            }
            int gf = (L > 0) ? ops[0] : 0, gb = (L > 0) ? ops[L - 1] : 0;
            for (int i = 0; i < L; ++i) {
                const int kb = L - 1 - i;
                const int gf_n = (i + 1 < L) ? ops[i + 1] : 0;
                const int gb_n = (kb > 0) ? ops[kb - 1] : 0;
                // ---- backward step kb (4 outcomes): store e_kb, E_new^T = E^T . G ----
                if (mrow < 4 && rowok) {
                    double* hp = H + (size_t)kb * 16;
                    *reinterpret_cast<double2*>(hp) = make_double2(a0, a1);
                    *reinterpret_cast<double2*>(hp + 8) = make_double2(a2, a3);
                }
                const double* bp = bfrag + gb * 256 + lane;
                double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0, x00 = 0.0, x01 = 0.0, x10 = 0.0, x11 = 0.0;
                dmma884(d00, d01, a0, bp[0]);   dmma884(d10, d11, a0, bp[32]);
                dmma884(x00, x01, a1, bp[64]);  dmma884(x10, x11, a1, bp[96]);
                dmma884(d00, d01, a2, bp[128]); dmma884(d10, d11, a2, bp[160]);
                dmma884(x00, x01, a3, bp[192]); dmma884(x10, x11, a3, bp[224]);
                // ---- forward step i ----
                if (first) {
                    const double2* s2 = reinterpret_cast<const double2*>(fx + cur * 16 + half * 8);
                    const double2 s0 = s2[0], s1 = s2[1], s2v = s2[2], s3 = s2[3];
                    const double* fp = ffrag + gf * 256 + lane;
                    double f0 = fp[0] * s0.x, f1 = fp[32] * s0.y, f2 = fp[64] * s1.x, f3 = fp[96] * s1.y;
                    f0 = fma(fp[128], s2v.x, f0); f1 = fma(fp[160], s2v.y, f1);
                    f2 = fma(fp[192], s3.x, f2); f3 = fma(fp[224], s3.y, f3);
                    double v = (f0 + f1) + (f2 + f3);
                    v += shfl_xor_f64(v, 16);
                    cur ^= 1;
                    if (lane < 16) { fx[cur * 16 + lane] = v; S[(size_t)(i + 1) * 16 + lane] = v; }
                    __syncwarp();
                }
                a0 = d00 + x00; a1 = d01 + x01; a2 = d10 + x10; a3 = d11 + x11;
                gf = gf_n; gb = gb_n;
            }
            // e_0 row
            if (mrow < 4 && rowok) {
                double* hp = H + (size_t)L * 16;
                *reinterpret_cast<double2*>(hp) = make_double2(a0, a1);
                *reinterpret_cast<double2*>(hp + 8) = make_double2(a2, a3);
            }
            if (probs && nin > 0) {
                // p_o = E_o . s_L ; s_L is in the scratch row L (written by this warp) -- for later groups re-read it
                __syncwarp();
                const double* sL = first ? (fx + cur * 16) : nullptr;
                double pr = 0.0;
                if (rowok) {
                    const double* Er = E + ei * 16;
                    if (first) pr = Er[2 * q] * sL[2 * q] + Er[2 * q + 1] * sL[2 * q + 1] + Er[8 + 2 * q] * sL[8 + 2 * q] + Er[9 + 2 * q] * sL[9 + 2 * q];
                    else {
                        const double* sg = S + (size_t)L * 16;
                        pr = Er[2 * q] * sg[2 * q] + Er[2 * q + 1] * sg[2 * q + 1] + Er[8 + 2 * q] * sg[8 + 2 * q] + Er[9 + 2 * q] * sg[9 + 2 * q];
                    }
                }
                pr += shfl_xor_f64(pr, 1);
                pr += shfl_xor_f64(pr, 2);
                if (rowok && q == 0) probs[a.out_el[og + mrow]] = pr;
            }
            __syncwarp();
            if (o1 == o0) break;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase B.  one warp per (circuit, outcome); grid-stride over circuits, 4 outcome warps per circuit slot.
// dynamic smem: n_ops*4*32 int2 (column map fragments) + 2*SPAM_MAX ints
// ------------------------------------------------------------------------------------------------------------
#define A2P_WARPS 8      // 2 circuits x 4 outcomes per CTA

__global__ void __launch_bounds__(A2P_WARPS * 32)
k_accum_d16(AtomDev a, ModelDev m, TwoPhaseDev tp, D16Args args, int c_begin, int c_end)
{
    extern __shared__ __align__(16) unsigned char sm2b[];
    int2* cm_s = reinterpret_cast<int2*>(sm2b);                         // [n_ops*4][32]
    int* spamc_s = reinterpret_cast<int*>(cm_s + a.n_ops * 4 * 32);     // [SPAM_MAX]
    int* spamw_s = spamc_s + D16_SPAM_MAX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < a.n_ops * 4 * 32; idx += blockDim.x) {
        const int g = idx >> 7, tile = (idx >> 5) & 3, l = idx & 31;
        const int i = 8 * (tile >> 1) + (l >> 2), jc = 8 * (tile & 1) + 2 * (l & 3);
        int2 cc = *reinterpret_cast<const int2*>(args.colmap + g * 256 + i * 16 + jc);
        if (cc.y == cc.x + 1 && cc.x >= 0 && ((cc.x | (int)(args.ld & 1)) & 1) == 0) cc.y = -2;
        cm_s[idx] = cc;
    }
    const int n_spam_s = args.n_spam < D16_SPAM_MAX ? args.n_spam : D16_SPAM_MAX;
    for (int t = threadIdx.x; t < n_spam_s; t += blockDim.x) { spamc_s[t] = args.spam_col[t]; spamw_s[t] = args.spam_w[t]; }
    __syncthreads();

    const int mrow = lane >> 2, q = lane & 3;
    const int slot = warp >> 2, ow = warp & 3;             // circuit slot within the CTA, outcome within group
    const int cstride = gridDim.x * (A2P_WARPS / 4);
    for (int c = c_begin + blockIdx.x * (A2P_WARPS / 4) + slot; c < c_end; c += cstride) {
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const int o0 = a.out_ptr[c], o1 = a.out_ptr[c + 1];
        const int prep = a.circ_prep[c];
        const double* S = tp.scratch + (size_t)tp.srow[c] * 16;
        const uint16_t* pa = tp.bperm + p0;
        const uint16_t* ca = tp.bcnt + (size_t)c * a.n_ops;
        for (int qo = o0 + ow; qo < o1; qo += 4) {
            const int ei = a.out_eff[qo];
            const int64_t el = a.out_el[qo];
            const double* H = S + (size_t)(L + 1) * 16 * (1 + (qo - o0));
            double* Jr = args.J + el * args.ld;
            const double* hb = H + mrow;
            const double* sb = S + mrow;
            int tb = 0;
            for (int g = 0; g < a.n_ops; ++g) {
                const int cg = ca[g];
                double acc[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r] = 0.0;
                for (int t0 = 0; t0 < cg; t0 += 4) {
                    const bool v = (t0 + q) < cg;
                    const int s = v ? pa[tb + t0 + q] : 0;
                    const double* hp = hb + (size_t)s * 16;
                    const double* sp = sb + (size_t)s * 16;
                    const double ea = v ? hp[0] : 0.0, eb = v ? hp[8] : 0.0;
                    const double sa = v ? sp[0] : 0.0, sbv = v ? sp[8] : 0.0;
                    dmma884(acc[0], acc[1], ea, sa);
                    dmma884(acc[2], acc[3], ea, sbv);
                    dmma884(acc[4], acc[5], eb, sa);
                    dmma884(acc[6], acc[7], eb, sbv);
                }
                tb += cg;
                const int2* cm = cm_s + g * 128 + lane;
#pragma unroll
                for (int tile = 0; tile < 4; ++tile) {
                    const int2 cc = cm[tile * 32];
                    const double v0 = acc[tile * 2], v1 = acc[tile * 2 + 1];
                    if (cc.y == -2) {
                        *reinterpret_cast<double2*>(Jr + cc.x) = make_double2(v0, v1);
                    } else {
                        if (cc.x >= 0) Jr[cc.x] = v0;
                        if (cc.y >= 0) Jr[cc.y] = v1;
                    }
                }
            }
            // SPAM / unmapped columns
            const int w_rho0 = (int)m.off_rho + prep * 16, w_eff0 = (int)m.off_eff + ei * 16;
            const double* e0 = H + (size_t)L * 16;
            const double* sL = S + (size_t)L * 16;
            for (int t = lane; t < args.n_spam; t += 32) {
                const int w = t < D16_SPAM_MAX ? spamw_s[t] : args.spam_w[t];
                const int col = t < D16_SPAM_MAX ? spamc_s[t] : args.spam_col[t];
                double val = 0.0;
                if (w >= w_rho0 && w < w_rho0 + 16) val = e0[w - w_rho0];
                else if (w >= w_eff0 && w < w_eff0 + 16) val = sL[w - w_eff0];
                Jr[col] = val;
            }
        }
    }
}
