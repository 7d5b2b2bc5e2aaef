// kernels_d16.cuh -- the 2-qubit (d = 16) Jacobian kernel: the BASELINE.json headline path.
//
// One CTA (4 warps) owns one circuit at a time (persistent, longest circuits first).
//   warp 0      : forward chain s_k = G_k s_{k-1}, all L+1 states kept in shared memory
//   warp w      : outcome w, w+4, ... of the circuit: backward chain e_{k-1} = G_k^T e_k fused with the
//                 rank-1 accumulation  W_g += e_k (x) s_{k-1}  into REGISTER accumulators
//                 (NG gates x 16x16 doubles spread over the 32 lanes = NG*8 doubles per lane)
//   epilogue    : the accumulators are the Jacobian row for a fully parameterised gate
//                 (dG[i,j]/dtheta_p = delta_{p,(i,j)}); they are stored straight to J through a column
//                 map (colmap[w] = Jacobian column of member element w, or -1), 128-byte coalesced.
// The only HBM traffic that scales with the problem is the Jacobian store: 8*(n_params+1) bytes per
// circuit outcome -- this kernel is HBM-write bound by construction.
//
// Arithmetic restated from the reference: dense acton (opcreps.cpp:40-54), adjoint acton
// (opcreps.cpp:56-68), effect dot (effectcreps.cpp:39-45); derivative structure
// matrixforwardsim.py:1059-1139 with dprod = sum_k (suffix) dG_k (prefix) (:729-792).
#pragma once
#include "common.cuh"

#define D16_WARPS 4

struct D16Args {
    const int32_t* colmap;    // [n_w]  J column of W index w (gate part used by the register epilogue), -1 = none
    const int32_t* spam_col;  // [n_spam] columns NOT fed by a gate element ...
    const int32_t* spam_w;    // [n_spam] ... and the rho/effect W index feeding each (or -1 -> zero)
    int n_spam;
    double* J;                // [n_elements][ld]
    int64_t ld;
    double* probs;            // [n_elements] or nullptr
};

// shared memory: gf/gb fragments (NG*4*32 double2 each), states (max_depth+1)*16, evec 4*2*16
__host__ __device__ inline size_t d16_smem_bytes(int ng, int max_depth) {
    return (size_t)ng * 4 * 32 * 16 * 2 + (size_t)(max_depth + 1) * 16 * 8 + D16_WARPS * 2 * 16 * 8;
}

template <int NG>
__global__ void __launch_bounds__(D16_WARPS * 32, 4)
k_dprobs_d16(AtomDev a, ModelDev m, D16Args args)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* gf = reinterpret_cast<double2*>(smem_raw);          // forward fragments  G[i][half*8+2q..]
    double2* gb = gf + NG * 4 * 32;                               // backward fragments G[half*8+2q..][j]
    double* states = reinterpret_cast<double*>(gb + NG * 4 * 32); // [(max_depth+1)][16]
    double* evec = states + (size_t)(a.max_depth + 1) * 16;       // [warp][2][16]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lo = lane & 15, half = lane >> 4;

    const double* G = m.M;
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;

    // stage gate fragments once per CTA (persistent kernel)
    for (int idx = threadIdx.x; idx < a.n_ops * 4 * 32; idx += blockDim.x) {
        const int g = idx >> 7, q = (idx >> 5) & 3, l = idx & 31;
        const int r = l & 15, h = l >> 4;
        const double* Gg = G + g * 256;
        gf[idx] = make_double2(Gg[r * 16 + h * 8 + 2 * q], Gg[r * 16 + h * 8 + 2 * q + 1]);
        gb[idx] = make_double2(Gg[(h * 8 + 2 * q) * 16 + r], Gg[(h * 8 + 2 * q + 1) * 16 + r]);
    }
    __syncthreads();

    double* ev = evec + warp * 32;

    for (int c = blockIdx.x; c < a.n_circ; c += gridDim.x) {
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const int32_t* ops = a.circ_ops + p0;
        const int prep = a.circ_prep[c];

        // ---------------- forward chain (warp 0) ----------------
        if (warp == 0) {
            if (lane < 16) states[lane] = rho[prep * 16 + lane];
            __syncwarp();
            for (int k = 0; k < L; ++k) {
                const int g = ops[k];
                const double2* s2 = reinterpret_cast<const double2*>(states + k * 16 + half * 8);
                const double2* f = gf + g * 128 + lane;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 sv = s2[q];
                    const double2 fv = f[q * 32];
                    a0 = fma(fv.x, sv.x, a0);
                    a1 = fma(fv.y, sv.y, a1);
                }
                double v = a0 + a1;
                v += shfl_xor_f64(v, 16);
                if (lane < 16) states[(k + 1) * 16 + lane] = v;
                __syncwarp();
            }
        }
        __syncthreads();

        // ---------------- backward chains + accumulation (one outcome per warp) ----------------
        const int o0 = a.out_ptr[c], o1 = a.out_ptr[c + 1];
        for (int q = o0 + warp; q < o1; q += D16_WARPS) {
            const int ei = a.out_eff[q];
            const int64_t el = a.out_el[q];
            double acc[NG][8];
#pragma unroll
            for (int g = 0; g < NG; ++g)
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[g][r] = 0.0;

            const double e_init = E[ei * 16 + lo];
            const double sL = states[L * 16 + lo];
            if (args.probs) {
                double pr = (half == 0) ? e_init * sL : 0.0;
#pragma unroll
                for (int mk = 16; mk > 0; mk >>= 1) pr += shfl_xor_f64(pr, mk);
                if (lane == 0) args.probs[el] = pr;
            }
            int cur = 0;
            if (lane < 16) ev[lane] = e_init;
            __syncwarp();
            double e_own = e_init;   // e[lo]
            for (int k = L - 1; k >= 0; --k) {
                const int g = ops[k];
                const double2* e2 = reinterpret_cast<const double2*>(ev + cur * 16 + half * 8);
                double e8[8];
#pragma unroll
                for (int r = 0; r < 4; ++r) { const double2 t = e2[r]; e8[2 * r] = t.x; e8[2 * r + 1] = t.y; }
                const double sj = states[k * 16 + lo];
                // rank-1 update of this gate's accumulator: W_g[half*8+r][lo] += e[half*8+r] * s[lo]
                switch (g) {
#define D16_CASE(GI) case GI: if (GI < NG) { _Pragma("unroll") for (int r = 0; r < 8; ++r) acc[GI < NG ? GI : 0][r] = fma(e8[r], sj, acc[GI < NG ? GI : 0][r]); } break;
                    D16_CASE(0) D16_CASE(1) D16_CASE(2) D16_CASE(3)
                    D16_CASE(4) D16_CASE(5) D16_CASE(6) D16_CASE(7)
#undef D16_CASE
                    default: break;
                }
                // e_new[lo] = sum_i G[i][lo] e[i]   (this lane: i in half*8..half*8+7, then pair-sum)
                const double2* f = gb + g * 128 + lane;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const double2 fv = f[r * 32];
                    a0 = fma(fv.x, e8[2 * r], a0);
                    a1 = fma(fv.y, e8[2 * r + 1], a1);
                }
                double v = a0 + a1;
                v += shfl_xor_f64(v, 16);
                e_own = v;
                cur ^= 1;
                if (lane < 16) ev[cur * 16 + lane] = v;
                __syncwarp();
            }

            // ---------------- epilogue: store the Jacobian row ----------------
            double* Jr = args.J + el * args.ld;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g < a.n_ops) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int w = g * 256 + (half * 8 + r) * 16 + lo;
                        const int col = __ldg(args.colmap + w);
                        if (col >= 0) Jr[col] = acc[g][r];
                    }
                }
            }
            // SPAM / unmapped columns: d p/d rho = e_0, d p/d E_j = s_L for the outcome's own effect, else 0
            const int64_t w_rho0 = m.off_rho + (int64_t)prep * 16, w_eff0 = m.off_eff + (int64_t)ei * 16;
            // publish e_0 for arbitrary-lane access
            if (lane < 16) ev[cur * 16 + lane] = e_own;
            __syncwarp();
            for (int t = lane; t < args.n_spam; t += 32) {
                const int64_t w = args.spam_w[t];
                double val = 0.0;
                if (w >= w_rho0 && w < w_rho0 + 16) val = ev[cur * 16 + (int)(w - w_rho0)];
                else if (w >= w_eff0 && w < w_eff0 + 16) val = states[L * 16 + (int)(w - w_eff0)];
                Jr[args.spam_col[t]] = val;
            }
            __syncwarp();
        }
        __syncthreads();   // states are overwritten by the next circuit
    }
}
