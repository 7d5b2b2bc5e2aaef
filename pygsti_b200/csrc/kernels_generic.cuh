// kernels_generic.cuh -- dimension-generic FP64 kernels (any d = 4^n, any number of gates).
//
// These are the correctness-first members of the kernel family: one warp owns one circuit and walks
// its op sequence; they are used for dimensions / gate counts the specialised kernels do not cover,
// for the forward-difference mode and for the Hessian.  Arithmetic restated from the reference:
//   state propagation  out[i] = sum_j G[i][j] v[j]   (pygsti/evotypes/densitymx/opcreps.cpp:40-54)
//   outcome            p = sum_i E[i] v[i]           (pygsti/evotypes/densitymx/effectcreps.cpp:39-45)
//   circuit walk       pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:224-283
#pragma once
#include "common.cuh"

#define GEN_WARPS 4   // warps per CTA in the generic kernels

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

// ---------------------------------------------------------------------------------------------
// probs: out[b*batch_stride + el*el_stride] for model b = blockIdx.y (b = 0 for the plain call;
// b > 0 are the perturbed models of the forward-difference mode).
// dynamic smem: GEN_WARPS * 2 * D doubles.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(GEN_WARPS * 32)
k_probs_generic(AtomDev a, const double* __restrict__ Mbase, const double* __restrict__ Gtbase,
                int64_t n_w, double* __restrict__ out, int64_t el_stride, int64_t batch_stride)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* buf0 = smem + (size_t)warp * 2 * D;
    double* buf1 = buf0 + D;
    const int b = blockIdx.y;
    const double* M = Mbase + (int64_t)b * n_w;
    const double* Gt = Gtbase + (int64_t)b * a.n_ops * D * D;
    const double* rho = M + (int64_t)a.n_ops * D * D;
    const double* E = rho + (int64_t)a.n_rho * D;
    double* o = out + (int64_t)b * batch_stride;

    const int gw = blockIdx.x * GEN_WARPS + warp;
    const int nw = gridDim.x * GEN_WARPS;
    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c], p1 = a.circ_ptr[c + 1];
        const double* r = rho + (int64_t)a.circ_prep[c] * D;
        double* s = buf0; double* t = buf1;
        for (int i = lane; i < D; i += 32) s[i] = r[i];
        __syncwarp();
        for (uint32_t k = p0; k < p1; ++k) {
            const double* Gg = Gt + (int64_t)a.circ_ops[k] * D * D;
            for (int i = lane; i < D; i += 32) {
                double acc = 0.0;
#pragma unroll 8
                for (int j = 0; j < D; ++j) acc += Gg[j * D + i] * s[j];
                t[i] = acc;
            }
            __syncwarp();
            double* x = s; s = t; t = x;
        }
        for (int q = a.out_ptr[c]; q < a.out_ptr[c + 1]; ++q) {
            const double* e = E + (int64_t)a.out_eff[q] * D;
            double part = 0.0;
            for (int i = lane; i < D; i += 32) part += e[i] * s[i];
            part = warp_sum(part);
            if (lane == 0) o[(int64_t)a.out_el[q] * el_stride] = part;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// W accumulate: W[el][w] = d p_el / d(member element w), W pre-zeroed, row stride ldw.
// One warp per circuit; forward states kept in a per-warp global scratch (L2 resident).
// Adjoint recursion (SURVEY.md App. B.2 == matrixforwardsim.py:1059-1139 for fully
// parameterised members):
//   s_0 = rho, s_k = G_k s_{k-1};   e_L = E_j, e_{k-1} = G_k^T e_k
//   W[el, G_k block] += e_k (x) s_{k-1};  W[el, rho] += e_0;  W[el, E_j] += s_L
// dynamic smem: GEN_WARPS * 2 * D doubles.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(GEN_WARPS * 32)
k_w_generic(AtomDev a, ModelDev m, double* __restrict__ W, int64_t ldw, double* __restrict__ probs,
            double* __restrict__ scratch /* [total warps][(max_depth+1)*D] */)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* eb0 = smem + (size_t)warp * 2 * D;
    double* eb1 = eb0 + D;
    const double* G = m.M;
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;
    const int gw = blockIdx.x * GEN_WARPS + warp;
    const int nw = gridDim.x * GEN_WARPS;
    double* st = scratch + (int64_t)gw * (a.max_depth + 1) * D;

    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const int prep = a.circ_prep[c];
        const int32_t* ops = a.circ_ops + p0;
        for (int i = lane; i < D; i += 32) st[i] = rho[(int64_t)prep * D + i];
        __syncwarp();
        for (int k = 0; k < L; ++k) {
            const double* Gg = m.Gt + (int64_t)ops[k] * D * D;
            const double* s = st + (int64_t)k * D;
            double* t = st + (int64_t)(k + 1) * D;
            for (int i = lane; i < D; i += 32) {
                double acc = 0.0;
#pragma unroll 8
                for (int j = 0; j < D; ++j) acc += Gg[j * D + i] * s[j];
                t[i] = acc;
            }
            __syncwarp();
        }
        const double* sL = st + (int64_t)L * D;
        for (int q = a.out_ptr[c]; q < a.out_ptr[c + 1]; ++q) {
            const int ei = a.out_eff[q];
            const int64_t el = a.out_el[q];
            double* Wr = W + el * ldw;
            const double* ev = E + (int64_t)ei * D;
            double part = 0.0;
            for (int i = lane; i < D; i += 32) {
                part += ev[i] * sL[i];
                Wr[m.off_eff + (int64_t)ei * D + i] += sL[i];
            }
            if (probs) { part = warp_sum(part); if (lane == 0) probs[el] = part; }
            double* e = eb0; double* en = eb1;
            for (int i = lane; i < D; i += 32) e[i] = ev[i];
            __syncwarp();
            for (int k = L - 1; k >= 0; --k) {
                const int g = ops[k];
                const double* s = st + (int64_t)k * D;
                double* Wg = Wr + (int64_t)g * D * D;
                for (int idx = lane; idx < D * D; idx += 32) {
                    const int i = idx / D, j = idx - i * D;
                    Wg[idx] += e[i] * s[j];
                }
                const double* Gg = G + (int64_t)g * D * D;
                for (int j = lane; j < D; j += 32) {
                    double acc = 0.0;
#pragma unroll 8
                    for (int i = 0; i < D; ++i) acc += Gg[i * D + j] * e[i];
                    en[j] = acc;
                }
                __syncwarp();
                double* x = e; e = en; en = x;
            }
            for (int i = lane; i < D; i += 32) Wr[m.off_rho + (int64_t)prep * D + i] += e[i];
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tangent W accumulate (Hessian of members LINEAR in their parameters; forward-over-reverse).
// For the tangent direction dM = dM/dtheta_p (batch b = blockIdx.y; dM[b] = [dG | drho | dE], dGt[b] its
// transposed gates):   ds_0 = drho, ds_k = G_k ds_{k-1} + dG_k s_{k-1};   de_L = dE_j,
// de_{k-1} = G_k^T de_k + dG_k^T e_k;   d/dtheta_p of the W row:
//   Wp[el, G_k block] += de_k (x) s_{k-1} + e_k (x) ds_{k-1};  Wp[el, rho] += de_0;  Wp[el, E_j] += ds_L
// so that  H[el, p, q] = sum_w Wp[el, w] D[w, q]   (second derivatives of the members vanish).
// Equals MatrixForwardSimulator._hprobs_from_rho_e (matrixforwardsim.py:1141-1287) for linear members.
// Scratch per warp: 2 * (max_depth+1) * D doubles (states and tangent states).
// dynamic smem: GEN_WARPS * 4 * D doubles.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(GEN_WARPS * 32)
k_w_tangent_generic(AtomDev a, ModelDev m, const double* __restrict__ dMb, const double* __restrict__ dGtb,
                    double* __restrict__ Wb, int64_t ldw, double* __restrict__ scratch, const unsigned char* __restrict__ need)
{
    // need[g] (g < n_ops), need[n_ops] (state preparations), need[n_ops + 1] (effects): only these blocks of the W row are
    // read by the contraction that follows (the rectangle's second parameter axis touches one or two members), so only
    // these are accumulated; nullptr = all.
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* eb0 = smem + (size_t)warp * 4 * D;
    double* eb1 = eb0 + D; double* db0 = eb1 + D; double* db1 = db0 + D;
    const int b = blockIdx.y;
    const double* dM = dMb + (int64_t)b * m.n_w;
    const double* dGt = dGtb + (int64_t)b * a.n_ops * D * D;
    double* W = Wb + (int64_t)b * a.n_elements * ldw;
    const double* G = m.M; const double* rho = m.M + m.off_rho; const double* E = m.M + m.off_eff;
    const double* dG = dM; const double* drho = dM + m.off_rho; const double* dE = dM + m.off_eff;
    const int gw = blockIdx.x * GEN_WARPS + warp;
    const int nw = gridDim.x * GEN_WARPS;
    double* st = scratch + ((int64_t)b * nw + gw) * 2 * (a.max_depth + 1) * D;
    double* dst = st + (int64_t)(a.max_depth + 1) * D;

    for (int c = gw; c < a.n_circ; c += nw) {
        const uint32_t p0 = a.circ_ptr[c];
        const int L = (int)(a.circ_ptr[c + 1] - p0);
        const int prep = a.circ_prep[c];
        const int32_t* ops = a.circ_ops + p0;
        for (int i = lane; i < D; i += 32) { st[i] = rho[(int64_t)prep * D + i]; dst[i] = drho[(int64_t)prep * D + i]; }
        __syncwarp();
        for (int k = 0; k < L; ++k) {
            const double* Gg = m.Gt + (int64_t)ops[k] * D * D;
            const double* dGg = dGt + (int64_t)ops[k] * D * D;
            const double* s = st + (int64_t)k * D; const double* ds = dst + (int64_t)k * D;
            for (int i = lane; i < D; i += 32) {
                double acc = 0.0, dacc = 0.0;
                for (int j = 0; j < D; ++j) {
                    acc += Gg[j * D + i] * s[j];
                    dacc += Gg[j * D + i] * ds[j] + dGg[j * D + i] * s[j];
                }
                st[(int64_t)(k + 1) * D + i] = acc; dst[(int64_t)(k + 1) * D + i] = dacc;
            }
            __syncwarp();
        }
        const double* dsL = dst + (int64_t)L * D;
        for (int q = a.out_ptr[c]; q < a.out_ptr[c + 1]; ++q) {
            const int ei = a.out_eff[q];
            double* Wr = W + (int64_t)a.out_el[q] * ldw;
            double* e = eb0; double* en = eb1; double* de = db0; double* den = db1;
            for (int i = lane; i < D; i += 32) {
                e[i] = E[(int64_t)ei * D + i]; de[i] = dE[(int64_t)ei * D + i];
                if (!need || need[a.n_ops + 1]) Wr[m.off_eff + (int64_t)ei * D + i] += dsL[i];
            }
            __syncwarp();
            for (int k = L - 1; k >= 0; --k) {
                const int g = ops[k];
                const double* s = st + (int64_t)k * D; const double* ds = dst + (int64_t)k * D;
                double* Wg = Wr + (int64_t)g * D * D;
                if (!need || need[g]) {
                    for (int idx = lane; idx < D * D; idx += 32) {
                        const int i = idx / D, j = idx - i * D;
                        Wg[idx] += de[i] * s[j] + e[i] * ds[j];
                    }
                }
                const double* Gg = G + (int64_t)g * D * D;
                const double* dGg = dG + (int64_t)g * D * D;
                for (int j = lane; j < D; j += 32) {
                    double acc = 0.0, dacc = 0.0;
                    for (int i = 0; i < D; ++i) {
                        acc += Gg[i * D + j] * e[i];
                        dacc += Gg[i * D + j] * de[i] + dGg[i * D + j] * e[i];
                    }
                    en[j] = acc; den[j] = dacc;
                }
                __syncwarp();
                double* x = e; e = en; en = x; x = de; de = den; den = x;
            }
            if (!need || need[a.n_ops]) for (int i = lane; i < D; i += 32) Wr[m.off_rho + (int64_t)prep * D + i] += de[i];
            __syncwarp();
        }
    }
}

// dense tangent models: dMb[b][w] = D[w, cols[b0 + b]]
__global__ void k_tangent_models(int64_t n_w, const int32_t* __restrict__ cols, int b0,
                                 const int32_t* __restrict__ cptr, const int32_t* __restrict__ crow,
                                 const double* __restrict__ cval, double* __restrict__ dMb)
{
    const int b = blockIdx.y;
    const int p = cols[b0 + b];
    double* dst = dMb + (int64_t)b * n_w;
    for (int t = cptr[p] + blockIdx.x * blockDim.x + threadIdx.x; t < cptr[p + 1]; t += gridDim.x * blockDim.x)
        dst[crow[t]] = cval[t];
}

// out[((el * n1) + a0 + b) * n2 + c] = sum_t cval[t] * Wb[b][el][crow[t]]   over column cols2[c] of D
__global__ void __launch_bounds__(128)
k_contract_hess(const double* __restrict__ Wb, int64_t ldw, int64_t n_el, int nb, int a0, int n1, int n2,
                const int32_t* __restrict__ cols2, const int32_t* __restrict__ cptr,
                const int32_t* __restrict__ crow, const double* __restrict__ cval, double* __restrict__ out, int accumulate)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= n2) return;
    const int p = cols2[c];
    const int tb = cptr[p], te = cptr[p + 1];
    for (int64_t r = blockIdx.y; r < n_el * nb; r += gridDim.y) {
        const int64_t el = r / nb; const int b = (int)(r - el * nb);
        const double* Wr = Wb + ((int64_t)b * n_el + el) * ldw;
        double acc = 0.0;
        for (int t = tb; t < te; ++t) acc += cval[t] * Wr[crow[t]];
        double* o = out + (el * n1 + a0 + b) * n2 + c;
        *o = accumulate ? *o + acc : acc;
    }
}

// Second-derivative term of the Hessian for members NOT linear in their parameters:
//   out[(el * n1 + a) * n2 + b] = sum_t kval[t] * W[el][krow[t]]   over the entries t of key (a, b) = ukey[k]
// (d2M_w / dtheta_a dtheta_b from member.hessian_wrt_params; == the `_hoperation` / SPAM-hessian terms of
// MatrixForwardSimulator._hprobs_from_rho_e, matrixforwardsim.py:1196-1237).  One thread per key, grid.y over elements.
__global__ void __launch_bounds__(128)
k_hess_d2(const double* __restrict__ W, int64_t ldw, int64_t n_el, int n1, int n2, int nk,
          const int64_t* __restrict__ ukey, const int32_t* __restrict__ kptr, const int32_t* __restrict__ krow,
          const double* __restrict__ kval, double* __restrict__ out)
{
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= nk) return;
    const int64_t key = ukey[k];
    const int tb = kptr[k], te = kptr[k + 1];
    for (int64_t el = blockIdx.y; el < n_el; el += gridDim.y) {
        const double* Wr = W + el * ldw;
        double acc = 0.0;
        for (int t = tb; t < te; ++t) acc += kval[t] * Wr[krow[t]];
        out[el * (int64_t)n1 * n2 + key] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// J = W . D  with D in CSC (per parameter column p: entries [cptr[p], cptr[p+1]) of (row w, val)).
// grid.x over column tiles of 128, grid.y over element rows; coalesced stores along p.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_contract_csc(const double* __restrict__ W, int64_t ldw, int64_t n_el, int n_params,
               const int32_t* __restrict__ cptr, const int32_t* __restrict__ crow,
               const double* __restrict__ cval, double* __restrict__ J, int64_t ldj,
               const double* __restrict__ row_scale)
{
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_params) return;
    const int b = cptr[p], e = cptr[p + 1];
    for (int64_t el = blockIdx.y; el < n_el; el += gridDim.y) {
        const double* Wr = W + el * ldw;
        double acc = 0.0;
        for (int t = b; t < e; ++t) acc += cval[t] * Wr[crow[t]];
        J[el * ldj + p] = row_scale ? acc * row_scale[el] : acc;
    }
}

// ---------------------------------------------------------------------------------------------
// forward-difference helpers (reference semantics, pyx:349-378)
// ---------------------------------------------------------------------------------------------
// perturbed models, step 1: Mb[b] = M   (step 2, k_perturb_apply, adds eps * D[:, p0+b]; then the gates are transposed)
__global__ void k_perturb_models(const double* __restrict__ M, int64_t n_w, double* __restrict__ Mb)
{
    const int b = blockIdx.y;
    double* dst = Mb + (int64_t)b * n_w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_w; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = M[i];
}
__global__ void k_perturb_apply(int64_t n_w, int p0, double eps,
                                const int32_t* __restrict__ cptr, const int32_t* __restrict__ crow,
                                const double* __restrict__ cval, double* __restrict__ Mb)
{
    const int b = blockIdx.y;
    double* dst = Mb + (int64_t)b * n_w;
    const int beg = cptr[p0 + b], end = cptr[p0 + b + 1];
    for (int t = beg + blockIdx.x * blockDim.x + threadIdx.x; t < end; t += gridDim.x * blockDim.x)
        dst[crow[t]] += eps * cval[t];   // CSC rows are unique within a column (duplicates summed at upload)
}
// Gt[b][g][j][i] = M[b][g][i][j]
__global__ void k_transpose_gates(const double* __restrict__ Mb, int64_t n_w, int n_ops, int D,
                                  double* __restrict__ Gtb)
{
    const int b = blockIdx.y;
    const double* G = Mb + (int64_t)b * n_w;
    double* Gt = Gtb + (int64_t)b * n_ops * D * D;
    const int64_t n = (int64_t)n_ops * D * D;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = idx / (D * D); const int r = (int)(idx - g * D * D);
        const int j = r / D, i = r - j * D;
        Gt[idx] = G[g * D * D + (int64_t)i * D + j];
    }
}
// out[el*ld + p0 + b] = (Pb[b][el] - P0[el]) / eps
__global__ void k_fd_finish(const double* __restrict__ Pb, const double* __restrict__ P0, int64_t n_el,
                            int nb, int p0, double eps, double* __restrict__ out, int64_t ld)
{
    const int b = threadIdx.x & 31;            // 32 columns per block-row for semi-coalesced stores
    const int64_t el0 = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int bb = b; bb < nb; bb += 32)
        for (int64_t el = el0; el < n_el; el += (int64_t)gridDim.x * (blockDim.x >> 5))
            out[el * ld + p0 + bb] = (Pb[(int64_t)bb * n_el + el] - P0[el]) / eps;
}
