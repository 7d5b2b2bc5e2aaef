// kernels_lindblad.cuh -- dense algebra of Lindblad-parameterised members on the device (second part).
//
// Replaces, per parameter-vector update of a CPTPLND / H+S / GLND model, the host work of
//   LindbladErrorgen._update_rep            L = Re sum_i c_i B_i                        (lindbladerrorgen.py:700-708)
//   LindbladErrorgen.deriv_wrt_params       dL_p = Re sum_i dc_i/dtheta_p B_i           (lindbladerrorgen.py:1342-1384, lindbladcoefficients.py:943-990)
//   ExpErrorgenOp._update_rep / deriv       E = expm(L), dE_p = Dexp(L)[dL_p]           (experrorgenop.py:114-125, 213-262)
//   ComposedOp / ComposedState / ComposedPOVMEffect to_dense + deriv_wrt_params         (composition with the static part)
// The coefficients c and their Jacobian dc/dtheta stay on the host (cheap, parameterisation-specific).  The per-thread arithmetic is
// csrc/lindblad_core.h (validated on the host).  Small problem (BASELINE config 4: 7 error generators x 240 parameters, d = 16): three
// launches, one thread per output element / per (generator, parameter); no attempt at tensor cores.
#pragma once
#include "lindblad_core.h"

struct LindDev {
    int d, n_eg, n_mem;
    const int* eg_ncoeff;      // [n_eg]
    const int* eg_npar;        // [n_eg]
    const long long* eg_boff;  // [n_eg] offset of B_e (in units of d*d matrices) = prefix sum of n_coeff
    const long long* eg_doff;  // [n_eg] offset of dc_e (elements) = prefix sum of n_coeff * n_par
    const long long* eg_poff;  // [n_eg + 1] prefix sum of n_par (index of (e, p) pairs)
    const double* B_re; const double* B_im;      // [sum n_coeff][d*d]
    const double* c_re; const double* c_im;      // [sum n_coeff]
    const double* dc_re; const double* dc_im;    // per generator [n_coeff][n_par]
    double* L;                 // [n_eg][d*d]
    double* dL;                // [sum n_par][d*d]
    double* E;                 // [n_eg][d*d]
    double* dE;                // [sum n_par][d*d]
    double* work;              // [sum n_par + n_eg][7*d*d]  (per thread: its own copy of E + the 6 d^2 work area of lb_expm_frechet)
    // members
    const int* m_kind;         // 0 op, 1 state, 2 effect
    const int* m_eg;
    const long long* m_soff;   // offset of the static part (doubles)
    const double* stat;
    double* val; double* dval;
};

// stage 1: L_e and dL_{e,p}; one thread per (row of the (e, p) list incl. the value row, matrix element)
__global__ void k_lind_errgen(LindDev a, long long n_rows /* = sum n_par + n_eg */) {
    const int n = a.d * a.d;
    const long long total = n_rows * n;
    const long long npar_total = a.eg_poff[a.n_eg];
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / n; const int el = (int)(idx - row * n);
        int e; long long p = -1;
        if (row < npar_total) {
            e = 0; while (row >= a.eg_poff[e + 1]) ++e;
            p = row - a.eg_poff[e];
        } else e = (int)(row - npar_total);
        const int nc = a.eg_ncoeff[e], np_ = a.eg_npar[e];
        const double* br = a.B_re + a.eg_boff[e] * n + el;
        const double* bi = a.B_im + a.eg_boff[e] * n + el;
        double acc = 0.0;
        if (p < 0) {
            const double* cr = a.c_re + a.eg_boff[e]; const double* ci = a.c_im + a.eg_boff[e];
            for (int t = 0; t < nc; ++t) acc = fma(cr[t], br[(long long)t * n], fma(-ci[t], bi[(long long)t * n], acc));
            a.L[(long long)e * n + el] = acc;
        } else {
            const double* dr = a.dc_re + a.eg_doff[e] + p; const double* di = a.dc_im + a.eg_doff[e] + p;
            for (int t = 0; t < nc; ++t) acc = fma(dr[(long long)t * np_], br[(long long)t * n], fma(-di[(long long)t * np_], bi[(long long)t * n], acc));
            a.dL[row * n + el] = acc;
        }
    }
}

// stage 2: one thread per (e, p) row: dE_{e,p} (and E_e from the extra row of each generator)
__global__ void k_lind_expm(LindDev a, long long n_rows) {
    const int n = a.d * a.d;
    const long long npar_total = a.eg_poff[a.n_eg];
    const long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double* w = a.work + row * 7 * n;
    if (row < npar_total) {
        int e = 0; while (row >= a.eg_poff[e + 1]) ++e;
        // (this thread's own copy of E goes to the head of its work area; the E that is kept comes from the value row below)
        lb_expm_frechet(a.d, a.L + (long long)e * n, a.dL + row * n, w, a.dE + row * n, w + n);
    } else {
        const int e = (int)(row - npar_total);
        lb_expm_frechet(a.d, a.L + (long long)e * n, nullptr, a.E + (long long)e * n, nullptr, w);
    }
}

// stage 3: composition with the static part; one thread per output element of val / dval
//   vptr / dptr [n_mem + 1]: prefix sums of the members' sizes (d*d or d) and of size * n_par(generator of the member);
//   dval of member m is [size][n_par] row-major, as `deriv_wrt_params` returns it
__global__ void k_lind_compose(LindDev a, long long n_out_val, long long n_out_dval, const long long* __restrict__ vptr,
                               const long long* __restrict__ dptr) {
    const int d = a.d, n = d * d;
    const long long total = n_out_val + n_out_dval;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const bool is_val = idx < n_out_val;
        const long long o = is_val ? idx : idx - n_out_val;
        const long long* ptr = is_val ? vptr : dptr;
        int m = 0; while (o >= ptr[m + 1]) ++m;
        const long long local = o - ptr[m];
        const int e = a.m_eg[m], kind = a.m_kind[m];
        const int np_ = a.eg_npar[e];
        const double* S = a.stat + a.m_soff[m];
        int w_el; const double* X;                       // element of the member, matrix exp(L) or dE_p
        if (is_val) { w_el = (int)local; X = a.E + (long long)e * n; }
        else { w_el = (int)(local / np_); const long long p = local - (long long)w_el * np_; X = a.dE + (a.eg_poff[e] + p) * n; }
        double acc = 0.0;
        if (kind == 0) {                                 // op: (X . T)[i][j]
            const int i = w_el / d, j = w_el - i * d;
            for (int k = 0; k < d; ++k) acc = fma(X[i * d + k], S[k * d + j], acc);
        } else if (kind == 1) {                          // state: (X . rho0)[i]
            for (int k = 0; k < d; ++k) acc = fma(X[w_el * d + k], S[k], acc);
        } else {                                         // effect: (X^T . e0)[i]
            for (int k = 0; k < d; ++k) acc = fma(X[k * d + w_el], S[k], acc);
        }
        if (is_val) a.val[o] = acc; else a.dval[o] = acc;
    }
}
