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

// ---- d = 16: the same recursion with one WARP per row and the 16 x 16 products on the FP64 tensor cores --------------------------------
// k_lind_expm above (one THREAD per (generator, parameter), ~65 serial 16 x 16 products out of a private global work area) took 41.6 ms
// for BASELINE config 4's 7 generators x 240 parameters.  Everything that does not depend on the parameter is computed once per
// generator by k_lind_gen16: the scaled generator X, the Taylor terms T_k = X^k / k! (k = 0..18), E and its squarings E_q.  k_lind_dexp16
// (one warp per (generator, parameter)) then runs  dT_k = (T_{k-1} dX + dT_{k-1} X) / k,  dE = sum_k dT_k,  and  dE <- E_q dE + dE E_q
// per squaring: two 16 x 16 x 16 products per step as 2 x 16 DMMA m8n8k4, X and dX held as B fragments in registers, T_{k-1} / E_q read
// as fragments from global memory (shared by the 240 warps of a generator through L1 / L2), dT / dE staged in shared memory (row
// stride 20: conflict-free fragment loads).  Same arithmetic as lb_expm_frechet (host-validated), different summation order.
#define LB16_LD 20
struct Lind16Work { double* T; double* Eq; double* X; int* s; };     // [n_eg][19][256], [n_eg][61][256], [n_eg][256], [n_eg]
__device__ __forceinline__ Lind16Work lind16_work(const LindDev& a) {
    Lind16Work w; const size_t n = 256;
    w.T = a.work; w.Eq = a.work + (size_t)a.n_eg * 19 * n; w.X = a.work + (size_t)a.n_eg * 80 * n;
    w.s = reinterpret_cast<int*>(a.work + (size_t)a.n_eg * 81 * n);
    return w;
}
// c[mt][nt] += A . B for 16 x 16 matrices; A row-major with leading dimension lda, B given as fragments bf[kk][nt]
__device__ __forceinline__ void lb16_mma_ab(const double* A, int lda, const double (&bf)[4][2], double2 (&c)[2][2], unsigned lg, unsigned lt) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const double a0 = A[lg * lda + 4 * kk + lt], a1 = A[(8 + lg) * lda + 4 * kk + lt];
        dmma884(c[0][0].x, c[0][0].y, a0, bf[kk][0]); dmma884(c[0][1].x, c[0][1].y, a0, bf[kk][1]);
        dmma884(c[1][0].x, c[1][0].y, a1, bf[kk][0]); dmma884(c[1][1].x, c[1][1].y, a1, bf[kk][1]);
    }
}
// B fragments of a row-major matrix: bf[kk][nt] = B[4 kk + lt][8 nt + lg]
__device__ __forceinline__ void lb16_bfrag(const double* B, int ldb, double scale, double (&bf)[4][2], unsigned lg, unsigned lt) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { bf[kk][0] = B[(4 * kk + lt) * ldb + lg] * scale; bf[kk][1] = B[(4 * kk + lt) * ldb + 8 + lg] * scale; }
}
// C-layout registers -> row-major matrix
__device__ __forceinline__ void lb16_store(double* M, int ld, const double2 (&c)[2][2], unsigned lg, unsigned lt) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) *reinterpret_cast<double2*>(M + (8 * mt + lg) * ld + 8 * nt + 2 * lt) = c[mt][nt];
}

// one warp per generator
__global__ void __launch_bounds__(32) k_lind_gen16(LindDev a) {
    __shared__ __align__(16) double sm[2][16 * LB16_LD];
    const int e = blockIdx.x, lane = threadIdx.x;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    const Lind16Work w = lind16_work(a);
    const double* L = a.L + (size_t)e * 256;
    // ||L||_1 = max column sum
    double cs = 0.0;
    if (lane < 16) for (int i = 0; i < 16; ++i) cs += fabs(L[i * 16 + lane]);
#pragma unroll
    for (int mk = 16; mk > 0; mk >>= 1) cs = fmax(cs, __shfl_xor_sync(0xffffffffu, cs, mk));
    int s = 0; double scale = 1.0;
    while (cs * scale > 0.5 && s < 60) { scale *= 0.5; ++s; }
    double xb[4][2];
    lb16_bfrag(L, 16, scale, xb, lg, lt);
    double* X = w.X + (size_t)e * 256; double* T = w.T + (size_t)e * 19 * 256; double* Eq = w.Eq + (size_t)e * 61 * 256;
    for (int i = lane; i < 256; i += 32) { X[i] = L[i] * scale; const double id = ((i >> 4) == (i & 15)) ? 1.0 : 0.0; T[i] = id; sm[0][(i >> 4) * LB16_LD + (i & 15)] = id; }
    if (lane == 0) w.s[e] = s;
    double2 E[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { E[mt][nt].x = (8 * mt + lg == 8 * nt + 2 * lt) ? 1.0 : 0.0; E[mt][nt].y = (8 * mt + lg == 8 * nt + 2 * lt + 1) ? 1.0 : 0.0; }
    __syncwarp();
    int cur = 0;
    for (int k = 1; k <= LB_TAYLOR_ORDER; ++k) {               // T_k = T_{k-1} X / k
        double2 c[2][2] = {{make_double2(0.0, 0.0), make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0), make_double2(0.0, 0.0)}};
        lb16_mma_ab(sm[cur], LB16_LD, xb, c, lg, lt);
        const double inv = 1.0 / (double)k;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { c[mt][nt].x *= inv; c[mt][nt].y *= inv; E[mt][nt].x += c[mt][nt].x; E[mt][nt].y += c[mt][nt].y; }
        lb16_store(sm[cur ^ 1], LB16_LD, c, lg, lt);
        lb16_store(T + (size_t)k * 256, 16, c, lg, lt);
        __syncwarp();
        cur ^= 1;
    }
    lb16_store(Eq, 16, E, lg, lt);                              // E_0 = Taylor sum; E_{q+1} = E_q E_q
    lb16_store(sm[0], LB16_LD, E, lg, lt);
    __syncwarp();
    cur = 0;
    for (int q = 0; q < s; ++q) {
        double eb[4][2];
        lb16_bfrag(sm[cur], LB16_LD, 1.0, eb, lg, lt);
        double2 c[2][2] = {{make_double2(0.0, 0.0), make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0), make_double2(0.0, 0.0)}};
        lb16_mma_ab(sm[cur], LB16_LD, eb, c, lg, lt);
        lb16_store(sm[cur ^ 1], LB16_LD, c, lg, lt);
        lb16_store(Eq + (size_t)(q + 1) * 256, 16, c, lg, lt);
        __syncwarp();
        cur ^= 1;
    }
    for (int i = lane; i < 256; i += 32) a.E[(size_t)e * 256 + i] = sm[cur][(i >> 4) * LB16_LD + (i & 15)];
}

// one warp per (generator, parameter) row
#define LB16_WARPS 4
__global__ void __launch_bounds__(LB16_WARPS * 32) k_lind_dexp16(LindDev a, long long n_par_rows) {
    __shared__ __align__(16) double sm[LB16_WARPS][16 * LB16_LD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lg = (unsigned)lane >> 2, lt = (unsigned)lane & 3u;
    const long long row = (long long)blockIdx.x * LB16_WARPS + warp;
    if (row >= n_par_rows) return;
    int e = 0; while (row >= a.eg_poff[e + 1]) ++e;
    const Lind16Work w = lind16_work(a);
    const int s = w.s[e];
    const double scale = scalbn(1.0, -s);
    const double* T = w.T + (size_t)e * 19 * 256; const double* Eq = w.Eq + (size_t)e * 61 * 256;
    double xb[4][2], dxb[4][2];
    lb16_bfrag(w.X + (size_t)e * 256, 16, 1.0, xb, lg, lt);
    lb16_bfrag(a.dL + row * 256, 16, scale, dxb, lg, lt);
    double* dT = sm[warp];
    for (int i = lane; i < 16 * LB16_LD; i += 32) dT[i] = 0.0;
    double2 dE[2][2] = {{make_double2(0.0, 0.0), make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0), make_double2(0.0, 0.0)}};
    __syncwarp();
    for (int k = 1; k <= LB_TAYLOR_ORDER; ++k) {               // dT_k = (T_{k-1} dX + dT_{k-1} X) / k
        double2 c[2][2] = {{make_double2(0.0, 0.0), make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0), make_double2(0.0, 0.0)}};
        lb16_mma_ab(T + (size_t)(k - 1) * 256, 16, dxb, c, lg, lt);
        lb16_mma_ab(dT, LB16_LD, xb, c, lg, lt);
        const double inv = 1.0 / (double)k;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { c[mt][nt].x *= inv; c[mt][nt].y *= inv; dE[mt][nt].x += c[mt][nt].x; dE[mt][nt].y += c[mt][nt].y; }
        __syncwarp();                                           // every lane has read dT_{k-1}
        lb16_store(dT, LB16_LD, c, lg, lt);
        __syncwarp();
    }
    for (int q = 0; q < s; ++q) {                               // dE <- E_q dE + dE E_q
        lb16_store(dT, LB16_LD, dE, lg, lt);
        __syncwarp();
        double db[4][2], eb[4][2];
        lb16_bfrag(dT, LB16_LD, 1.0, db, lg, lt);
        lb16_bfrag(Eq + (size_t)q * 256, 16, 1.0, eb, lg, lt);
        double2 c[2][2] = {{make_double2(0.0, 0.0), make_double2(0.0, 0.0)}, {make_double2(0.0, 0.0), make_double2(0.0, 0.0)}};
        lb16_mma_ab(Eq + (size_t)q * 256, 16, db, c, lg, lt);
        lb16_mma_ab(dT, LB16_LD, eb, c, lg, lt);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dE[mt][nt] = c[mt][nt];
        __syncwarp();
    }
    lb16_store(a.dE + row * 256, 16, dE, lg, lt);
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
