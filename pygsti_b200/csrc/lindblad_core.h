// lindblad_core.h -- arithmetic core of the Lindblad row (second part): the dense exponential of an
// error generator and its Frechet derivative, as plain per-thread C++ that compiles for the host and for the device.
//
// Reference: ExpErrorgenOp._update_rep (scipy.linalg.expm, pygsti/modelmembers/operations/experrorgenop.py:114-125) and
// ExpErrorgenOp.deriv_wrt_params (`_d_exp_x`, commutator series, :213-262, :722-830).  Here (E, dE) = (exp(L), Dexp(L)[dL]) come from ONE
// scaling-and-squaring recursion on the pair (L, dL): the pair is the block upper-triangular matrix [[L, dL], [0, L]], whose
// exponential is [[E, dE], [0, E]], and pairs multiply as (A, D)(B, F) = (AB, AF + DB) -- three d x d products instead of the eight of a
// 2d x 2d product.  Scaling: X = L / 2^s with ||X||_1 <= 1/2; Taylor to order LB_TAYLOR_ORDER (remainder < 1e-18 relative); s squarings.
//
// Validation: on the host (tests/test_lindblad_core.py: against scipy.linalg.expm / expm_frechet on random matrices with 1-norms up to
// 12, and against the reference's own CPTPLND gate through oracle/oracle_lindblad.py) and on the device through
// csrc/kernels_lindblad.cuh (b200_lindblad_members; tests/test_gpu_synthetic.py, tests/test_gpu_pygsti_dropin.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

#define LB_TAYLOR_ORDER 18

// C = A . B   (d x d, row-major; C must not alias A or B)
LB_HD void lb_matmul(int d, const double* A, const double* B, double* C) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double acc = 0.0;
            for (int k = 0; k < d; ++k) acc = fma(A[i * d + k], B[k * d + j], acc);
            C[i * d + j] = acc;
        }
}
// C += A . B
LB_HD void lb_matmul_acc(int d, const double* A, const double* B, double* C) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double acc = C[i * d + j];
            for (int k = 0; k < d; ++k) acc = fma(A[i * d + k], B[k * d + j], acc);
            C[i * d + j] = acc;
        }
}

// E = exp(L) and, if dL != nullptr, dE = Dexp(L)[dL].   work: 6 d^2 doubles (4 d^2 suffice when dL == nullptr).
LB_HD void lb_expm_frechet(int d, const double* L, const double* dL, double* E, double* dE, double* work) {
    const int n = d * d;
    double* X = work;            // scaled generator
    double* T = work + n;        // current Taylor term (value part)
    double* Tn = work + 2 * n;   // scratch for products
    double* dX = work + 3 * n;   // scaled direction
    double* dT = work + 4 * n;   // current Taylor term (derivative part)
    double* dTn = work + 5 * n;
    // ||L||_1 = max column sum
    double nrm = 0.0;
    for (int j = 0; j < d; ++j) {
        double cs = 0.0;
        for (int i = 0; i < d; ++i) cs += fabs(L[i * d + j]);
        nrm = cs > nrm ? cs : nrm;
    }
    int s = 0;
    double scale = 1.0;
    while (nrm * scale > 0.5 && s < 60) { scale *= 0.5; ++s; }
    for (int i = 0; i < n; ++i) { X[i] = L[i] * scale; T[i] = 0.0; E[i] = 0.0; }
    for (int i = 0; i < d; ++i) { T[i * d + i] = 1.0; E[i * d + i] = 1.0; }
    if (dL) for (int i = 0; i < n; ++i) { dX[i] = dL[i] * scale; dT[i] = 0.0; dE[i] = 0.0; }
    // Taylor: term_k = term_{k-1} . (X, dX) / k
    for (int k = 1; k <= LB_TAYLOR_ORDER; ++k) {
        const double inv = 1.0 / (double)k;
        if (dL) {
            lb_matmul(d, T, dX, dTn);            // T_{k-1} dX
            lb_matmul_acc(d, dT, X, dTn);        // + dT_{k-1} X
            for (int i = 0; i < n; ++i) { dT[i] = dTn[i] * inv; dE[i] += dT[i]; }
        }
        lb_matmul(d, T, X, Tn);
        for (int i = 0; i < n; ++i) { T[i] = Tn[i] * inv; E[i] += T[i]; }
    }
    // squaring: (E, dE) <- (E E, E dE + dE E)
    for (int q = 0; q < s; ++q) {
        if (dL) {
            lb_matmul(d, E, dE, dTn);
            lb_matmul_acc(d, dE, E, dTn);
            for (int i = 0; i < n; ++i) dE[i] = dTn[i];
        }
        lb_matmul(d, E, E, Tn);
        for (int i = 0; i < n; ++i) E[i] = Tn[i];
    }
}

// L = Re sum_i c_i B_i for complex coefficients / term superoperators given as separate real and imaginary parts
// (LindbladErrorgen._update_rep, dense branch, lindbladerrorgen.py:700-708):  Re(c B) = c_re B_re - c_im B_im
LB_HD void lb_errorgen(int d, int n_coeff, const double* c_re, const double* c_im, const double* B_re, const double* B_im, double* L) {
    const int n = d * d;
    for (int i = 0; i < n; ++i) L[i] = 0.0;
    for (int t = 0; t < n_coeff; ++t) {
        const double cr = c_re[t], ci = c_im[t];
        const double* br = B_re + (long long)t * n;
        const double* bi = B_im + (long long)t * n;
        for (int i = 0; i < n; ++i) L[i] = fma(cr, br[i], fma(-ci, bi[i], L[i]));
    }
}
