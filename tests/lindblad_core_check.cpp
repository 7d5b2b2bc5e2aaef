// C entry points around pygsti_b200/csrc/lindblad_core.h for tests/test_lindblad_core.py (host build with g++; no GPU).
#include <vector>
#include "../pygsti_b200/csrc/lindblad_core.h"

extern "C" {
// batch of directions for one generator: E [d*d], dE [n_dir][d*d]
void lbc_expm_frechet(int d, const double* L, int n_dir, const double* dL, double* E, double* dE) {
    std::vector<double> work((size_t)6 * d * d), Etmp((size_t)d * d);
    if (n_dir == 0) { lb_expm_frechet(d, L, nullptr, E, nullptr, work.data()); return; }
    for (int p = 0; p < n_dir; ++p)
        lb_expm_frechet(d, L, dL + (size_t)p * d * d, p == 0 ? E : Etmp.data(), dE + (size_t)p * d * d, work.data());
}
void lbc_errorgen(int d, int n_coeff, const double* c_re, const double* c_im, const double* B_re, const double* B_im, double* L) {
    lb_errorgen(d, n_coeff, c_re, c_im, B_re, B_im, L);
}
}
