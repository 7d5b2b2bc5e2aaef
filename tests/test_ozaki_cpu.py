"""The arithmetic of the tcgen05 J^T J (csrc/kernels_ozaki.cuh), restated in numpy and checked on the CPU: the digit extraction of
k_oz_slice (magic-number rounding, exact remainders), the digit range the int8 operands rely on, the exactness bound of the int32
level sums and the accuracy of the recombination against a long-double product.  (The CUDA kernels themselves are tested on the GPU:
tests/test_gpu_parity.py::test_scaled_jacobian_and_jtj, test_jtj_tcgen05_full_size_vs_fp64.)"""
import numpy as np

T = 8


def slice_digits(J):
    """digits[t] (int8) and column exponents e with  J[:, p] 2^-e_p = sum_t digits[t][:, p] 2^(-6-7t)  up to 2^-56."""
    m = np.max(np.abs(J), axis=0)
    e = np.where(m > 0, np.frexp(m)[1], 0)
    y = J * np.exp2(6.0 - e)[None, :]                       # |y| < 64
    digs = []
    for t in range(T):
        magic = 6755399441055744.0 / float(1 << (7 * t))    # 1.5 2^(52 - 7t)
        mm = y + magic                                       # rounds y to a multiple of 2^-7t (round-to-nearest-even)
        lo = mm.view(np.int64) & 0xFF                        # low mantissa byte = the digit in two's complement
        d = np.where(lo >= 128, lo - 256, lo).astype(np.int64)
        y = y - (mm - magic)                                 # exact
        digs.append(d)
    return digs, e, y


def test_digits_are_int8_and_exact():
    rng = np.random.default_rng(0)
    J = rng.standard_normal((500, 37)) * 10.0 ** rng.integers(-8, 8, size=37)[None, :]
    J[:, 3] = 0.0; J[7, 5] = np.max(np.abs(J[:, 5])) * 0.999999; J[::3, 9] *= 1e-12
    digs, e, rem = slice_digits(J)
    for t, d in enumerate(digs):
        assert d.min() >= -64 and d.max() <= 64, t                                      # what the int8 operands assume
    rec = sum(d.astype(np.float64) * 2.0 ** (-6 - 7 * t) for t, d in enumerate(digs))   # exact in float64? every term is: sum it in long double
    recl = sum(d.astype(np.longdouble) * np.longdouble(2.0) ** (-6 - 7 * t) for t, d in enumerate(digs))
    scaled = (J * np.exp2(-e.astype(np.float64))[None, :]).astype(np.longdouble)
    assert np.max(np.abs(recl - scaled)) <= 2.0 ** -56                                   # half a unit of the last digit
    assert np.max(np.abs(rem)) <= 2.0 ** -50 * 2.0 ** 6 and rec.shape == J.shape


def test_level_sums_fit_int32_and_recombine_to_fp64_accuracy():
    rng = np.random.default_rng(1)
    nE, Np = 3000, 24
    J = rng.standard_normal((nE, Np)) * 10.0 ** rng.integers(-3, 4, size=Np)[None, :]
    digs, e, _ = slice_digits(J)
    # the bound the host code uses for a K slice: 8 pairs x 65 504 rows x 64^2 < 2^31
    assert 8 * 65504 * 64 * 64 < 2 ** 31
    C = np.zeros((Np, Np), dtype=np.longdouble)
    for t in range(T):
        lvl = np.zeros((Np, Np), dtype=np.int64)
        for a in range(t + 1):
            lvl += digs[a].T @ digs[t - a]                  # the exact integer GEMMs of level t = a + b
        assert np.max(np.abs(lvl)) < 2 ** 31
        C += lvl.astype(np.longdouble) * np.longdouble(2.0) ** (-12 - 7 * t)
    C = (C * (np.longdouble(2.0) ** (e[:, None] + e[None, :]))).astype(np.float64)
    ref = (J.astype(np.longdouble).T @ J.astype(np.longdouble)).astype(np.float64)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))   # |C_ij| <= scale_ij
    assert np.max(np.abs(C - ref) / scale) <= 1e-13         # dropped pairs a + b >= 8: 2^-62 x (column maxima product) x nE
    assert np.array_equal(C, C.T)
