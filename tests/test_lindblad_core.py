"""Host validation of pygsti_b200/csrc/lindblad_core.h -- the per-thread arithmetic of the next on-device row (exp of an error
generator + Frechet derivative) -- compiled with g++: against scipy on random matrices and against the reference's own
CPTPLND gate through oracle/oracle_lindblad.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.linalg as la

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lbc") / "liblbc.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "lindblad_core_check.cpp")], check=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.lbc_expm_frechet.argtypes = [C.c_int, dp, C.c_int, dp, dp, dp]
    lib.lbc_errorgen.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _expm_frechet(core, L, dL):
    d = L.shape[0]
    L = np.ascontiguousarray(L, float); dL = np.ascontiguousarray(dL, float)
    E = np.empty((d, d)); dE = np.empty((dL.shape[0], d, d))
    core.lbc_expm_frechet(d, _p(L), dL.shape[0], _p(dL), _p(E), _p(dE))
    return E, dE


@pytest.mark.parametrize("d", [4, 16])
@pytest.mark.parametrize("norm", [1e-3, 0.3, 2.0, 12.0])
def test_expm_frechet_vs_scipy(core, d, norm):
    rng = np.random.default_rng(d + int(norm * 1000))
    L = rng.standard_normal((d, d)); L *= norm / np.linalg.norm(L, 1)
    dL = rng.standard_normal((3, d, d))
    E, dE = _expm_frechet(core, L, dL)
    E_ref = la.expm(L)
    scale = max(1.0, np.max(np.abs(E_ref)))
    assert np.max(np.abs(E - E_ref)) <= 1e-12 * scale
    for p in range(3):
        dE_ref = la.expm_frechet(L, dL[p], compute_expm=False)
        assert np.max(np.abs(dE[p] - dE_ref)) <= 1e-11 * max(1.0, np.max(np.abs(dE_ref)))


def test_against_reference_cptplnd_gate(core):
    ref = os.path.join(REPO, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "pygsti")) and ref not in sys.path:
        sys.path.insert(0, ref)
    pytest.importorskip("pygsti", reason="reference install (baseline/_ref) not present")
    from pygsti.modelpacks import smq2Q_XYCNOT
    from oracle import oracle_lindblad as ol
    model = smq2Q_XYCNOT.target_model("CPTPLND")
    rng = np.random.default_rng(3)
    model.from_vector(model.to_vector() + 2e-2 * rng.standard_normal(model.num_params))
    op = model.operations[list(model.operations.keys())[-1]]          # the CNOT: ComposedOp([static target, ExpErrorgenOp])
    target = op.factorops[0].to_dense("HilbertSchmidt")
    B, c, dc = ol.lindblad_inputs(op.factorops[1].errorgen)
    d = target.shape[0]
    # error generator through the C core (real / imaginary parts separately)
    L = np.empty((d, d))
    core.lbc_errorgen(d, c.size, _p(np.ascontiguousarray(c.real)), _p(np.ascontiguousarray(c.imag)),
                      _p(np.ascontiguousarray(B.real)), _p(np.ascontiguousarray(B.imag)), _p(L))
    assert np.max(np.abs(L - op.factorops[1].errorgen.to_dense("HilbertSchmidt"))) <= 1e-13
    dL = ol.errorgen_derivs(dc, B)
    E, dE = _expm_frechet(core, L, dL)
    G = E @ target
    dG = np.einsum('pij,jk->pik', dE, target).reshape(dL.shape[0], -1).T
    assert np.max(np.abs(G - op.to_dense("HilbertSchmidt"))) <= 1e-12
    assert np.max(np.abs(dG - op.deriv_wrt_params())) <= 1e-10
