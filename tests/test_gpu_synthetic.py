"""Edge cases on the GPU against the oracle, with synthetic layouts (no pyGSTi): empty circuits, duplicate circuits,
cache links, outcome subsets, more than 4 / fewer than 4 effects, more than 8 gates, several preps, every kernel
family (d = 4, 16, 64, 256), permuted and general derivative maps, every d = 16 code path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_np as onp
from tests import synth

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(gpu_ctx, t, G, rho, E, D):
    at = gpu_ctx.upload_atom(t)
    at.set_model(G, rho, E)
    at.set_derivs(D)
    p = np.full(t.n_elements, np.nan)
    at.fill_probs(p)
    J = np.full((t.n_elements, D.n_params), np.nan)
    p2 = np.full(t.n_elements, np.nan)
    at.fill_dprobs(J, p2)
    info = at.info()
    at.free()
    return p, p2, J, info


@pytest.mark.parametrize("dim,n_ops,n_rho,n_eff,depth", [(16, 5, 1, 4, 40), (16, 11, 2, 6, 30), (16, 3, 2, 2, 70),
                                                        (16, 20, 1, 8, 25), (4, 3, 1, 2, 50), (64, 4, 1, 8, 12)])
def test_fused_path_edge_cases(gpu_ctx, dim, n_ops, n_rho, n_eff, depth):
    circs = synth.random_circuits(40 if dim < 64 else 14, depth, n_ops, n_rho, n_eff, seed=dim + n_ops)
    t = synth.make_tables(dim, n_ops, n_rho, n_eff, circs)
    G, rho, E = synth.random_model(dim, n_ops, n_rho, n_eff, seed=1)
    D = synth.full_derivs(t)
    p, p2, J, info = _run(gpu_ctx, t, G, rho, E, D)
    po = onp.mapfill_probs(t, G, rho, E)
    Jo = onp.dprobs_analytic(t, G, rho, E, D)
    scale = max(1.0, np.max(np.abs(Jo)))
    assert np.max(np.abs(p - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    assert np.max(np.abs(p2 - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    assert np.max(np.abs(J - Jo)) <= 1e-11 * scale
    assert info["fused_path"] == 1


def test_permuted_and_partial_column_map(gpu_ctx):
    """unit partial permutation that is NOT contiguous: shuffled columns, some member elements without a parameter,
    some parameters fed by nothing (must come out exactly zero)."""
    from pygsti_b200.packing import DerivMap
    circs = synth.random_circuits(30, 30, 5, 1, 4, seed=3)
    t = synth.make_tables(16, 5, 1, 4, circs)
    G, rho, E = synth.random_model(16, 5, 1, 4, seed=2)
    n_w = 5 * 256 + 16 + 64
    rng = np.random.default_rng(0)
    keep = np.sort(rng.choice(n_w, size=n_w - 200, replace=False)).astype(np.int32)
    n_params = n_w + 37
    cols = rng.permutation(n_params)[:keep.size].astype(np.int32)
    D = DerivMap(n_w, n_params, keep, cols, np.ones(keep.size))
    p, p2, J, info = _run(gpu_ctx, t, G, rho, E, D)
    Jo = onp.dprobs_analytic(t, G, rho, E, D)
    assert info["fused_path"] == 1
    assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    unfed = np.setdiff1d(np.arange(n_params), cols)
    assert np.all(J[:, unfed] == 0.0)


@pytest.mark.parametrize("dim,n_ops,n_eff", [(16, 6, 4), (16, 9, 5), (4, 3, 2), (64, 3, 8)])
def test_general_derivative_map(gpu_ctx, dim, n_ops, n_eff):
    circs = synth.random_circuits(24 if dim < 64 else 10, 20 if dim < 64 else 8, n_ops, 2, n_eff, seed=7)
    t = synth.make_tables(dim, n_ops, 2, n_eff, circs)
    G, rho, E = synth.random_model(dim, n_ops, 2, n_eff, seed=5)
    D = synth.random_derivs(t, 57, density=0.03, seed=1)
    p, p2, J, info = _run(gpu_ctx, t, G, rho, E, D)
    Jo = onp.dprobs_analytic(t, G, rho, E, D)
    assert info["fused_path"] == 0
    assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))


def test_probs_d256(gpu_ctx):
    circs = synth.random_circuits(10, 6, 3, 1, 16, seed=11, subsets=False)
    t = synth.make_tables(256, 3, 1, 16, circs)
    G, rho, E = synth.random_model(256, 3, 1, 16, seed=4)
    at = gpu_ctx.upload_atom(t); at.set_model(G, rho, E)
    p = np.full(t.n_elements, np.nan); at.fill_probs(p); at.free()
    po = onp.mapfill_probs(t, G, rho, E)
    assert np.max(np.abs(p - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))


@pytest.mark.parametrize("mode", ["trie", "generic"])
def test_every_d16_code_path(mode):
    """The two d = 16 Jacobian implementations -- the trie kernels (default) and the generic W . D kernels that take over
    for layouts the trie path rejects (B200_D16_MODE=generic forces them) -- give the same answers; run in a subprocess
    because the knob is latched at first use."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from pygsti_b200 import engine
from oracle import oracle_np as onp
from tests import synth
ctx = engine.Context(0)
for (n_ops, n_rho, n_eff, depth, seed) in [(5, 1, 4, 60, 0), (7, 2, 3, 33, 1)]:
    circs = synth.random_circuits(50, depth, n_ops, n_rho, n_eff, seed=seed)
    t = synth.make_tables(16, n_ops, n_rho, n_eff, circs)
    G, rho, E = synth.random_model(16, n_ops, n_rho, n_eff, seed=seed)
    D = synth.full_derivs(t)
    at = ctx.upload_atom(t); at.set_model(G, rho, E); at.set_derivs(D)
    J = np.full((t.n_elements, D.n_params), np.nan); p = np.full(t.n_elements, np.nan)
    at.fill_dprobs(J, p)
    Jo = onp.dprobs_analytic(t, G, rho, E, D); po = onp.mapfill_probs(t, G, rho, E)
    assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo))), np.max(np.abs(J - Jo))
    assert np.max(np.abs(p - po)) <= 1e-12
print("ok")
''' % REPO
    mode, _, knobs = mode.partition(":")
    env = dict(os.environ, B200_D16_MODE=mode)
    for kv in filter(None, knobs.split(",")):
        k, v = kv.split("="); env[k] = v
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def _shared_param_derivs(t, n_params, seed=0):
    """Derivative map in the style of a crosstalk-free model: every parameter touches the SAME (i, j) positions of
    several gates (one 1-qubit gate embedded on different qubits), plus TP-POVM-like rows (one parameter feeding two
    effects with opposite signs) and a state-preparation block."""
    from pygsti_b200.packing import DerivMap
    d = t.dim
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for p in range(n_params - 2 * d):
        gates = rng.choice(t.n_ops, size=min(t.n_ops, 1 + p % 3), replace=False)
        ij = rng.integers(0, d * d, size=int(rng.integers(1, 9)))
        for g in gates:
            for w in ij:
                rows.append(g * d * d + int(w)); cols.append(p); vals.append(float(rng.standard_normal()))
    off_rho = t.n_ops * d * d
    off_eff = off_rho + t.n_rho * d
    for i in range(d):                                   # prep parameters (all preps share them)
        p = n_params - 2 * d + i
        for r in range(t.n_rho):
            rows.append(off_rho + r * d + i); cols.append(p); vals.append(1.0 + r)
    for i in range(d):                                   # effect parameters: +1 on effect 0, -1 on the last effect
        p = n_params - d + i
        rows.append(off_eff + i); cols.append(p); vals.append(1.0)
        rows.append(off_eff + (t.n_eff - 1) * d + i); cols.append(p); vals.append(-1.0)
    return DerivMap(off_eff + t.n_eff * d, n_params, np.asarray(rows, np.int32), np.asarray(cols, np.int32), np.asarray(vals))


@pytest.mark.parametrize("dim,n_ops,n_rho,n_eff,n_circ,depth,n_params", [(64, 6, 2, 8, 30, 70, 700), (64, 3, 1, 5, 12, 150, 300),
                                                                         (256, 3, 1, 16, 6, 9, 600)])
def test_level_batched_jacobian(gpu_ctx, dim, n_ops, n_rho, n_eff, n_circ, depth, n_params):
    """d = 64 / 256 Jacobian path (level-batched sweeps + sparse contraction): parameters shared between gates, several
    parameter tiles, gate buckets longer than one shared-memory stage, outcome subsets, row scaling, J^T J."""
    circs = synth.random_circuits(n_circ, depth, n_ops, n_rho, n_eff, seed=dim + depth)
    t = synth.make_tables(dim, n_ops, n_rho, n_eff, circs)
    G, rho, E = synth.random_model(dim, n_ops, n_rho, n_eff, seed=9)
    D = _shared_param_derivs(t, n_params, seed=2)
    p, p2, J, info = _run(gpu_ctx, t, G, rho, E, D)
    po = onp.mapfill_probs(t, G, rho, E)
    Jo = onp.dprobs_analytic(t, G, rho, E, D)
    sc = max(1.0, np.max(np.abs(Jo)))
    assert info["fused_path"] == 0
    assert np.max(np.abs(p2 - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    assert np.max(np.abs(J - Jo)) <= 1e-11 * sc
    # fused row scaling and J^T J / J^T f on the same path
    at = gpu_ctx.upload_atom(t); at.set_model(G, rho, E); at.set_derivs(D)
    rs = np.random.default_rng(1).standard_normal(t.n_elements)
    Js = np.full_like(J, np.nan)
    at.fill_dprobs(Js, row_scale=rs)
    assert np.max(np.abs(Js - Jo * rs[:, None])) <= 1e-11 * sc * max(1.0, np.max(np.abs(rs)))
    f = np.random.default_rng(2).standard_normal(t.n_elements)
    jtj, jtf = at.jtj(rs, f)
    Jr = Jo * rs[:, None]
    assert np.max(np.abs(jtj - Jr.T @ Jr)) <= 1e-9 * max(1.0, np.max(np.abs(Jr.T @ Jr)))
    assert np.max(np.abs(jtf - Jr.T @ f)) <= 1e-9 * max(1.0, np.max(np.abs(Jr.T @ f)))
    at.free()


def test_level_batched_matches_generic_path():
    """The level-batched d = 64 Jacobian (DMMA accumulate, and its scalar predecessor) and the correctness-first W-matrix
    path agree (subprocess: env latch)."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from pygsti_b200 import engine
from tests import synth
ctx = engine.Context(0)
circs = synth.random_circuits(16, 20, 4, 1, 8, seed=5)
t = synth.make_tables(64, 4, 1, 8, circs)
G, rho, E = synth.random_model(64, 4, 1, 8, seed=3)
D = synth.random_derivs(t, 90, density=0.01, seed=4)
at = ctx.upload_atom(t); at.set_model(G, rho, E); at.set_derivs(D)
J = np.full((t.n_elements, D.n_params), np.nan); at.fill_dprobs(J)
np.save(sys.argv[1], J); print("ok", ctx.launch_count)
''' % REPO
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for tag, extra in (("lj", {}), ("gen", {"B200_NO_LEVELJ": "1"}), ("lj_v1", {"B200_LJ_ACCUM_V1": "1"}), ("lj_v2", {"B200_LJ_ACCUM_V2": "1"})):
            f = os.path.join(td, tag + ".npy")
            r = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, **extra), capture_output=True, text=True, timeout=600)
            assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
            outs.append(np.load(f))
    assert np.all(np.isfinite(outs[0]))
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-11 * max(1.0, np.max(np.abs(outs[1])))   # DMMA accumulate vs W-matrix path
    assert np.max(np.abs(outs[2] - outs[1])) <= 1e-11 * max(1.0, np.max(np.abs(outs[1])))   # scalar accumulate vs W-matrix path
    assert np.max(np.abs(outs[3] - outs[1])) <= 1e-11 * max(1.0, np.max(np.abs(outs[1])))   # CTA-per-tile DMMA accumulate (v2)


def test_repeated_calls_with_changing_models(gpu_ctx):
    """The trie path resets only the table rows a chain waits for (k_trie_prepare): successive calls on one atom with
    DIFFERENT models, and a model repeated, must each match the oracle; the phase-timing events must not change results."""
    circs = synth.random_circuits(60, 45, 5, 2, 4, seed=11)
    t = synth.make_tables(16, 5, 2, 4, circs)
    D = synth.full_derivs(t)
    at = gpu_ctx.upload_atom(t)
    at.set_derivs(D)
    models = [synth.random_model(16, 5, 2, 4, seed=s) for s in (21, 22, 21, 23)]
    gpu_ctx.phase_timing(True)
    for k, (G, rho, E) in enumerate(models):
        at.set_model(G, rho, E)
        J = np.full((t.n_elements, D.n_params), np.nan)
        p = np.full(t.n_elements, np.nan)
        at.fill_dprobs(J, p)
        if k % 2 == 0:
            p1 = np.full(t.n_elements, np.nan)
            at.fill_probs(p1)
            assert np.max(np.abs(p1 - p)) <= 1e-13
        po = onp.mapfill_probs(t, G, rho, E)
        Jo = onp.dprobs_analytic(t, G, rho, E, D)
        assert np.max(np.abs(p - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
        assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    ms, n = gpu_ctx.phase_ms()
    gpu_ctx.phase_timing(False)
    assert n == len(models) and all(x > 0.0 for x in ms)
    at.free()


def test_affine_model_update_on_device(gpu_ctx):
    """b200_atom_bind_params / set_params: M = M_const + D theta on the device for a general affine map (several
    parameters per element, elements without parameters), against numpy; the Jacobian of the updated model against the
    oracle; binding is refused for a derivative map over a parameter sub-block that does not match theta."""
    from pygsti_b200.packing import DerivMap
    circs = synth.random_circuits(20, 25, 4, 1, 4, seed=4)
    t = synth.make_tables(16, 4, 1, 4, circs)
    n_w = 4 * 256 + 16 + 64
    rng = np.random.default_rng(3)
    n_params = 300
    nnz = 2000
    rows = rng.integers(0, n_w, size=nnz).astype(np.int32); cols = rng.integers(0, n_params, size=nnz).astype(np.int32)
    keep = np.unique(rows.astype(np.int64) * n_params + cols, return_index=True)[1]
    rows, cols = rows[keep], cols[keep]
    vals = rng.standard_normal(rows.size)
    D = DerivMap(n_w, n_params, rows, cols, vals)
    Dd = np.zeros((n_w, n_params)); Dd[rows, cols] = vals
    Mc = 0.1 * rng.standard_normal(n_w)
    th0 = 0.05 * rng.standard_normal(n_params)
    M0 = Mc + Dd @ th0
    split = lambda M: (M[:1024].reshape(4, 16, 16), M[1024:1040].reshape(1, 16), M[1040:].reshape(4, 16))
    at = gpu_ctx.upload_atom(t)
    at.set_model(*split(M0)); at.set_derivs(D)
    with pytest.raises(Exception):
        at.bind_params(th0[:-1])                      # wrong length: refused
    at.bind_params(th0)
    for k in range(3):
        th = 0.05 * rng.standard_normal(n_params)
        at.set_params(th)
        M = Mc + Dd @ th
        assert np.max(np.abs(at.get_model() - M)) <= 1e-14
        J = np.full((t.n_elements, n_params), np.nan); p = np.full(t.n_elements, np.nan)
        at.fill_dprobs(J, p)
        G, rho, E = split(M)
        assert np.max(np.abs(p - onp.mapfill_probs(t, G, rho, E))) <= 1e-12
        Jo = onp.dprobs_analytic(t, G, rho, E, D)
        assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    at.free()


@pytest.mark.parametrize("d", [4, 16])
def test_lindblad_members_on_device(gpu_ctx, d):
    """b200_lindblad_members (second part): L = Re sum c_i B_i, exp(L), its Frechet derivatives and the composition
    with the static parts, for random complex coefficients / term superoperators and two generators shared by several members,
    against the numpy oracle (oracle/oracle_lindblad.py, itself pinned against the reference's CPTPLND / H+S / GLND members)."""
    from types import SimpleNamespace as NS
    from oracle import oracle_lindblad as ol
    rng = np.random.default_rng(100 + d)
    errgens = []
    for n_coeff, n_par, scale in ((12, 9, 0.15), (5, 4, 0.6)):
        B = (rng.standard_normal((n_coeff, d, d)) + 1j * rng.standard_normal((n_coeff, d, d))) / d
        c = scale * (rng.standard_normal(n_coeff) + 1j * rng.standard_normal(n_coeff))
        dc = rng.standard_normal((n_coeff, n_par)) + 1j * rng.standard_normal((n_coeff, n_par))
        errgens.append(NS(B_re=np.ascontiguousarray(B.real), B_im=np.ascontiguousarray(B.imag), c=c, dc=dc, B=B))
    members = [NS(kind="op", errgen=0, static=rng.standard_normal((d, d))), NS(kind="rho", errgen=1, static=rng.standard_normal(d)),
               NS(kind="op", errgen=1, static=rng.standard_normal((d, d))), NS(kind="eff", errgen=0, static=rng.standard_normal(d)),
               NS(kind="eff", errgen=0, static=rng.standard_normal(d))]
    out = gpu_ctx.lindblad_members(d, errgens, members)
    assert len(out) == len(members)
    for m, (val, dval) in zip(members, out):
        e = errgens[m.errgen]
        if m.kind == "op":
            v_ref, dv_ref = ol.composed_gate(e.c, e.dc, e.B, m.static)
            v_ref = v_ref.ravel()
        elif m.kind == "rho":
            v_ref, dv_ref = ol.composed_state(e.c, e.dc, e.B, m.static)
        else:
            v_ref, dv_ref = ol.composed_effect(e.c, e.dc, e.B, m.static)
        assert val.shape == v_ref.shape and dval.shape == dv_ref.shape
        assert np.max(np.abs(val - v_ref)) <= 1e-12 * max(1.0, np.max(np.abs(v_ref)))
        assert np.max(np.abs(dval - dv_ref)) <= 1e-11 * max(1.0, np.max(np.abs(dv_ref)))


@pytest.mark.parametrize("dim,n_ops", [(256, 9), (64, 6)])
def test_factored_model_embedded_composed(gpu_ctx, dim, n_ops):
    """Gates as products of small operations embedded on 1-2 qubits (the device form of OpCRep_Composed / OpCRep_Embedded,
    opcreps.cpp:93-158, 242-276): probabilities through the factor programs equal the dense-product path and the oracle, and
    the dense matrices built on the device equal the Kronecker construction (packing.factored_to_dense)."""
    from pygsti_b200.packing import FactoredModel, factored_to_dense
    nq = {64: 3, 256: 4}[dim]
    rng = np.random.default_rng(dim)
    fptr, f_nq, f_t, f_off, mats, off = [0], [], [], [], [], 0
    for g in range(n_ops):
        for _ in range(int(rng.integers(0 if g == 0 else 1, 4))):          # gate 0 may be the empty product (identity)
            k = int(rng.integers(1, 3))
            tg = [int(x) for x in rng.choice(nq, size=k, replace=False)]     # any order, e.g. (2, 0)
            small = np.eye(4 ** k) * 0.8 + 0.3 * rng.standard_normal((4 ** k, 4 ** k)) / 2 ** k
            f_nq.append(k); f_t.append(tg + [-1] * (4 - k)); f_off.append(off); mats.append(small.ravel()); off += small.size
        fptr.append(len(f_nq))
    n_eff = 8
    rho = rng.standard_normal((1, dim)) / np.sqrt(dim); E = rng.standard_normal((n_eff, dim)) / np.sqrt(dim)
    fm = FactoredModel(n_qubits=nq, op_fptr=np.asarray(fptr, np.int32), f_nq=np.asarray(f_nq, np.int32),
                       f_targets=np.asarray(f_t, np.int32).reshape(-1, 4), f_moff=np.asarray(f_off, np.int64),
                       mats=np.concatenate(mats), rho=rho, E=E)
    G = factored_to_dense(fm, dim)
    circs = synth.random_circuits(40, 30, n_ops, 1, n_eff, seed=3)
    t = synth.make_tables(dim, n_ops, 1, n_eff, circs)
    at = gpu_ctx.upload_atom(t); at.set_model_factored(fm)
    p = np.full(t.n_elements, np.nan); at.fill_probs(p)
    Md = at.get_model()
    assert np.max(np.abs(Md[:n_ops * dim * dim].reshape(n_ops, dim, dim) - G)) <= 1e-13 * max(1.0, np.max(np.abs(G)))
    po = onp.mapfill_probs(t, G, rho, E)
    assert np.max(np.abs(p - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    at.set_model(G, rho, E)                                                  # the dense path on the same model
    pd = np.full(t.n_elements, np.nan); at.fill_probs(pd)
    assert np.max(np.abs(p - pd)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    # the derivative path on a factored model uses the device-built dense matrices
    D = synth.random_derivs(t, 40, density=0.0005, seed=5)
    at.set_model_factored(fm); at.set_derivs(D)
    J = np.full((t.n_elements, 40), np.nan); at.fill_dprobs(J)
    Jo = onp.dprobs_analytic(t, G, rho, E, D)
    assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    at.free()


def _random_factored(dim, n_ops, rng, share=False):
    from pygsti_b200.packing import FactoredModel
    nq = {64: 3, 256: 4}[dim]
    fptr, f_nq, f_t, f_off, mats, off = [0], [], [], [], [], 0
    for g in range(n_ops):
        for _ in range(int(rng.integers(0 if g == 0 else 1, 3))):
            k = int(rng.integers(1, 3))
            tg = [int(x) for x in rng.choice(nq, size=k, replace=False)]
            small = np.eye(4 ** k) * 0.8 + 0.3 * rng.standard_normal((4 ** k, 4 ** k)) / 2 ** k
            if share and f_nq and k == f_nq[0]:                 # this factor re-uses the first factor's matrix (one gate on another qubit)
                f_nq.append(k); f_t.append(tg + [-1] * (4 - k)); f_off.append(f_off[0])
                continue
            f_nq.append(k); f_t.append(tg + [-1] * (4 - k)); f_off.append(off); mats.append(small.ravel()); off += small.size
        fptr.append(len(f_nq))
    n_eff = 8
    rho = rng.standard_normal((2, dim)) / np.sqrt(dim); E = rng.standard_normal((n_eff, dim)) / np.sqrt(dim)
    return FactoredModel(n_qubits=nq, op_fptr=np.asarray(fptr, np.int32), f_nq=np.asarray(f_nq, np.int32),
                         f_targets=np.asarray(f_t, np.int32).reshape(-1, 4), f_moff=np.asarray(f_off, np.int64),
                         mats=np.concatenate(mats), rho=rho, E=E)


@pytest.mark.parametrize("dim,n_ops,share", [(64, 7, False), (64, 6, True), (256, 6, False)])
def test_factored_jacobian_matches_oracle(gpu_ctx, dim, n_ops, share):
    """The Jacobian straight from the factor programs (b200_atom_set_derivs_factored, kernels_factoredj.cuh: forward states per
    factor step, backward walk with DMMA chain + accumulate, CSC contraction over factor-space elements) against the numpy
    restatement (oracle_np.dprobs_factored) on random factor programs: 1- and 2-qubit factors on any qubits in any order, empty
    products, several factors per layer, shared matrices, a random sparse factor-space map; row scale; probabilities."""
    from pygsti_b200.packing import DerivMap
    rng = np.random.default_rng(dim + n_ops)
    fm = _random_factored(dim, n_ops, rng, share)
    n_eff = fm.E.shape[0]
    circs = synth.random_circuits(24, 20, n_ops, 2, n_eff, seed=4)
    t = synth.make_tables(dim, n_ops, 2, n_eff, circs)
    n_wf = fm.mats.size + (2 + n_eff) * dim
    Np = 150
    nnz = 4000
    Df = DerivMap(n_wf, Np, rng.integers(0, n_wf, nnz), rng.integers(0, Np, nnz), rng.standard_normal(nnz))
    at = gpu_ctx.upload_atom(t); at.set_model_factored(fm); at.set_derivs_factored(Df)
    J = np.full((t.n_elements, Np), np.nan); p = np.full(t.n_elements, np.nan)
    at.fill_dprobs(J, p)
    Jo = onp.dprobs_factored(t, fm, Df)
    from pygsti_b200.packing import factored_to_dense
    po = onp.mapfill_probs(t, factored_to_dense(fm, dim), fm.rho, fm.E)
    assert np.max(np.abs(p - po)) <= 1e-12 * max(1.0, np.max(np.abs(po)))
    assert np.max(np.abs(J - Jo)) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    w = rng.uniform(-2, 2, t.n_elements)
    Js = np.full((t.n_elements, Np), np.nan); at.fill_dprobs(Js, row_scale=w)
    assert np.max(np.abs(Js - Jo * w[:, None])) <= 1e-11 * max(1.0, np.max(np.abs(Jo)))
    at.free()


def test_jtj_tcgen05_nonfinite_column(gpu_ctx):
    """A Jacobian column that holds NaN / Inf cannot be cut into int8 digits: the tcgen05 J^T J must flag it (NaN in that row and
    column of the result, as an FP64 contraction delivers) and leave every other entry exact."""
    dim, n_ops, n_eff = 4, 3, 2
    G, rho, E = synth.random_model(dim, n_ops, 1, n_eff, seed=2)
    circs = synth.random_circuits(700, 12, n_ops, 1, n_eff, seed=9, subsets=False)
    t = synth.make_tables(dim, n_ops, 1, n_eff, circs)
    D = synth.random_derivs(t, 70, density=0.05, seed=3)
    at = gpu_ctx.upload_atom(t); at.set_model(G, rho, E); at.set_derivs(D)
    J = np.empty((t.n_elements, 70)); at.fill_dprobs(J)
    p_bad = int(D.cols[0])
    vals = D.vals.copy(); vals[0] = np.inf                       # one derivative entry infinite: column p_bad of J becomes Inf / NaN
    from pygsti_b200.packing import DerivMap
    at.set_derivs(DerivMap(D.n_w, D.n_params, D.rows, D.cols, vals))
    try:
        gpu_ctx.set_jtj_mode(8); X, _ = at.jtj(np.ones(t.n_elements), np.zeros(t.n_elements))
    finally:
        gpu_ctx.set_jtj_mode(-1)
    ok = np.ones(70, bool); ok[p_bad] = False
    assert np.isnan(X[p_bad]).all() and np.isnan(X[:, p_bad]).all()
    ref = J[:, ok].T @ J[:, ok]
    assert np.max(np.abs(X[np.ix_(ok, ok)] - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    at.free()
