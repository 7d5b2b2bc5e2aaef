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
    # Size-independent identity checked on ALL 273 340 rows.  Every member of this model is `full`: its parameters ARE its
    # elements, and p is homogeneous of degree (number of occurrences) in each member, so by Euler's theorem
    #     sum_{params of gate g} theta * dp/dtheta = count_g(circuit) * p,   and = p for the prep and for the effect of the row.
    D = a["D"]
    M = np.concatenate([a["G"].ravel(), a["rho"].ravel(), a["E"].ravel()])
    theta = np.zeros(Np); theta[D.cols] = M[D.rows]
    counts = _gate_counts(t)                                             # [n_rows, n_ops]
    row_of_el = np.empty(nE, np.int64); eff_of_el = np.empty(nE, np.int64)
    row_of_el[t.out_el] = np.repeat(np.arange(t.n_rows), np.diff(t.out_ptr)); eff_of_el[t.out_el] = t.out_eff
    scale = max(1.0, float(np.max(np.abs(p))))
    for g in range(t.n_ops):
        cols = D.cols[(D.rows >= g * 256) & (D.rows < (g + 1) * 256)]
        lhs = J[:, cols] @ theta[cols]
        assert np.max(np.abs(lhs - counts[row_of_el, g] * p)) <= 1e-10 * scale * max(1, counts.max()), g
    off_rho, off_eff = t.n_ops * 256, t.n_ops * 256 + t.n_rho * 16
    cols = D.cols[(D.rows >= off_rho) & (D.rows < off_eff)]
    assert np.max(np.abs(J[:, cols] @ theta[cols] - p)) <= 1e-10 * scale
    for e in range(t.n_eff):
        cols = D.cols[(D.rows >= off_eff + e * 16) & (D.rows < off_eff + (e + 1) * 16)]
        lhs = J[:, cols] @ theta[cols]
        assert np.max(np.abs(lhs - np.where(eff_of_el == e, p, 0.0))) <= 1e-10 * scale, e
    at.free()


def _gate_counts(t):
    """Occurrences of every gate in the FULL circuit of every prefix-table row (cache links followed)."""
    counts = np.zeros((t.n_rows, t.n_ops), np.int64)
    slot = {}
    for k in range(t.n_rows):
        c = np.bincount(t.row_ops[t.row_ptr[k]:t.row_ptr[k + 1]], minlength=t.n_ops)
        if t.row_istart[k] >= 0:
            c = c + counts[slot[int(t.row_istart[k])]]
        counts[k] = c
        if t.row_icache[k] >= 0:
            slot[int(t.row_icache[k])] = k
    return counts


def test_full_size_config4_cptplnd(gpu_ctx, load_case):
    """BASELINE config 4 at the size SURVEY 8d names (CPTPLND, Np = 1680, 7860 circuits / 31 440 outcomes): probabilities
    against the reference's Map simulator, sampled Jacobian rows and a sampled 64 x 64 Hessian rectangle against the
    reference's Matrix simulator (goldens: make_golden.py c4_gst16_layout), plus the TP property on all circuits."""
    c = load_case("c4_gst16_layout")
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    nE, Np = c.n_elements, c.num_params
    J = engine.pinned_empty((nE, Np)); p = np.empty(nE)
    at.fill_dprobs(J, p)
    st = int(c["probs_map_stride"])
    assert np.max(np.abs(p[::st] - c["probs_map_sample"])) <= PROBS_TOL
    assert abs(p.sum() - float(c["probs_map_sum"])) <= 1e-9
    rows = c["dprobs_matrix_sample_elements"]
    assert np.max(np.abs(J[rows] - c["dprobs_matrix_sample_rows"])) <= DPROBS_TOL
    t = a["tables"]
    # CPTP model: every circuit's probabilities sum to 1 and every circuit's Jacobian rows sum to 0, for ALL circuits
    assert np.max(np.abs(np.add.reduceat(p[t.out_el], t.out_ptr[:-1]) - 1.0)) <= 1e-12
    assert np.max(np.abs(np.add.reduceat(J[t.out_el], t.out_ptr[:-1], axis=0))) <= 1e-10
    r = c["hess_rects"][0]
    p1, p2 = np.arange(r[0], r[1]), np.arange(r[2], r[3])
    H = np.empty((nE, p1.size, p2.size))
    at.fill_hprobs(p1, p2, H, c.hess_map("H2r0"))
    assert np.max(np.abs(H[rows] - c["hprobs_matrix_rect0_sample_rows"])) <= 1e-9
    assert np.max(np.abs(np.add.reduceat(H[t.out_el], t.out_ptr[:-1], axis=0))) <= 1e-9     # second derivative of sum_j p_j = 1
    at.free()


@pytest.mark.parametrize("factored", [False, True])
def test_full_size_config3_d64(gpu_ctx, load_case, factored):
    """BASELINE config 3 on one GPU, 5000 of the 50 000 random circuits of the bench stream (d = 64, Np = 775, depth U{1..256}; model
    tensors from the reference): sampled circuits against the oracle (itself pinned to the reference's Matrix goldens for this
    model, tests/test_oracle_cpu.py), and for ALL rows the size-independent TP property: sum_j p_j = 1, sum_j dp_j = 0.
    Both device forms of the model: dense 64 x 64 gates (level-batched sweeps) and factor programs + factor-space map
    (kernels_factoredj.cuh, the path bench.py times)."""
    from pygsti_b200 import fixtures as fx
    from oracle import oracle_c
    c = load_case("c3_3q_localnoise_sub"); a = c.atoms[0]
    n_ops, n_eff, Np = a["tables"].n_ops, a["tables"].n_eff, a["D"].n_params
    t, circs = fx.random_layout(64, n_ops, n_eff, 50000, 256, seed=0, rows=(0, 5000))
    at = gpu_ctx.upload_atom(t)
    if factored:
        at.set_model_factored(a["fm"]); at.set_derivs_factored(a["Df"])
    else:
        at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
    J = engine.pinned_empty((t.n_elements, Np)); p = np.empty(t.n_elements)
    at.fill_dprobs(J, p)
    assert np.max(np.abs(p.reshape(-1, n_eff).sum(axis=1) - 1.0)) <= 1e-12
    assert np.max(np.abs(J.reshape(-1, n_eff, Np).sum(axis=1))) <= 1e-10
    orc = oracle_c.Oracle("port")
    for lo in (0, 2497, 4990):
        sub, _ = fx.random_layout(64, n_ops, n_eff, 50000, 256, seed=0, rows=(lo, lo + 6))
        Jo, po = orc.dprobs_analytic(sub, a["G"], a["rho"], a["E"], a["D"])
        sl = slice(lo * n_eff, (lo + 6) * n_eff)
        assert np.max(np.abs(p[sl] - po)) <= PROBS_TOL
        assert np.max(np.abs(J[sl] - Jo)) <= DPROBS_TOL
    # J^T J / J^T f with an ODD number of parameters (775): the hand-written DMMA reduction against numpy
    rs = np.random.default_rng(0).uniform(0.5, 1.5, t.n_elements); f = np.random.default_rng(1).standard_normal(t.n_elements)
    JTJ, JTf = at.jtj(rs, f)
    Js = J * rs[:, None]
    ref = Js.T @ Js
    assert np.max(np.abs(JTJ - ref)) <= 1e-11 * np.max(np.abs(ref)) and np.array_equal(JTJ, JTJ.T)
    assert np.max(np.abs(JTf - Js.T @ f)) <= 1e-11 * np.max(np.abs(Js.T @ f))
    at.free()


def test_full_size_config5_d256(gpu_ctx):
    """BASELINE config 5 at full size (d = 256, 14 layer labels, 5000 random circuits of depth U{1..128}, 16 outcomes): sampled
    circuits against the oracle and, for ALL circuits, linearity in the effects: with E_15 := sum of the other effects the
    last outcome's probability must equal the sum of the others."""
    from pygsti_b200 import fixtures as fx
    from oracle import oracle_c
    G, rho, E = fx.random_dense_model(256, 14, 1, 16, seed=1)
    E = E.copy(); E[15] = E[:15].sum(axis=0)
    t, circs = fx.random_layout(256, 14, 16, 5000, 128, seed=0)
    at = gpu_ctx.upload_atom(t); at.set_model(G, rho, E)
    p = np.empty(t.n_elements); at.fill_probs(p)
    P = p.reshape(-1, 16)
    assert np.max(np.abs(P[:, :15].sum(axis=1) - P[:, 15])) <= 1e-12 * max(1.0, np.max(np.abs(P)))
    orc = oracle_c.Oracle("port")
    for lo in (0, 2500, 4980):
        sub, _ = fx.random_layout(256, 14, 16, 5000, 128, seed=0, rows=(lo, lo + 20))
        po = orc.mapfill_probs(sub, G, rho, E)
        assert np.max(np.abs(p[lo * 16:(lo + 20) * 16] - po)) <= PROBS_TOL * max(1.0, np.max(np.abs(po)))
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


@pytest.mark.parametrize("jtj_mode", [0, 8, 7])
@pytest.mark.parametrize("name", ["c2_2q_full_sub", "c4_2q_cptplnd_sub", "c1_1q_tp", "c3_3q_localnoise_sub"])
def test_scaled_jacobian_and_jtj(gpu_ctx, load_case, name, jtj_mode):
    """b200_fill_dprobs_scaled / b200_jtj (fused objective Jacobian fill) on every kernel path: fused d=16 (trie
    epilogue), general d=16 (W + contraction), d=4 and d=64 generic; the contraction by the FP64 DMMA SYRK (mode 0) and by
    the tcgen05 int8 Ozaki SYRK with 8 / 7 digits (forced here: the automatic mode uses it from 4096 rows up)."""
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
    gpu_ctx.set_jtj_mode(jtj_mode)
    try:
        JTJ, JTf = at.jtj(w, f)
    finally:
        gpu_ctx.set_jtj_mode(-1)
    R = Jref.T @ Jref
    assert np.max(np.abs(JTJ - R)) <= 1e-10 * max(1.0, np.max(np.abs(R)))
    assert np.max(np.abs(JTJ - JTJ.T)) == 0.0
    assert np.max(np.abs(JTf - Jref.T @ f)) <= 1e-10 * max(1.0, np.max(np.abs(Jref.T @ f)))
    at.free()


def test_jtj_tcgen05_full_size_vs_fp64(gpu_ctx, load_case):
    """BASELINE config 2 at full size (273 340 x 1360): the tcgen05 Ozaki J^T J (8 and 7 digits; several K slices per tile,
    int32 accumulators near their exactness bound) against the FP64 DMMA SYRK and against a long-double host product."""
    c = load_case("c2_full_layout")
    a = c.atoms[0]
    at = _atom(gpu_ctx, a)
    rng = np.random.default_rng(5)
    w = rng.uniform(0.5, 1.5, size=c.n_elements) * 10.0 ** rng.integers(-3, 4, size=c.n_elements)   # rows over 7 decades
    f = rng.standard_normal(c.n_elements)
    res = {}
    try:
        for mode in (0, 8, 7):
            gpu_ctx.set_jtj_mode(mode)
            res[mode] = at.jtj(w, f)
    finally:
        gpu_ctx.set_jtj_mode(-1)
    R, Rf = res[0]
    sc = np.max(np.abs(R))
    for mode, tol in ((8, 1e-13), (7, 1e-12)):
        X, Xf = res[mode]
        assert np.isfinite(X).all()
        assert np.max(np.abs(X - X.T)) == 0.0
        assert np.max(np.abs(X - R)) <= tol * sc, mode
        assert np.max(np.abs(Xf - Rf)) <= 1e-12 * np.max(np.abs(Rf)), mode
    J = np.empty((c.n_elements, c.num_params))
    at.fill_dprobs(J, row_scale=w)
    cols = [0, 5, 401, c.num_params - 1]
    ref = np.zeros((len(cols), c.num_params), dtype=np.longdouble)          # float64 products of 4096-row blocks, summed in long double
    for r0 in range(0, c.n_elements, 4096):
        ref += J[r0:r0 + 4096, cols].T @ J[r0:r0 + 4096]
    ref = ref.astype(np.float64)
    for mode, tol in ((0, 1e-13), (8, 1e-13), (7, 1e-12)):
        assert np.max(np.abs(res[mode][0][cols] - ref)) <= tol * sc, mode
    at.free()


def test_factored_jacobian_golden_c3(gpu_ctx, load_case):
    """BASELINE config 3's model family (3-qubit crosstalk-free full TP model, d = 64, Np = 775): the reference holds its layers as
    EmbeddedOps; the engine takes the same factors + the factor-space derivative map (packing.pack_derivs_factored) and must
    reproduce the reference's Matrix simulator -- with only the factor-space map set, and with both maps set (then the factored
    kernels run and the dense map stays available to the Hessian paths)."""
    c = load_case("c3_3q_localnoise_sub")
    a = c.atoms[0]
    assert "fm" in a and "Df" in a
    ref = c["dprobs_matrix"]
    for both in (False, True):
        at = gpu_ctx.upload_atom(a["tables"])
        at.set_model_factored(a["fm"])
        if both:
            at.set_derivs(a["D"])
        at.set_derivs_factored(a["Df"])
        n0 = gpu_ctx.launch_count
        J = np.full((c.n_elements, c.num_params), np.nan); p = np.full(c.n_elements, np.nan)
        at.fill_dprobs(J, p)
        assert gpu_ctx.launch_count - n0 == 2          # k_fj_forward + k_fj_backward: the factored path ran
        assert np.max(np.abs(p - c["probs_matrix"])) <= 1e-12
        assert np.max(np.abs(J - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref)))
        at.free()
