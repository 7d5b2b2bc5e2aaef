"""Host-side logic of the pyGSTi plug-in (no GPU): class plumbing that the reference's traps demand
(SURVEY.md App. A 'Boundary traps'), packing of layouts/models, and loud failure without a device."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
if os.path.isdir(os.path.join(REF, "pygsti")) and REF not in sys.path:
    sys.path.insert(0, REF)
pygsti = pytest.importorskip("pygsti", reason="reference install (baseline/_ref) not present")

from pygsti.modelpacks import smq1Q_XYI  # noqa: E402
from pygsti_b200 import packing, engine, _lib  # noqa: E402
from pygsti_b200.forwardsim import B200ForwardSimulator  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


@pytest.fixture(scope="module")
def model():
    return smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)


def test_sim_assignment_copy_and_serialization(model):
    m = model.copy()
    m.sim = B200ForwardSimulator(num_atoms=2, derivative_mode='fd')
    assert isinstance(m.sim, B200ForwardSimulator) and m.sim.model is m
    assert m.sim.calclib.__name__ == "pygsti_b200.calclib"
    m2 = m.copy()                      # Model.copy deep-copies the simulator through __getstate__
    assert isinstance(m2.sim, B200ForwardSimulator) and m2.sim.model is m2
    assert m2.sim.derivative_mode == 'fd' and m2.sim.calclib.__name__ == "pygsti_b200.calclib"
    s = m.sim.copy()                   # MapForwardSimulator.copy hard-codes its class: must be overridden
    assert isinstance(s, B200ForwardSimulator) and s._num_atoms == 2
    st = m.sim.to_nice_serialization()
    s2 = B200ForwardSimulator.from_nice_serialization(st)
    assert isinstance(s2, B200ForwardSimulator) and s2.derivative_mode == 'fd'


def test_layout_is_the_reference_layout_and_no_jacobian_table(model):
    m = model.copy()
    m.sim = B200ForwardSimulator()
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data
    layout = m.sim.create_layout(circuits, array_types=('e', 'ep'))
    assert type(layout).__name__ == "MapCOPALayout"
    assert not hasattr(layout.atoms[0], "jac_table")     # trap 3: the generic calclib would build it


def test_layouts_skip_the_host_prefix_cache_plan_by_default(model):
    """The engine rebuilds prefix + suffix sharing itself: layouts are created without pyGSTi's cache plan unless asked;
    the packed tables of both variants describe the same circuits (same oracle probabilities)."""
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    outs = []
    for kw, want_cache in (({}, False), ({"host_prefix_cache": True}, True), ({"max_cache_size": 5}, True)):
        m = model.copy()
        m.sim = B200ForwardSimulator(**kw)
        layout = m.sim.create_layout(circuits, array_types=('e', 'ep'))
        atom = layout.atoms[0]
        assert (atom.cache_size > 0) == want_cache
        s = m.sim.copy()
        assert s._max_cache_size == m.sim._max_cache_size and s.host_prefix_cache == m.sim.host_prefix_cache
        t = packing.pack_atom(atom, m.dim)
        mt = packing.pack_model(m, atom, m.dim)
        p = np.empty(layout.num_elements)
        p[:] = onp.mapfill_probs(t, mt.G, mt.rho, mt.E)
        outs.append({c: p[layout.indices(c)] for c in circuits})
    for c in circuits:
        assert np.max(np.abs(outs[0][c] - outs[1][c])) <= 1e-15 and np.max(np.abs(outs[0][c] - outs[2][c])) <= 1e-15


def test_packing_reproduces_reference_probs_through_the_oracle(model):
    m = model.copy()
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    m.sim = 'map'
    layout = m.sim.create_layout(circuits, array_types=('e', 'ep'))
    ref = np.empty(layout.num_elements); m.sim.bulk_fill_probs(ref, layout)
    atom = layout.atoms[0]
    t = packing.pack_atom(atom, m.dim)
    mt = packing.pack_model(m, atom, m.dim)
    out = onp.mapfill_probs(t, mt.G, mt.rho, mt.E)
    assert np.max(np.abs(out - ref)) <= 1e-14
    D = packing.pack_derivs(m, atom, m.dim, slice(10, 30))
    assert D.n_params == 20 and D.cols.max() < 20
    Dfull = packing.pack_derivs(m, atom, m.dim)
    J = onp.dprobs_analytic(t, mt.G, mt.rho, mt.E, Dfull)
    Jb = onp.dprobs_analytic(t, mt.G, mt.rho, mt.E, D)
    assert np.max(np.abs(J[:, 10:30] - Jb)) == 0.0


def test_fill_fails_loudly_without_gpu(model):
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    m = model.copy()
    m.sim = B200ForwardSimulator()
    circuits = smq1Q_XYI.create_gst_experiment_design(1).all_circuits_needing_data
    with pytest.raises(_lib.B200Error):
        m.sim.bulk_probs(circuits)
    with pytest.raises(ValueError):
        B200ForwardSimulator(derivative_mode='nope')


@pytest.mark.parametrize("pack,param,n_lind", [("smq1Q_XYI", "CPTPLND", 3 + 1 + 2), ("smq1Q_XYI", "H+S", 6), ("smq2Q_XYCNOT", "CPTPLND", 5 + 1 + 4)])
def test_lindblad_inputs_reassemble_model_and_derivative_map(pack, param, n_lind):
    """`packing.pack_lindblad` (host side of the planned on-device update of Lindblad members): the model tensors M and the
    derivative map D rebuilt from (B, c, dc/dtheta, static part) with the oracle's dense algebra equal pack_model / pack_derivs."""
    import importlib
    from oracle import oracle_lindblad as ol
    mp = importlib.import_module("pygsti.modelpacks." + pack)
    m = mp.target_model(param)
    rng = np.random.default_rng(2)
    m.from_vector(m.to_vector() + 1e-2 * rng.standard_normal(m.num_params))
    m.sim = B200ForwardSimulator()
    circuits = mp.create_gst_experiment_design(1).all_circuits_needing_data
    atom = m.sim.create_layout(circuits, array_types=('e', 'ep')).atoms[0]
    d = m.dim
    mt = packing.pack_model(m, atom, d)
    M_ref = np.concatenate([mt.G.ravel(), mt.rho.ravel(), mt.E.ravel()])
    D_ref = packing.pack_derivs(m, atom, d)
    D_dense = np.zeros((D_ref.n_w, D_ref.n_params)); D_dense[D_ref.rows, D_ref.cols] = D_ref.vals
    li = packing.pack_lindblad(m, atom, d)
    assert len(li.members) == n_lind and not li.host_members
    M = np.full(M_ref.size, np.nan); Dd = np.zeros_like(D_dense)
    for lm in li.members:
        eg = li.errgens[lm.errgen]
        B = eg.B_re + 1j * eg.B_im
        if lm.kind == "op":
            val, dval = ol.composed_gate(eg.c, eg.dc, B, lm.static)
            val = val.ravel()
        elif lm.kind == "rho":
            val, dval = ol.composed_state(eg.c, eg.dc, B, lm.static)
        else:
            val, dval = ol.composed_effect(eg.c, eg.dc, B, lm.static)
        M[lm.w_offset:lm.w_offset + val.size] = val
        Dd[lm.w_offset:lm.w_offset + val.size][:, lm.gpindices] += dval
    assert np.max(np.abs(M - M_ref)) <= 1e-12
    assert np.max(np.abs(Dd - D_dense)) <= 1e-10
    # packing.assemble_lindblad with the oracle standing in for the device call (engine.Context.lindblad_members)
    outs = []
    for lm in li.members:
        eg = li.errgens[lm.errgen]
        B = eg.B_re + 1j * eg.B_im
        f = {"op": ol.composed_gate, "rho": ol.composed_state, "eff": ol.composed_effect}[lm.kind]
        val, dval = f(eg.c, eg.dc, B, lm.static)
        outs.append((np.ravel(val), dval))
    for pidx in (None, slice(3, 11)):
        mt2, D2 = packing.assemble_lindblad(li, outs, m, atom, d, pidx)
        Dr = packing.pack_derivs(m, atom, d, pidx)
        assert max(np.max(np.abs(mt2.G - mt.G)), np.max(np.abs(mt2.rho - mt.rho)), np.max(np.abs(mt2.E - mt.E))) <= 1e-12
        A = np.zeros((D2.n_w, D2.n_params)); A[D2.rows, D2.cols] = D2.vals
        R = np.zeros((Dr.n_w, Dr.n_params)); R[Dr.rows, Dr.cols] = Dr.vals
        assert A.shape == R.shape and np.max(np.abs(A - R)) <= 1e-10


def test_member_hessian_cache_equals_filtered_calls_and_follows_the_parameters():
    """pack_hessians slices rectangles out of each member's full Hessian, computed once per parameter vector: same numbers as the
    members' own filtered `hessian_wrt_params(l1, l2)`, recomputed when the parameters change."""
    m = smq1Q_XYI.target_model('CPTPLND')
    rng = np.random.default_rng(0)
    m.sim = B200ForwardSimulator()
    atom = m.sim.create_layout(smq1Q_XYI.create_gst_experiment_design(1).all_circuits_needing_data, array_types=('e', 'ep')).atoms[0]
    ops, rhos, effs = packing._members(m, atom)
    lo1, hi1, lo2, hi2 = 3, 20, 10, 40
    for step in range(2):
        m.from_vector(m.to_vector() + 1e-2 * rng.standard_normal(m.num_params))
        H = packing.pack_hessians(m, atom, 4, slice(lo1, hi1), slice(lo2, hi2))
        dense = np.zeros((H.n_w, H.n1, H.n2)); dense[H.rows, H.a, H.b] = H.vals
        ref = np.zeros_like(dense); off = 0
        for group, size in ((ops, 16), (rhos, 4), (effs, 4)):
            for mem in group:
                gp = packing._gp_array(mem.gpindices)
                l1 = [i for i, g in enumerate(gp) if lo1 <= g < hi1]; l2 = [i for i, g in enumerate(gp) if lo2 <= g < hi2]
                if gp.size and mem.has_nonzero_hessian() and l1 and l2:
                    Hm = np.real(np.asarray(mem.hessian_wrt_params(l1, l2))).reshape(size, len(l1), len(l2))
                    ref[off:off + size][:, (gp[l1] - lo1)[:, None], (gp[l2] - lo2)[None, :]] += Hm
                off += size
        assert np.max(np.abs(ref)) > 0 and np.max(np.abs(dense - ref)) <= 1e-14


def test_factored_packing_reproduces_to_dense():
    """Factor programs of pyGSTi's non-dense reps (ComposedOp of EmbeddedOp; opcreps.cpp:93-158, 242-276): what the device
    does with them (packing.factored_to_dense is the numpy restatement of k_factored_to_dense) equals the reference's own
    to_dense for every layer label of a 4-qubit cloud-crosstalk model and of a 3-qubit local-noise model, bit for bit; models
    whose layers are not products of embedded 1-2 qubit operations are declined (dense packing then)."""
    from pygsti.processors import QubitProcessorSpec
    from pygsti.models import modelconstruction as mc
    from pygsti.forwardsims import MapForwardSimulator
    from pygsti.circuits import Circuit
    rng = np.random.default_rng(7)
    for nq, build in ((4, lambda ps: mc.create_cloud_crosstalk_model(ps, lindblad_error_coeffs={
            ('Gxpi2', 0): {('H', 'X'): 0.01, ('S', 'Z'): 0.005}, ('Gcnot', 1, 2): {('H', 'ZZ'): 0.01, ('S', 'XX'): 0.003}})),
                      (3, lambda ps: mc.create_crosstalk_free_model(ps, ideal_gate_type='full TP', ideal_spam_type='full TP'))):
        pspec = QubitProcessorSpec(nq, ['Gxpi2', 'Gypi2', 'Gcnot'], geometry='line')
        model = build(pspec)
        prim = list(model.primitive_op_labels)
        circs = [Circuit([prim[int(rng.integers(len(prim)))] for _ in range(int(rng.integers(1, 9)))], line_labels=tuple(range(nq)))
                 for _ in range(10)]
        model.sim = MapForwardSimulator()
        atom = model.sim.create_layout(circs, array_types=('e',)).atoms[0]
        d = 4 ** nq
        mt = packing.pack_model(model, atom, d)
        fm = packing.pack_model_factored(model, atom, d)
        assert fm is not None and fm.n_qubits == nq and fm.op_fptr[-1] == fm.f_nq.size >= mt.G.shape[0]
        assert np.array_equal(packing.factored_to_dense(fm, d), mt.G)
        assert np.array_equal(fm.rho, mt.rho) and np.array_equal(fm.E, mt.E)
    m1 = smq1Q_XYI.target_model()                           # dense 1-qubit gates: not a factored model
    m1.sim = MapForwardSimulator()
    a1 = m1.sim.create_layout(smq1Q_XYI.create_gst_experiment_design(1).all_circuits_needing_data, array_types=('e',)).atoms[0]
    assert packing.pack_model_factored(m1, a1, 4) is None


def test_factor_space_derivative_map_equals_dense_map():
    """packing.pack_derivs_factored (rows = entries of the small embedded operations, what b200_atom_set_derivs_factored takes) is
    consistent with packing.pack_derivs (rows = entries of the dense layer operations, pinned against the reference): for the
    3-qubit crosstalk-free model every layer is ONE embedded factor, so  d(dense layer)/d theta = embed(d(factor)/d theta)
    -- the chain rule the device applies implicitly; SPAM rows are identical."""
    from pygsti.processors import QubitProcessorSpec
    from pygsti.models import modelconstruction as mc
    from pygsti.forwardsims import MapForwardSimulator
    from pygsti.circuits import Circuit
    pspec = QubitProcessorSpec(3, ['Gxpi2', 'Gypi2', 'Gcnot'], geometry='line')
    model = mc.create_crosstalk_free_model(pspec, ideal_gate_type='full TP', ideal_spam_type='full TP')
    rng = np.random.default_rng(3)
    model.from_vector(model.to_vector() + 0.01 * rng.standard_normal(model.num_params))
    prim = list(model.primitive_op_labels)
    circs = [Circuit([prim[int(rng.integers(len(prim)))] for _ in range(int(rng.integers(1, 7)))], line_labels=(0, 1, 2)) for _ in range(12)]
    model.sim = MapForwardSimulator()
    atom = model.sim.create_layout(circs, array_types=('e', 'ep')).atoms[0]
    d = 64
    fm = packing.pack_model_factored(model, atom, d)
    Df = packing.pack_derivs_factored(model, atom, d, fm)
    Dd = packing.pack_derivs(model, atom, d)
    assert fm is not None and Df is not None and Df.n_params == Dd.n_params == model.num_params
    n_ops = fm.op_fptr.size - 1
    assert np.array_equal(np.diff(fm.op_fptr), np.ones(n_ops, dtype=fm.op_fptr.dtype))      # one factor per layer label
    dense_f = np.zeros((Df.n_w, Df.n_params)); np.add.at(dense_f, (Df.rows, Df.cols), Df.vals)
    dense_d = np.zeros((Dd.n_w, Dd.n_params)); np.add.at(dense_d, (Dd.rows, Dd.cols), Dd.vals)
    n_mats = fm.mats.size
    assert np.array_equal(dense_f[n_mats:], dense_d[n_ops * d * d:])                        # rho and effect rows
    idx = np.arange(d)
    for g in range(n_ops):
        k = int(fm.f_nq[g]); ds = 4 ** k
        shifts = [2 * (3 - 1 - int(q)) for q in fm.f_targets[g, :k]]
        tmask = 0
        for sh in shifts:
            tmask |= 3 << sh
        t = np.zeros(d, dtype=np.int64)
        for sh in shifts:
            t = (t << 2) | ((idx >> sh) & 3)
        rest = idx & ~tmask
        same_rest = rest[:, None] == rest[None, :]
        small_rows = int(fm.f_moff[g]) + (t[:, None] * ds + t[None, :])                     # factor-space row feeding dense entry (i, j)
        emb = np.where(same_rest[:, :, None], dense_f[small_rows], 0.0).reshape(d * d, -1)
        assert np.max(np.abs(emb - dense_d[g * d * d:(g + 1) * d * d])) <= 1e-14, g
    assert np.count_nonzero(dense_f) < np.count_nonzero(dense_d)                            # every small entry appears 4 / 16 times in the dense layer


def test_stock_objective_hooks_install_and_restore():
    """The simulator installs the objective-function / LM hooks once; they gate on the simulator and the objective (penalty rows,
    omitted outcomes and `derivative_mode='fd'` go to pyGSTi's own methods) and `uninstall_hooks` restores pyGSTi."""
    from pygsti.objectivefns import objectivefns as _of
    from pygsti.layouts.distlayout import DistributableCOPALayout as _DL
    from pygsti.data import simulate_data
    from pygsti_b200 import objective as fused
    fused.uninstall_hooks()
    orig = (_of.TimeIndependentMDCObjectiveFunction.dlsvec, _DL.fill_jtj)
    m = smq1Q_XYI.target_model("full TP").depolarize(op_noise=0.05, spam_noise=0.02)
    circuits = smq1Q_XYI.create_gst_experiment_design(1).all_circuits_needing_data
    ds = simulate_data(m, circuits, 100, seed=1)
    m.sim = B200ForwardSimulator()
    assert fused._HOOKS and _of.TimeIndependentMDCObjectiveFunction.dlsvec is not orig[0] and _DL.fill_jtj is not orig[1]
    o = _of.Chi2Function.create_from(m, ds, circuits, method_names=('lsvec', 'dlsvec'))
    assert fused._sim_wants_hooks(o) and fused._plain(o)
    o_pen = _of.Chi2Function.create_from(m, ds, circuits, method_names=('lsvec', 'dlsvec'), penalties={'cptp_penalty_factor': 1.0})
    assert not fused._plain(o_pen)                           # penalty rows: the reference's own dlsvec
    m2 = m.copy(); m2.sim = B200ForwardSimulator(derivative_mode='fd')
    assert not fused._plain(_of.Chi2Function.create_from(m2, ds, circuits, method_names=('lsvec', 'dlsvec')))
    m3 = m.copy(); m3.sim = B200ForwardSimulator(fused_objective=False)
    assert not fused._sim_wants_hooks(_of.Chi2Function.create_from(m3, ds, circuits, method_names=('lsvec', 'dlsvec')))
    # arrays the hooks did not produce fall through to pyGSTi's own fill_jtj
    J = np.random.default_rng(0).standard_normal((o.layout.num_elements, m.num_params)); jtj = np.empty((m.num_params,) * 2)
    o.layout.fill_jtj(J, jtj)
    assert np.allclose(jtj, J.T @ J)
    fused.uninstall_hooks()
    assert (_of.TimeIndependentMDCObjectiveFunction.dlsvec, _DL.fill_jtj) == orig and not fused._HOOKS
    fused.install_hooks()
