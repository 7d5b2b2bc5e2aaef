"""
End-to-end drop-in tests on the GPU THROUGH pyGSTi's own API: ``model.sim = B200ForwardSimulator()``.

These re-state the reference's own forward-simulator tests with the new simulator added to the list of
simulators under test (pyGSTi test/unit/objects/test_forwardsim.py:278-348 ForwardSimConsistencyTester,
:351-378 ForwardSimIntegrationTester, :58-146 smoke tests; test/unit/mpi/run_me_with_mpiexec.py:176-261
atom / param-block equivalence).  They need the reference install (baseline/_ref, built by
oracle/build_ref.py; it is git-ignored but travels to the GPU box) and skip when pyGSTi is not importable;
tests/test_gpu_parity.py covers the same kernels against committed golden vectors without pyGSTi.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
if os.path.isdir(os.path.join(REF, "pygsti")) and REF not in sys.path:
    sys.path.insert(0, REF)
pygsti = pytest.importorskip("pygsti", reason="reference install (baseline/_ref) not present")

import scipy.linalg as la  # noqa: E402
from pygsti.modelpacks import smq1Q_XYI, smq2Q_XYCNOT  # noqa: E402
from pygsti.forwardsims import MapForwardSimulator, MatrixForwardSimulator  # noqa: E402
from pygsti.circuits import create_lsgst_circuit_lists  # noqa: E402
from pygsti_b200.forwardsim import B200ForwardSimulator  # noqa: E402

PROBS_TOL = 1e-14   # colinearity tolerances of the reference's own tester (test_forwardsim.py:280-281)
JACS_TOL = 1e-10


def _bulk_arrays(model, sim, circuits, want_jac=True):
    m = model.copy()
    m.sim = sim
    pr = m.sim.bulk_probs(circuits)
    probs = np.concatenate([np.array(list(pr[c].values())) for c in circuits])
    jac = None
    if want_jac:
        dp = m.sim.bulk_dprobs(circuits)
        jac = np.concatenate([np.array(list(dp[c].values())) for c in circuits])
    return probs, jac


def _colinearities(rows):
    _, _, vt = la.svd(rows, full_matrices=False)
    v = vt[0, :]
    col = (rows / la.norm(rows, axis=1)[:, None]) @ v
    if np.count_nonzero(col < 0) > col.size / 2:
        col *= -1
    return col


def test_consistency_tester_with_b200_simulator_added():
    model = smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    circuits = create_lsgst_circuit_lists(model, smq1Q_XYI.prep_fiducials(), smq1Q_XYI.meas_fiducials(),
                                          smq1Q_XYI.germs(), [4])[0]
    sims = [MapForwardSimulator(), MatrixForwardSimulator(), B200ForwardSimulator(),
            B200ForwardSimulator(derivative_mode='fd')]
    res = [_bulk_arrays(model, s, circuits) for s in sims]
    pcl = _colinearities(np.vstack([r[0] for r in res]))
    assert np.all(pcl >= 1 - PROBS_TOL), pcl
    jcl = _colinearities(np.stack([r[1].ravel() for r in res]))
    assert np.all(jcl >= 1 - JACS_TOL), jcl
    # and the absolute bar of BASELINE.json: 1e-10 vs the reference (probs vs Map, Jacobian vs analytic Matrix)
    assert np.max(np.abs(res[2][0] - res[0][0])) <= 1e-10
    assert np.max(np.abs(res[2][1] - res[1][1])) <= 1e-10


def test_model_probabilities_known_answer():
    """test_model.py:418-486 style: probabilities equal the explicit numpy product E^T Gy Gx rho."""
    model = smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    model.sim = B200ForwardSimulator()
    c = pygsti.circuits.Circuit([('Gxpi2', 0), ('Gypi2', 0)], line_labels=(0,))
    p = model.probabilities(c)
    Gx = model.operations[('Gxpi2', 0)].to_dense(); Gy = model.operations[('Gypi2', 0)].to_dense()
    rho = model.preps['rho0'].to_dense()
    for k, E in model.povms['Mdefault'].items():
        expect = float(E.to_dense() @ Gy @ Gx @ rho)
        assert abs(p[(k,)] - expect) <= 1e-12


@pytest.mark.parametrize("param", ["full", "full TP", "CPTPLND", "H+S"])
def test_parameterizations_vs_matrix_sim(param):
    model = smq1Q_XYI.target_model(param)
    v = model.to_vector(); rng = np.random.default_rng(1)
    model.from_vector(v + 5e-3 * rng.standard_normal(v.size))
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data[:150]
    pm, jm = _bulk_arrays(model, MatrixForwardSimulator(), circuits)
    pb, jb = _bulk_arrays(model, B200ForwardSimulator(), circuits)
    assert np.max(np.abs(pm - pb)) <= 1e-10
    assert np.max(np.abs(jm - jb)) <= 1e-10


def test_atoms_and_param_blocks_equal_single_atom():
    """run_me_with_mpiexec.py:176-261 logic without MPI: num_atoms in {1,4}, param_blk_sizes in {None,15}."""
    model = smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    ref = None
    for natoms, blk in [(1, None), (4, None), (1, (15,)), (4, (15,))]:
        m = model.copy()
        m.sim = B200ForwardSimulator(num_atoms=natoms, param_blk_sizes=blk)
        layout = m.sim.create_layout(circuits, array_types=('e', 'ep'))
        p = np.empty(layout.num_elements); dp = np.full((layout.num_elements, m.num_params), np.nan)
        m.sim.bulk_fill_dprobs(dp, layout, pr_array_to_fill=p)
        cur = {}
        for i, c in enumerate(layout.circuits):
            idx, outs = layout.indices_and_outcomes_for_index(i)
            cur[c] = (p[idx].copy(), dp[idx].copy())
        if ref is None:
            ref = cur
        else:
            for c in ref:
                assert np.max(np.abs(ref[c][0] - cur[c][0])) <= 1e-13
                assert np.max(np.abs(ref[c][1] - cur[c][1])) <= 1e-12


@pytest.mark.parametrize("param,tol", [("full", 1e-10), ("full TP", 1e-10), ("CPTPLND", 1e-10), ("H+S", 1e-10)])
def test_hprobs_vs_matrix_sim(param, tol):
    """Fully analytic device Hessian, 1e-10 vs the reference's analytic Matrix simulator: full / TP members are linear
    in their parameters; CPTPLND / H+S members add their own `hessian_wrt_params` term (b200_fill_hprobs).  The
    reference's own Map simulator (FD of FD) is off by ~3e-3 on the CPTPLND case."""
    model = smq1Q_XYI.target_model(param)
    v = model.to_vector(); rng = np.random.default_rng(2)
    model.from_vector(v + 5e-3 * rng.standard_normal(v.size))
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data[:30]
    mm = model.copy(); mm.sim = MatrixForwardSimulator()
    hm = mm.sim.bulk_hprobs(circuits)
    mb = model.copy(); mb.sim = B200ForwardSimulator()
    hb = mb.sim.bulk_hprobs(circuits)
    worst = 0.0
    for c in circuits:
        for k in hm[c]:
            worst = max(worst, np.max(np.abs(hm[c][k] - hb[c][k])))
    assert worst <= tol, worst


def test_hprobs_fd_driver_fallback():
    """analytic_hessian=False keeps the reference's own FD driver (mapforwardsim.py:394-438) over the analytic device
    Jacobian: limited by that driver's truncation error (hessian_eps = 1e-5)."""
    model = smq1Q_XYI.target_model("CPTPLND")
    v = model.to_vector(); rng = np.random.default_rng(2)
    model.from_vector(v + 5e-3 * rng.standard_normal(v.size))
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data[:10]
    mm = model.copy(); mm.sim = MatrixForwardSimulator()
    hm = mm.sim.bulk_hprobs(circuits)
    mb = model.copy(); mb.sim = B200ForwardSimulator(analytic_hessian=False)
    hb = mb.sim.bulk_hprobs(circuits)
    worst = max(np.max(np.abs(hm[c][k] - hb[c][k])) for c in circuits for k in hm[c])
    assert worst <= 3e-3, worst


def test_hprobs_rectangles_match_full_blocks():
    """iter_hprobs_by_rectangle (forwardsim.py:787-878) semantics used by the MLE Hessian."""
    model = smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data[:20]
    model.sim = B200ForwardSimulator()
    layout = model.sim.create_layout(circuits, array_types=('e', 'ep', 'epp'))
    full = np.empty((layout.num_elements, model.num_params, model.num_params))
    model.sim.bulk_fill_hprobs(full, layout)
    sl1, sl2 = slice(5, 25), slice(30, 60)
    for s1, s2, hblk in model.sim.iter_hprobs_by_rectangle(layout, [(sl1, sl2)], False):
        assert np.max(np.abs(hblk - full[:, s1, s2])) <= 1e-12


def test_two_qubit_full_model_vs_reference():
    model = smq2Q_XYCNOT.target_model().depolarize(op_noise=0.01, spam_noise=0.01)
    circuits = smq2Q_XYCNOT.create_gst_experiment_design(2).all_circuits_needing_data[::7][:120]
    pmap, _ = _bulk_arrays(model, MapForwardSimulator(), circuits, want_jac=False)
    pb, jb = _bulk_arrays(model, B200ForwardSimulator(), circuits)
    assert np.max(np.abs(pmap - pb)) <= 1e-10
    _, jm = _bulk_arrays(model, MatrixForwardSimulator(), circuits[:40])
    n = jm.shape[0]
    assert np.max(np.abs(jm - jb[:n])) <= 1e-10


def test_gst_fit_reaches_reference_quality():
    """ForwardSimIntegrationTester (test_forwardsim.py:351-378): full GST on noiseless data, 2*DeltaLogL <= 0.05."""
    from pygsti.protocols import gst, ProtocolData
    from pygsti.data import simulate_data
    from pygsti.tools import two_delta_logl
    design = smq1Q_XYI.create_gst_experiment_design(max_max_length=8)
    datagen = smq1Q_XYI.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    ds = simulate_data(datagen, design.all_circuits_needing_data, 20000, sample_error='none')
    proto = gst.GateSetTomography(smq1Q_XYI.target_model("full TP"), 'stdgaugeopt', name="testGST")
    results = proto.run(ProtocolData(design, ds), simulator=B200ForwardSimulator())
    mdl = results.estimates["testGST"].models['stdgaugeopt']
    assert isinstance(results.estimates["testGST"].models['final iteration estimate'].sim, B200ForwardSimulator)
    assert two_delta_logl(mdl, ds) <= 0.05


def _random_circuits(model, n, max_depth, line_labels, seed=0):
    rng = np.random.default_rng(seed)
    prim = list(model.primitive_op_labels)
    out = []
    for _ in range(n):
        L = int(rng.integers(1, max_depth + 1))
        out.append(pygsti.circuits.Circuit([prim[int(rng.integers(len(prim)))] for _ in range(L)], line_labels=line_labels))
    return out


def test_three_qubit_local_noise_model_d64():
    """BASELINE config 3 model family: 3-qubit crosstalk-free model (embedded / composed layer ops, tensor-product
    SPAM -> densified through to_dense at the boundary), d = 64.  probs vs the reference Cython Map simulator; the
    Jacobian of this model family is pinned against the reference's Matrix simulator by the committed golden
    fixture c3_3q_localnoise_sub (tests/test_gpu_parity.py)."""
    from pygsti.processors import QubitProcessorSpec
    from pygsti.models import modelconstruction as mc
    pspec = QubitProcessorSpec(3, ['Gxpi2', 'Gypi2', 'Gcnot'], geometry='line')
    model = mc.create_crosstalk_free_model(pspec, ideal_gate_type='full TP', ideal_spam_type='full TP')
    v = model.to_vector(); rng = np.random.default_rng(0)
    model.from_vector(v + 0.01 * rng.standard_normal(v.size))
    circuits = _random_circuits(model, 40, 40, (0, 1, 2), seed=5)
    pmap, _ = _bulk_arrays(model, MapForwardSimulator(), circuits, want_jac=False)
    pb, jb = _bulk_arrays(model, B200ForwardSimulator(), circuits)
    assert np.max(np.abs(pmap - pb)) <= 1e-10
    # FD check of the analytic device Jacobian against the reference Map simulator's own finite differences
    _, jmap = _bulk_arrays(model, MapForwardSimulator(), circuits[:6])
    n = jmap.shape[0]
    assert np.max(np.abs(jmap - jb[:n])) <= 2e-5
    # the simulator handed the layers over as factor programs and the derivative map in factor space: the Jacobian above came from
    # the factored kernels (csrc/kernels_factoredj.cuh) and must equal the reference's analytic (Matrix) Jacobian to 1e-10
    from pygsti_b200 import calclib as _cl
    m = model.copy(); m.sim = B200ForwardSimulator()
    layout = m.sim.create_layout(circuits[:6], array_types=('e', 'ep'))
    J = np.empty((layout.num_elements, m.num_params)); m.sim.bulk_fill_dprobs(J, layout)
    ents = [e for a in layout.atoms for e in _cl._ATOM_CACHE[a].values()]
    assert ents and all(e.get("fm") is not None for e in ents)
    mm = model.copy(); mm.sim = MatrixForwardSimulator()
    lm = mm.sim.create_layout(circuits[:6], array_types=('e', 'ep'))
    Jm = np.empty((lm.num_elements, mm.num_params)); mm.sim.bulk_fill_dprobs(Jm, lm)
    for i, c in enumerate(circuits[:6]):
        ib, ob = layout.indices_and_outcomes_for_index(i); im, om = lm.indices_and_outcomes(c)
        ib = np.arange(ib.start, ib.stop) if isinstance(ib, slice) else np.asarray(ib)
        im = np.arange(im.start, im.stop) if isinstance(im, slice) else np.asarray(im)
        lut = {o: int(k) for o, k in zip(om, im)}
        for o, k in zip(ob, ib):
            assert np.max(np.abs(J[int(k)] - Jm[lut[o]])) <= 1e-10


def test_four_qubit_cloud_crosstalk_model_d256():
    """BASELINE config 5 model family: 4-qubit cloud-crosstalk model, d = 256, NON-dense reference reps
    (OpCRep_Embedded / Composed / ExpErrorgen, computational POVM; evotype.py:97 prefer_dense_reps = False for
    dim > 64).  The engine sees them as dense superoperators (to_dense) and must reproduce the reference."""
    from pygsti.processors import QubitProcessorSpec
    from pygsti.models import modelconstruction as mc
    pspec = QubitProcessorSpec(4, ['Gxpi2', 'Gypi2', 'Gcnot'], geometry='line')
    errs = {('Gxpi2', 0): {('H', 'X'): 0.01, ('S', 'Z'): 0.005}, ('Gypi2', 1): {('H', 'Y'): 0.02, ('S', 'Z'): 0.004},
            ('Gcnot', 1, 2): {('H', 'ZZ'): 0.01, ('S', 'XX'): 0.003}}
    model = mc.create_cloud_crosstalk_model(pspec, lindblad_error_coeffs=errs)
    circuits = _random_circuits(model, 12, 12, (0, 1, 2, 3), seed=7)
    pmap, _ = _bulk_arrays(model, MapForwardSimulator(), circuits, want_jac=False)
    pb, _ = _bulk_arrays(model, B200ForwardSimulator(), circuits, want_jac=False)
    assert pmap.shape == pb.shape and pmap.size == 12 * 16
    assert np.max(np.abs(pmap - pb)) <= 1e-10


def test_stock_objective_and_lm_use_the_fused_path():
    """Missing-drop-in row of round 1: a STOCK objective function (`objfn.dlsvec`, `objfn.dterms`) and the products LM forms right
    after it (`layout.fill_jtj / fill_jtf`, `ari.norm2_jac`; simplerlm.py:663-678) on a B200ForwardSimulator go through the fused
    kernels (hooks installed by the simulator), and give what pyGSTi's own host code gives."""
    from pygsti.data import simulate_data
    from pygsti.objectivefns import objectivefns as _objfns
    from pygsti.optimize import arraysinterface as _ari
    from pygsti_b200 import objective as fused
    target = smq1Q_XYI.target_model("full TP")
    datagen = target.depolarize(op_noise=0.1, spam_noise=0.05)
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    ds = simulate_data(datagen, circuits, 1000, seed=1234)
    start = target.depolarize(op_noise=0.06, spam_noise=0.03)

    def make(**kw):
        m = start.copy(); m.sim = B200ForwardSimulator(**kw)
        return _objfns.Chi2Function.create_from(m, ds, circuits, method_names=('lsvec', 'dlsvec', 'dterms'))

    assert fused._HOOKS, "hooks are installed by the simulator constructor"
    hooked, plain = make(), make(fused_objective=False)
    v = start.to_vector()
    J1 = hooked.dlsvec(v); J0 = plain.dlsvec(v).copy()
    assert J1.ctypes.data in fused._LAST and fused._LAST[J1.ctypes.data]["kind"] == "dlsvec"       # the fused path ran
    assert np.max(np.abs(J1 - J0)) <= 1e-11 * max(1.0, np.max(np.abs(J0)))
    f = hooked.lsvec(v)
    nP = J1.shape[1]
    ari = _ari.DistributedArraysInterface(hooked.layout, 'normal', 0)
    jtj = np.empty((nP, nP)); jtf = np.empty(nP)
    n2 = ari.norm2_jac(J1)
    ari.fill_jtj(J1, jtj, None); ari.fill_jtf(J1, f, jtf)
    assert fused._LAST[J1.ctypes.data]["jtj"] is not None                                          # ... came from the device
    R = J0.T @ J0
    assert np.max(np.abs(jtj - R)) <= 1e-10 * np.max(np.abs(R))
    assert np.max(np.abs(jtf - J0.T @ plain.lsvec(v))) <= 1e-10 * max(1.0, np.max(np.abs(J0.T @ f)))
    assert abs(n2 - np.linalg.norm(J0) ** 2) <= 1e-10 * n2
    # a Jacobian the hooks do not know (a copy) and a model that has moved on fall back to pyGSTi's own code
    Jc = J0.copy(); jtj2 = np.empty((nP, nP))
    ari.fill_jtj(Jc, jtj2, None)
    assert np.max(np.abs(jtj2 - R)) <= 1e-12 * np.max(np.abs(R))
    hooked.model.from_vector(v + 1e-3)
    assert fused._record_for(J1) is None
    D1 = hooked.dterms(v); D0 = plain.dterms(v)
    assert np.max(np.abs(D1 - D0)) <= 1e-11 * max(1.0, np.max(np.abs(D0)))


def test_fused_objective_with_term_weights():
    """A TermWeighted objective ('normalized tvd': rows weighted by 1 / circuit size, objectivefns.py:5175-5192): the weights
    must reach the fused row scale -- dterms, dlsvec AND J^T J / J^T f -- exactly as `_reweight_jac` applies them."""
    from pygsti.data import simulate_data
    from pygsti.objectivefns import objectivefns as _objfns
    from pygsti_b200 import objective as fused
    target = smq1Q_XYI.target_model("full TP")
    datagen = target.depolarize(op_noise=0.1, spam_noise=0.05)
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    ds = simulate_data(datagen, circuits, 1000, seed=1234)
    start = target.depolarize(op_noise=0.06, spam_noise=0.03)

    def make(**kw):
        m = start.copy(); m.sim = B200ForwardSimulator(**kw)
        return _objfns.TVDFunction.create_from(m, ds, circuits, name='normalized tvd', method_names=('lsvec', 'dlsvec', 'dterms'))

    ours, same = make(), make(fused_objective=False)
    v = start.to_vector()
    w = fused._term_weights(ours)
    assert w.shape == (ours.nelements,) and np.ptp(w) > 0.1                 # genuinely non-trivial weights
    Jsame = same.dlsvec(v).copy(); Dsame = same.dterms(v).copy(); fsame = same.lsvec(v).copy()
    Df = fused.fused_dterms(ours, v).copy(); Jf = fused.fused_dlsvec(ours, v).copy()
    assert np.max(np.abs(Df - Dsame)) <= 1e-11 * max(1.0, np.max(np.abs(Dsame)))
    assert np.max(np.abs(Jf - Jsame)) <= 1e-11 * max(1.0, np.max(np.abs(Jsame)))
    JTJ, JTf = fused.fused_jtj(ours, v)
    R = Jsame.T @ Jsame
    assert np.max(np.abs(JTJ - R)) <= 1e-10 * np.max(np.abs(R))
    assert np.max(np.abs(JTf - Jsame.T @ fsame)) <= 1e-10 * max(1.0, np.max(np.abs(Jsame.T @ fsame)))


@pytest.mark.parametrize("objname", ["chi2", "logl"])
def test_fused_objective_jacobian_and_jtj(objname):
    """'next' row 8f-1: dterms / dlsvec with the row scaling fused into the kernel epilogue, and J^T J / J^T f without
    the Jacobian leaving the device, against the reference objective function evaluated with the reference Map and
    Matrix simulators."""
    from pygsti.data import simulate_data
    from pygsti.objectivefns import objectivefns as _objfns
    from pygsti_b200 import objective as fused
    target = smq1Q_XYI.target_model("full TP")
    # (noise levels chosen so that no lsvec entry sits at 0, where dlsvec = dterms * 0.5 / lsvec is discontinuous:
    #  there even the reference's own Map and Matrix simulators disagree by O(1e3) in J^T J)
    datagen = target.depolarize(op_noise=0.1, spam_noise=0.05)
    circuits = smq1Q_XYI.create_gst_experiment_design(4).all_circuits_needing_data
    ds = simulate_data(datagen, circuits, 1000, seed=1234)
    start = target.depolarize(op_noise=0.06, spam_noise=0.03)
    cls = _objfns.Chi2Function if objname == "chi2" else _objfns.PoissonPicDeltaLogLFunction

    def make(sim):
        m = start.copy(); m.sim = sim
        return cls.create_from(m, ds, circuits, method_names=('lsvec', 'dlsvec', 'dterms'))

    ref = make(MatrixForwardSimulator())      # independent simulator AND independent layout (element order differs)
    ours = make(B200ForwardSimulator())
    same = make(B200ForwardSimulator(fused_objective=False))   # reference objective-function code path on top of the GPU simulator
    v = start.to_vector()
    Jref = ref.dlsvec(v).copy(); fref = ref.lsvec(v).copy()
    Jsame = same.dlsvec(v).copy(); Dsame = same.dterms(v).copy()
    Jf = fused.fused_dlsvec(ours, v).copy()
    Df = fused.fused_dterms(ours, v).copy()
    # (1) fused row scaling == the reference's host-side `jac *= ...` passes, element by element
    assert np.max(np.abs(Jf - Jsame)) <= 1e-11 * max(1.0, np.max(np.abs(Jsame)))
    assert np.max(np.abs(Df - Dsame)) <= 1e-11 * max(1.0, np.max(np.abs(Dsame)))
    # (2) against the reference objective on the reference's analytic Matrix simulator, through quantities that do not
    #     depend on the element order of the layout: J^T J and J^T f (what the LM step consumes)
    JTJ, JTf = fused.fused_jtj(ours, v)
    R = Jref.T @ Jref
    assert np.max(np.abs(JTJ - R)) <= 1e-7 * np.max(np.abs(R))
    assert np.max(np.abs(JTf - Jref.T @ fref)) <= 1e-7 * max(1.0, np.max(np.abs(Jref.T @ fref)))
    assert np.max(np.abs(Jf.T @ Jf - JTJ)) <= 1e-10 * np.max(np.abs(R))


@pytest.mark.parametrize("param", ["full TP", "CPTPLND"])
def test_fused_mle_hessian(param):
    """'next' row 8f-2: the MLE Hessian assembled from rectangles that are evaluated AND reduced on the device
    (b200_hessian_block) against the reference objective function's own `hessian()` on the reference's analytic Matrix
    simulator (objectivefns.py:4892-4990, 1576-1693)."""
    from pygsti.data import simulate_data
    from pygsti.objectivefns import objectivefns as _objfns
    from pygsti_b200 import objective as fused
    target = smq1Q_XYI.target_model(param)
    datagen = smq1Q_XYI.target_model("full TP").depolarize(op_noise=0.1, spam_noise=0.05)
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data[:40]
    ds = simulate_data(datagen, circuits, 1000, seed=1234)
    v = target.to_vector(); rng = np.random.default_rng(5)
    start = target.copy(); start.from_vector(v + 5e-3 * rng.standard_normal(v.size))

    def make(sim):
        m = start.copy(); m.sim = sim
        return _objfns.PoissonPicDeltaLogLFunction.create_from(m, ds, circuits, method_names=('hessian',))

    Href = make(MatrixForwardSimulator()).hessian()
    ours = make(B200ForwardSimulator())
    H = fused.fused_hessian(ours, block_size=17)          # ragged blocks on purpose
    assert H.shape == Href.shape
    assert np.max(np.abs(H - H.T)) <= 1e-9 * np.max(np.abs(Href))
    assert np.max(np.abs(H - Href)) <= 1e-8 * np.max(np.abs(Href))
    Hsame = ours.hessian()                                # reference code path on the GPU simulator (host-side reduction)
    assert np.max(np.abs(H - Hsame)) <= 1e-8 * np.max(np.abs(Href))


def _as_index_list(ix, n):
    return list(range(*ix.indices(n))) if isinstance(ix, slice) else [int(i) for i in ix]


@pytest.mark.parametrize("param", ["full", "full TP"])
def test_device_model_update_follows_from_vector(param):
    """SURVEY 8f rank 3 (first part): with ``device_model_update=True`` only the parameter vector is uploaded once the
    parameters are bound; probabilities, Jacobian and the device copy of the model tensors must follow
    ``model.from_vector`` exactly as the host-packed path does, and a GST-style sequence of parameter updates on ONE
    layout must agree with the reference's Matrix simulator at every step."""
    from pygsti_b200 import calclib, packing
    model = smq2Q_XYCNOT.target_model(param).depolarize(op_noise=0.01, spam_noise=0.01)
    circuits = smq2Q_XYCNOT.create_gst_experiment_design(2).all_circuits_needing_data[:200]
    model.sim = B200ForwardSimulator(device_model_update=True)
    layout = model.sim.create_layout(circuits, array_types=('e', 'ep'))
    ref = model.copy(); ref.sim = MatrixForwardSimulator()
    rlayout = ref.sim.create_layout(circuits, array_types=('e', 'ep'))
    nE, Np = layout.num_elements, model.num_params
    # element correspondence between the two layouts (Map and Matrix layouts order their elements differently)
    ib, ir = [], []
    for c in circuits:
        eb = dict(zip(layout.outcomes(c), _as_index_list(layout.indices(c), nE)))
        er = dict(zip(rlayout.outcomes(c), _as_index_list(rlayout.indices(c), nE)))
        assert set(eb) == set(er)
        for o in eb:
            ib.append(eb[o]); ir.append(er[o])
    ib, ir = np.array(ib), np.array(ir)
    assert ib.size == nE
    rng = np.random.default_rng(5)
    v0 = model.to_vector()
    for step in range(4):
        v = v0 + (0.0 if step == 0 else 0.02) * rng.standard_normal(Np)
        model.from_vector(v); ref.from_vector(v)
        J = np.empty((nE, Np)); p = np.empty(nE); Jr = np.empty((nE, Np)); pr = np.empty(nE)
        model.sim.bulk_fill_dprobs(J, layout, pr_array_to_fill=p)
        ref.sim.bulk_fill_dprobs(Jr, rlayout, pr_array_to_fill=pr)
        assert np.max(np.abs(p[ib] - pr[ir])) <= 1e-12
        assert np.max(np.abs(J[ib] - Jr[ir])) <= 1e-10
        p2 = np.empty(nE)
        model.sim.bulk_fill_probs(p2, layout)                 # probs-only fill after the binding: parameter upload only
        assert np.max(np.abs(p2[ib] - pr[ir])) <= 1e-12
        atom = layout.atoms[0]
        ent = calclib._ATOM_CACHE[atom][calclib.get_context(calclib._device_for(model.sim, atom)).device]
        assert step == 0 or calclib._bound_to(ent, model)      # bound after the first full-Jacobian fill
        mt = packing.pack_model(model, atom, model.dim)
        M_host = np.concatenate([np.ravel(mt.G), np.ravel(mt.rho), np.ravel(mt.E)])
        assert np.max(np.abs(ent["atom"].get_model() - M_host)) <= 1e-15


def test_device_lindblad_members_match_host_path():
    """SURVEY 8f rank 3, second part: with ``device_lindblad=True`` the dense matrices and parameter derivatives of CPTPLND
    members come from the device (b200_lindblad_members); probabilities and Jacobian must equal the host-packed path and the
    reference's Matrix simulator."""
    model = smq1Q_XYI.target_model("CPTPLND")
    rng = np.random.default_rng(9)
    model.from_vector(model.to_vector() + 1e-2 * rng.standard_normal(model.num_params))
    circuits = smq1Q_XYI.create_gst_experiment_design(2).all_circuits_needing_data
    res = {}
    for name, sim in (("dev", B200ForwardSimulator(device_lindblad=True)), ("host", B200ForwardSimulator(device_lindblad=False)),
                      ("matrix", MatrixForwardSimulator())):
        res[name] = _bulk_arrays(model, sim, circuits)
    assert np.max(np.abs(res["dev"][0] - res["host"][0])) <= 1e-12 and np.max(np.abs(res["dev"][0] - res["matrix"][0])) <= 1e-12
    assert np.max(np.abs(res["dev"][1] - res["host"][1])) <= 1e-10 and np.max(np.abs(res["dev"][1] - res["matrix"][1])) <= 1e-10
    # BASELINE config 4's model family, d = 16: the warp-per-row DMMA recursion (k_lind_gen16 / k_lind_dexp16)
    m2 = smq2Q_XYCNOT.target_model("CPTPLND")
    m2.from_vector(m2.to_vector() + 1e-2 * rng.standard_normal(m2.num_params))
    c2 = smq2Q_XYCNOT.create_gst_experiment_design(1).all_circuits_needing_data[:12]
    r2 = {name: _bulk_arrays(m2, sim, c2) for name, sim in (("dev", B200ForwardSimulator(device_lindblad=True)),
                                                             ("host", B200ForwardSimulator(device_lindblad=False)))}
    assert np.max(np.abs(r2["dev"][0] - r2["host"][0])) <= 1e-12
    assert np.max(np.abs(r2["dev"][1] - r2["host"][1])) <= 1e-10
