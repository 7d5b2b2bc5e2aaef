#!/usr/bin/env python
"""
Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF (pyGSTi @ 6822f14,
installed into baseline/_ref by oracle/build_ref.py).  The reference ships no golden vectors for this
path, so these are outputs of the reference run in the build container:

  probs_map      <- Cython  MapForwardSimulator.bulk_fill_probs     (pyx:149-287)
  dprobs_map     <- Cython  MapForwardSimulator.bulk_fill_dprobs    (pyx:290-383, forward differences eps=1e-7)
  probs_matrix   <- MatrixForwardSimulator.bulk_fill_probs
  dprobs_matrix  <- MatrixForwardSimulator.bulk_fill_dprobs         (analytic, matrixforwardsim.py:1059-1139)
  hprobs_matrix  <- MatrixForwardSimulator.bulk_fill_hprobs         (analytic, matrixforwardsim.py:1141-1287)

together with the integer layout tables and dense model tensors (pygsti_b200.packing) the engine and
the oracle consume.  All arrays are stored in the element order of the *Map* layout.

Usage:   python tests/golden/make_golden.py [case ...]        (no args = all small cases)
         python tests/golden/make_golden.py c2_full_layout    (the bench workload; ~3 min)
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))

import pygsti  # noqa: E402
from pygsti.forwardsims import MapForwardSimulator, MatrixForwardSimulator  # noqa: E402
from pygsti.circuits import Circuit  # noqa: E402
from pygsti_b200 import packing  # noqa: E402


def _canon_perm(map_layout, other_layout, n_circuits):
    """perm such that other_array[perm] is in map-layout element order."""
    perm = np.empty(map_layout.num_elements, dtype=np.int64)
    # NB: with num_atoms > 1 the Map layout re-orders circuits (SURVEY App. A trap 8), so match by circuit
    assert len(map_layout.circuits) == n_circuits
    for i, circ in enumerate(map_layout.circuits):
        mi, mo = map_layout.indices_and_outcomes_for_index(i)
        oi, oo = other_layout.indices_and_outcomes(circ)
        mi = np.arange(mi.start, mi.stop) if isinstance(mi, slice) else np.asarray(mi)
        oi = np.arange(oi.start, oi.stop) if isinstance(oi, slice) else np.asarray(oi)
        lut = {o: int(ix) for o, ix in zip(oo, oi)}
        for o, ix in zip(mo, mi):
            perm[int(ix)] = lut[o]
    return perm


def build_case(name, model, circuits, num_atoms=None, want_hprobs=False, want_map_fd=True,
               want_matrix=True, extra=None, hess_rects=None):
    """hess_rects = [(slice1, slice2), ...]: store the reference Matrix simulator's Hessian for those parameter
    rectangles only (`hprobs_matrix_rect<i>`, via the rectangle iterator the MLE Hessian uses, forwardsim.py:787-878)
    together with the members' second derivatives (packing.pack_hessians: `a0_H2r<i>_*`; `a0_H2_*` for the full
    Hessian of want_hprobs) so that the GPU tests need no pyGSTi."""
    t0 = time.time()
    d = model.dim
    Np = model.num_params
    circuits = list(circuits)
    mdl = model.copy()
    mdl.sim = MapForwardSimulator(num_atoms=num_atoms)
    assert mdl.sim.calclib.__name__.endswith("mapforwardsim_calc_densitymx"), "need the Cython reference"
    layout = mdl.sim.create_layout(circuits, array_types=('e', 'ep'))
    nE = layout.num_elements
    out = {"n_atoms": np.array(len(layout.atoms)), "dim": np.array(d), "num_params": np.array(Np),
           "n_elements": np.array(nE), "param_vec": mdl.to_vector().copy()}
    for ia, atom in enumerate(layout.atoms):
        pre = "a%d_" % ia
        tabs = packing.pack_atom(atom, d)
        out.update(tabs.to_dict(pre))
        mt = packing.pack_model(mdl, atom, d)
        out[pre + "G"] = mt.G; out[pre + "rho"] = mt.rho; out[pre + "E"] = mt.E
        D = packing.pack_derivs(mdl, atom, d)
        out[pre + "D_rows"] = D.rows; out[pre + "D_cols"] = D.cols; out[pre + "D_vals"] = D.vals
        out[pre + "D_shape"] = np.array([D.n_w, D.n_params])
        es = atom.element_slice
        out[pre + "element_slice"] = np.array([es.start, es.stop])
        fm = packing.pack_model_factored(mdl, atom, d) if d >= 64 else None
        if fm is not None:                # layer operations = products of embedded 1-2 qubit operations: factor programs + factor-space D
            Df = packing.pack_derivs_factored(mdl, atom, d, fm)
            assert np.max(np.abs(packing.factored_to_dense(fm, d) - mt.G)) < 1e-13
            out[pre + "F_n_qubits"] = np.array(fm.n_qubits); out[pre + "F_op_fptr"] = fm.op_fptr; out[pre + "F_nq"] = fm.f_nq
            out[pre + "F_targets"] = fm.f_targets; out[pre + "F_moff"] = fm.f_moff; out[pre + "F_mats"] = fm.mats
            out[pre + "FD_rows"] = Df.rows; out[pre + "FD_cols"] = Df.cols; out[pre + "FD_vals"] = Df.vals
            out[pre + "FD_shape"] = np.array([Df.n_w, Df.n_params])
        for tag, (sl1, sl2) in ([("H2", (None, None))] if want_hprobs else []) + \
                [("H2r%d" % i, r) for i, r in enumerate(hess_rects or [])]:
            H2 = packing.pack_hessians(mdl, atom, d, sl1, sl2)
            out[pre + tag + "_rows"] = H2.rows; out[pre + tag + "_a"] = H2.a; out[pre + tag + "_b"] = H2.b
            out[pre + tag + "_vals"] = H2.vals; out[pre + tag + "_shape"] = np.array([H2.n_w, H2.n1, H2.n2])

    probs_map = np.empty(nE); mdl.sim.bulk_fill_probs(probs_map, layout)
    out["probs_map"] = probs_map
    if want_map_fd:
        dp = np.empty((nE, Np)); mdl.sim.bulk_fill_dprobs(dp, layout)
        out["dprobs_map"] = dp
        out["map_eps"] = np.array(mdl.sim.derivative_eps)

    if want_matrix:
        mm = model.copy()
        mm.sim = MatrixForwardSimulator()
        ml = mm.sim.create_layout(circuits, array_types=('e', 'ep', 'epp') if want_hprobs else ('e', 'ep'))
        perm = _canon_perm(layout, ml, len(circuits))
        pm = np.empty(nE); mm.sim.bulk_fill_probs(pm, ml)
        out["probs_matrix"] = pm[perm]
        dpm = np.empty((nE, Np)); mm.sim.bulk_fill_dprobs(dpm, ml)
        out["dprobs_matrix"] = dpm[perm]
        if want_hprobs:
            hp = np.empty((nE, Np, Np)); mm.sim.bulk_fill_hprobs(hp, ml)
            out["hprobs_matrix"] = hp[perm]
        if hess_rects:
            ml2 = mm.sim.create_layout(circuits, array_types=('e', 'ep', 'epp'))
            perm2 = _canon_perm(layout, ml2, len(circuits))
            for i, rect in enumerate(hess_rects):
                blocks = list(mm.sim.iter_hprobs_by_rectangle(ml2, [rect], False))
                assert len(blocks) == 1
                out["hprobs_matrix_rect%d" % i] = np.array(blocks[0][2])[perm2]
            out["hess_rects"] = np.array([[r[0].start, r[0].stop, r[1].start, r[1].stop] for r in hess_rects])
    if extra:
        out.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s nE=%d Np=%d d=%d atoms=%d  %.1fs  %.2f MB" %
          (name, nE, Np, d, len(layout.atoms), time.time() - t0, os.path.getsize(path) / 1e6))


def _subset(circuits, n, seed=0):
    circuits = list(circuits)
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(len(circuits), size=min(n, len(circuits)), replace=False))
    return [circuits[i] for i in idx]


def _random_circuits(model, n, max_depth, line_labels, seed=0):
    rng = np.random.default_rng(seed)
    prim = list(model.primitive_op_labels)
    out = []
    for _ in range(n):
        L = int(rng.integers(1, max_depth + 1))
        layers = [prim[int(rng.integers(len(prim)))] for _ in range(L)]
        out.append(Circuit(layers, line_labels=line_labels))
    return out


# ---------------------------------------------------------------------------------------------
def case_c1_1q_full():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model().depolarize(op_noise=0.05, spam_noise=0.025)   # as test_forwardsim.py:172,289
    build_case("c1_1q_full", m, mp.create_gst_experiment_design(8).all_circuits_needing_data)


def case_c1_1q_full_atoms3():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    build_case("c1_1q_full_atoms3", m, mp.create_gst_experiment_design(4).all_circuits_needing_data,
               num_atoms=3, want_map_fd=False)


def case_c1_1q_tp():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model('full TP').depolarize(op_noise=0.03, spam_noise=0.01)
    build_case("c1_1q_tp", m, mp.create_gst_experiment_design(4).all_circuits_needing_data)


def case_c1_1q_cptplnd():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model('CPTPLND')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 1e-2 * rng.standard_normal(v.size))
    build_case("c1_1q_cptplnd", m, mp.create_gst_experiment_design(2).all_circuits_needing_data)


def case_c1_1q_hess():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model().depolarize(op_noise=0.05, spam_noise=0.025)
    circs = _subset(mp.create_gst_experiment_design(4).all_circuits_needing_data, 10, seed=1)
    build_case("c1_1q_hess", m, circs, want_hprobs=True)


def case_c1_1q_tp_hess():
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model('full TP').depolarize(op_noise=0.03, spam_noise=0.01)
    circs = _subset(mp.create_gst_experiment_design(4).all_circuits_needing_data, 10, seed=2)
    build_case("c1_1q_tp_hess", m, circs, want_hprobs=True)


def case_c1_1q_cptplnd_hess():
    """1-qubit CPTPLND (every member non-linear in its parameters): full analytic Hessian of the Matrix simulator."""
    from pygsti.modelpacks import smq1Q_XYI as mp
    m = mp.target_model('CPTPLND')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 1e-2 * rng.standard_normal(v.size))
    circs = _subset(mp.create_gst_experiment_design(4).all_circuits_needing_data, 12, seed=4)
    build_case("c1_1q_cptplnd_hess", m, circs, want_hprobs=True, want_map_fd=False)


def case_c4_2q_cptplnd_hess():
    """BASELINE config 4 (2-qubit CPTPLND, Np = 1680; parameters: prep [0,240), POVM [240,480), five gates of 240):
    Hessian rectangles that overlap inside the prep and POVM blocks, inside two gates, and an off-diagonal one."""
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    m = mp.target_model('CPTPLND')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 1e-3 * rng.standard_normal(v.size))
    circs = _subset(mp.create_gst_experiment_design(2).all_circuits_needing_data, 8, seed=5)
    build_case("c4_2q_cptplnd_hess", m, circs, want_map_fd=False, hess_rects=[(slice(225, 265), slice(230, 278)), (slice(700, 740), slice(690, 738)), (slice(225, 265), slice(700, 748))])


def case_c2_2q_full_sub():
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    m = mp.target_model().depolarize(op_noise=0.01, spam_noise=0.01)
    circs = _subset(mp.create_gst_experiment_design(8).all_circuits_needing_data, 24, seed=0)
    build_case("c2_2q_full_sub", m, circs)


def case_c4_2q_cptplnd_sub():
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    m = mp.target_model('CPTPLND')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 1e-3 * rng.standard_normal(v.size))
    circs = _subset(mp.create_gst_experiment_design(4).all_circuits_needing_data, 10, seed=3)
    build_case("c4_2q_cptplnd_sub", m, circs, want_map_fd=False)


def case_c3_3q_localnoise_sub():
    from pygsti.processors import QubitProcessorSpec
    from pygsti.models import modelconstruction as mc
    pspec = QubitProcessorSpec(3, ['Gxpi2', 'Gypi2', 'Gcnot'], geometry='line')
    m = mc.create_crosstalk_free_model(pspec, ideal_gate_type='full TP', ideal_spam_type='full TP')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 0.01 * rng.standard_normal(v.size))
    circs = _random_circuits(m, 8, 24, (0, 1, 2), seed=0)
    build_case("c3_3q_localnoise_sub", m, circs, want_map_fd=False)


def case_c2_full_layout(lite=False, name="c2_full_layout"):
    """The BASELINE.json headline workload: smq2Q_XYCNOT full model, long-sequence GST design maxL=128.
    Only tables + model + sampled reference outputs are stored (the full Jacobian is 2.97 GB)."""
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    t0 = time.time()
    m = mp.target_model().depolarize(op_noise=0.01, spam_noise=0.01)
    circs = list(mp.create_gst_experiment_design(128, lite=lite).all_circuits_needing_data)
    m.sim = MapForwardSimulator()
    layout = m.sim.create_layout(circs, array_types=('e', 'ep'))
    print("layout built: %.1fs  nE=%d" % (time.time() - t0, layout.num_elements))
    atom = layout.atoms[0]
    d = m.dim
    tabs = packing.pack_atom(atom, d)
    mt = packing.pack_model(m, atom, d)
    D = packing.pack_derivs(m, atom, d)
    out = {"n_atoms": np.array(1), "dim": np.array(d), "num_params": np.array(m.num_params),
           "n_elements": np.array(layout.num_elements), "n_circuits": np.array(len(circs))}
    out.update(tabs.to_dict("a0_"))
    out["a0_G"] = mt.G; out["a0_rho"] = mt.rho; out["a0_E"] = mt.E
    out["a0_D_rows"] = D.rows; out["a0_D_cols"] = D.cols; out["a0_D_vals"] = D.vals
    out["a0_D_shape"] = np.array([D.n_w, D.n_params])
    out["a0_element_slice"] = np.array([0, layout.num_elements])
    probs = np.empty(layout.num_elements)
    t1 = time.time(); m.sim.bulk_fill_probs(probs, layout); t_probs = time.time() - t1
    # reference probs: keep every 8th element (float64) + a checksum of all of them
    out["probs_map_stride"] = np.array(8)
    out["probs_map_sample"] = probs[::8].copy()
    out["probs_map_sum"] = np.array(probs.sum())
    out["ref_probs_seconds_1core"] = np.array(t_probs)
    # reference analytic Jacobian rows for a sample of circuits, via the Matrix simulator on just those circuits
    rng = np.random.default_rng(0)
    sample = np.sort(rng.choice(len(circs), size=24, replace=False))
    mm = m.copy(); mm.sim = MatrixForwardSimulator()
    sc = [circs[i] for i in sample]
    ml = mm.sim.create_layout(sc, array_types=('e', 'ep'))
    dpm = np.empty((ml.num_elements, m.num_params)); mm.sim.bulk_fill_dprobs(dpm, ml)
    el_idx = []; rows = []
    for k, ci in enumerate(sample):
        mi, mo = layout.indices_and_outcomes_for_index(int(ci))
        oi, oo = ml.indices_and_outcomes_for_index(k)
        mi = np.arange(mi.start, mi.stop) if isinstance(mi, slice) else np.asarray(mi)
        oi = np.arange(oi.start, oi.stop) if isinstance(oi, slice) else np.asarray(oi)
        lut = {o: int(ix) for o, ix in zip(oo, oi)}
        for o, ix in zip(mo, mi):
            el_idx.append(int(ix)); rows.append(dpm[lut[o]])
    out["dprobs_matrix_sample_elements"] = np.array(el_idx)
    out["dprobs_matrix_sample_rows"] = np.array(rows)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %.1fs  %.2f MB (ref probs %.3fs)" % (name, time.time() - t0, os.path.getsize(path) / 1e6, t_probs))


def case_c4_gst16_layout():
    """BASELINE config 4 at the size SURVEY 8d names: smq2Q_XYCNOT `CPTPLND` model (Np = 1680; parameter vector perturbed by
    1e-3 N(0,1), seed 0), GST design maxL = 16 (7860 circuits, 31 440 outcomes).  Stored: tables, dense model tensors, the
    (general, non-permutation) derivative map, the members' second derivatives for ONE 64 x 64 Hessian rectangle inside a gate's
    parameter block (the unit of work of the MLE Hessian, `iter_hprobs_by_rectangle`), and reference outputs for a sample of
    circuits: probabilities of the Map simulator (every 8th element), and Jacobian rows + that Hessian rectangle from the
    Matrix simulator."""
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    t0 = time.time()
    m = mp.target_model('CPTPLND')
    v = m.to_vector(); rng = np.random.default_rng(0)
    m.from_vector(v + 1e-3 * rng.standard_normal(v.size))
    circs = list(mp.create_gst_experiment_design(16).all_circuits_needing_data)
    m.sim = MapForwardSimulator()
    layout = m.sim.create_layout(circs, array_types=('e', 'ep'))
    atom = layout.atoms[0]
    d = m.dim
    print("layout built: %.1fs  nE=%d circuits=%d" % (time.time() - t0, layout.num_elements, len(circs)))
    tabs = packing.pack_atom(atom, d)
    mt = packing.pack_model(m, atom, d)
    D = packing.pack_derivs(m, atom, d)
    rect = (slice(700, 764), slice(690, 754))
    H2 = packing.pack_hessians(m, atom, d, rect[0], rect[1])
    out = {"n_atoms": np.array(1), "dim": np.array(d), "num_params": np.array(m.num_params),
           "n_elements": np.array(layout.num_elements), "n_circuits": np.array(len(circs)), "param_vec": m.to_vector().copy()}
    out.update(tabs.to_dict("a0_"))
    out["a0_G"] = mt.G; out["a0_rho"] = mt.rho; out["a0_E"] = mt.E
    out["a0_D_rows"] = D.rows; out["a0_D_cols"] = D.cols; out["a0_D_vals"] = D.vals
    out["a0_D_shape"] = np.array([D.n_w, D.n_params])
    out["a0_element_slice"] = np.array([0, layout.num_elements])
    out["a0_H2r0_rows"] = H2.rows; out["a0_H2r0_a"] = H2.a; out["a0_H2r0_b"] = H2.b; out["a0_H2r0_vals"] = H2.vals
    out["a0_H2r0_shape"] = np.array([H2.n_w, H2.n1, H2.n2])
    out["hess_rects"] = np.array([[rect[0].start, rect[0].stop, rect[1].start, rect[1].stop]])
    probs = np.empty(layout.num_elements)
    t1 = time.time(); m.sim.bulk_fill_probs(probs, layout); t_probs = time.time() - t1
    out["probs_map_stride"] = np.array(8)
    out["probs_map_sample"] = probs[::8].copy()
    out["probs_map_sum"] = np.array(probs.sum())
    out["ref_probs_seconds_1core"] = np.array(t_probs)
    sample = np.sort(rng.choice(len(circs), size=12, replace=False))
    mm = m.copy(); mm.sim = MatrixForwardSimulator()
    sc = [circs[i] for i in sample]
    ml = mm.sim.create_layout(sc, array_types=('e', 'ep', 'epp'))
    dpm = np.empty((ml.num_elements, m.num_params)); mm.sim.bulk_fill_dprobs(dpm, ml)
    blocks = list(mm.sim.iter_hprobs_by_rectangle(ml, [rect], False))
    hp = np.array(blocks[0][2])
    el_idx, rows, hrows = [], [], []
    for k, ci in enumerate(sample):
        mi, mo = layout.indices_and_outcomes_for_index(int(ci))
        oi, oo = ml.indices_and_outcomes_for_index(k)
        mi = np.arange(mi.start, mi.stop) if isinstance(mi, slice) else np.asarray(mi)
        oi = np.arange(oi.start, oi.stop) if isinstance(oi, slice) else np.asarray(oi)
        lut = {o: int(ix) for o, ix in zip(oo, oi)}
        for o, ix in zip(mo, mi):
            el_idx.append(int(ix)); rows.append(dpm[lut[o]]); hrows.append(hp[lut[o]])
    out["dprobs_matrix_sample_elements"] = np.array(el_idx)
    out["dprobs_matrix_sample_rows"] = np.array(rows)
    out["hprobs_matrix_rect0_sample_rows"] = np.array(hrows)
    path = os.path.join(HERE, "c4_gst16_layout.npz")
    np.savez_compressed(path, **out)
    print("c4_gst16_layout: %.1fs  %.2f MB (ref probs %.3fs, H2 nnz %d)" % (time.time() - t0, os.path.getsize(path) / 1e6, t_probs, H2.rows.size))


def case_c2_lite_layout():
    case_c2_full_layout(lite=True, name="c2_lite_layout")


SMALL = ["c1_1q_full", "c1_1q_full_atoms3", "c1_1q_tp", "c1_1q_cptplnd", "c1_1q_hess", "c1_1q_tp_hess",
         "c2_2q_full_sub", "c4_2q_cptplnd_sub", "c3_3q_localnoise_sub", "c1_1q_cptplnd_hess", "c4_2q_cptplnd_hess"]

if __name__ == "__main__":
    names = sys.argv[1:] or SMALL
    for n in names:
        globals()["case_" + n]()
