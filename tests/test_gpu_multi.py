"""Multi-GPU path on real devices (skipped with < 2 GPUs): ranks shard ONE layout (ShardPlan), fill their shard on their GPU
through the C ABI (device-buffer entry points) into their slot of the sharded element axis, and reassemble it with ONE
in-place NCCL all-gather; the sharded J^T J / J^T f take one NCCL all-reduce.  Serial == N ranks, as the reference checks in
test/unit/mpi/run_me_with_mpiexec.py:186-261."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd, engine
    c = Case(name)
    a = c.atoms[0]
    plan = bd.ShardPlan(a["tables"], world)
    slot, n_loc = plan.slot, plan.n_local[rank]
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = engine.Context(rank, stream=stream.cuda_stream)
    at = ctx.upload_atom(plan.tables[rank]); at.set_model(a["G"], a["rho"], a["E"]); at.set_derivs(a["D"])
    Np = c.num_params
    J = torch.zeros((plan.n_rows_padded, Np), dtype=torch.float64, device="cuda")
    P = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda")
    at.fill_dprobs_dev(J[rank * slot:].data_ptr(), Np, P[rank * slot:].data_ptr())
    bd.allgather_slots(J); bd.allgather_slots(P)
    torch.cuda.synchronize()
    pos = plan.position
    st = int(c["probs_map_stride"])
    e1 = float(np.max(np.abs(P.cpu().numpy()[pos][::st] - c["probs_map_sample"])))
    rows = torch.as_tensor(pos[c["dprobs_matrix_sample_elements"]], device="cuda")
    e2 = float(np.max(np.abs(J[rows].cpu().numpy() - c["dprobs_matrix_sample_rows"])))
    # sharded J^T J / J^T f: per-rank b200_jtj_dev on the element shard + ONE NCCL all-reduce
    rs = np.random.default_rng(0).uniform(0.5, 1.5, c.n_elements); fv = np.random.default_rng(1).standard_normal(c.n_elements)
    glob = plan.to_original[rank]
    d_rs = torch.from_numpy(rs[glob]).cuda(); d_f = torch.from_numpy(fv[glob]).cuda()
    jtj = torch.empty((Np, Np), dtype=torch.float64, device="cuda"); jtf = torch.empty(Np, dtype=torch.float64, device="cuda")
    at.jtj_dev(jtj.data_ptr(), d_rs.data_ptr(), d_f.data_ptr(), jtf.data_ptr())
    bd.allreduce_jtj(jtj, jtf)
    torch.cuda.synchronize()
    rs_s = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda"); fv_s = torch.zeros_like(rs_s)
    rs_s[torch.from_numpy(pos).cuda()] = torch.from_numpy(rs).cuda(); fv_s[torch.from_numpy(pos).cuda()] = torch.from_numpy(fv).cuda()
    Js = J * rs_s[:, None]
    ref_jtj = Js.T @ Js; ref_jtf = Js.T @ fv_s
    e3 = float((jtj - ref_jtj).abs().max() / ref_jtj.abs().max())
    e4 = float((jtf - ref_jtf).abs().max() / ref_jtf.abs().max())
    # the same exchange FUSED into the fill: every rank's kernel stores its rows into all copies of a peer-mapped array
    JP = bd.PeerArray(ctx, plan.n_rows_padded, Np); PP = bd.PeerArray(ctx, plan.n_rows_padded, 1)
    Jt, Pt = JP.tensor(), PP.tensor()
    Jt.zero_(); Pt.zero_()
    torch.cuda.synchronize(); dist.barrier()
    peers = [r for r in range(world) if r != rank]
    at.fill_dprobs_bcast_dev(JP.row_ptr(rank, rank * slot), Np, PP.row_ptr(rank, rank * slot),
                             [JP.row_ptr(r, rank * slot) for r in peers], [PP.row_ptr(r, rank * slot) for r in peers])
    JP.sync()
    e5 = float((Jt - J).abs().max()); e6 = float((Pt - P).abs().max())         # bitwise the NCCL result
    del Jt, Pt
    JP.close(); PP.close()
    q.put((rank, e1, e2, e3, e4, e5, e6))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shard_fill_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "c2_lite_layout", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, e1, e2, e3, e4, e5, e6 in res:
        assert e1 <= 1e-12 and e2 <= 1e-10 and e3 <= 1e-10 and e4 <= 1e-10 and e5 == 0.0 and e6 == 0.0, res


def test_single_process_two_gpus_concurrent():
    """One Python process driving 2 GPUs (`B200ForwardSimulator(num_atoms=4, devices=[0, 1])`): the layout-level
    fills run the atoms of the two GPUs concurrently; results equal the one-GPU, one-atom fill circuit by circuit, and the
    concurrent dprobs is faster than the same layout filled atom after atom on one GPU is not asserted (timing is reported)."""
    import time
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    REF = os.path.join(REPO, "baseline", "_ref")
    if os.path.isdir(os.path.join(REF, "pygsti")) and REF not in sys.path:
        sys.path.insert(0, REF)
    pytest.importorskip("pygsti", reason="reference install (baseline/_ref) not present")
    from pygsti.modelpacks import smq2Q_XYCNOT
    from pygsti_b200.forwardsim import B200ForwardSimulator
    model = smq2Q_XYCNOT.target_model().depolarize(op_noise=0.01, spam_noise=0.01)
    circuits = smq2Q_XYCNOT.create_gst_experiment_design(8).all_circuits_needing_data
    out = {}
    for tag, kw in (("one", dict(num_atoms=1)), ("two", dict(num_atoms=4, devices=[0, 1]))):
        m = model.copy()
        m.sim = B200ForwardSimulator(**kw)
        layout = m.sim.create_layout(circuits, array_types=('e', 'ep'))
        p = np.empty(layout.num_elements); dp = np.full((layout.num_elements, m.num_params), np.nan)
        p_only = np.empty(layout.num_elements)
        m.sim.bulk_fill_probs(p_only, layout)
        m.sim.bulk_fill_dprobs(dp, layout, pr_array_to_fill=p)          # first call: uploads
        t0 = time.time(); m.sim.bulk_fill_dprobs(dp, layout, pr_array_to_fill=p); dt = time.time() - t0
        assert np.max(np.abs(p - p_only)) <= 1e-13
        cur = {}
        for i, c in enumerate(layout.circuits):
            idx, outs = layout.indices_and_outcomes_for_index(i)
            cur[c] = (p[idx].copy(), dp[idx].copy())
        out[tag] = (cur, dt, len(layout.atoms))
    assert out["two"][2] == 4
    if getattr(out["two"][0], "keys", None):
        for c in out["one"][0]:
            assert np.max(np.abs(out["one"][0][c][0] - out["two"][0][c][0])) <= 1e-13
            assert np.max(np.abs(out["one"][0][c][1] - out["two"][0][c][1])) <= 1e-12
    print("bulk_fill_dprobs through pyGSTi: 1 GPU %.1f ms, 2 GPUs (4 atoms, concurrent) %.1f ms" % (out["one"][1] * 1e3, out["two"][1] * 1e3))
