"""N>1 host logic on CPU: world_size-2 gloo.  Each rank takes its shard of a layout atom (pygsti_b200.dist),
evaluates it (here with the CPU oracle as the stand-in checker -- there is no GPU in this container), and
the shards are reassembled with the same all-gather the GPU path uses with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    from oracle import oracle_np as onp
    c = Case(name)
    a = c.atoms[0]
    local, glob = bd.shard_tables(a["tables"], rank, world)
    p = onp.mapfill_probs(local, a["G"], a["rho"], a["E"])
    J = onp.dprobs_analytic(local, a["G"], a["rho"], a["E"], a["D"])
    gi = torch.from_numpy(glob)
    P = bd.allgather_rows(torch.from_numpy(p), gi, c.n_elements).numpy()
    Jf = bd.allgather_rows(torch.from_numpy(J), gi, c.n_elements).numpy()
    ok = (np.max(np.abs(P - c["probs_map"])) <= 1e-14) and (np.max(np.abs(Jf - c["dprobs_matrix"])) <= 1e-12)
    q.put((rank, bool(ok), int(local.n_elements), int(local.num_state_propagations())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c1_1q_full", "c2_2q_full_sub"])
def test_two_rank_shard_and_allgather(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    work = sorted(r[3] for r in res)
    assert work[1] <= 1.35 * work[0] + 50        # shards are balanced


def test_shards_cover_every_element_once_full_layout():
    sys.path.insert(0, REPO)
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    t = Case("c2_lite_layout").atoms[0]["tables"]
    rows = bd.shard_rows(t, 8)
    assert sum(len(r) for r in rows) == t.n_rows
    L = bd.expanded_lengths(t)
    assert L.sum() == 654388 and L.max() <= 134
    loads = [int(((L[r] + 1)).sum()) for r in rows]
    assert max(loads) <= 1.1 * (sum(loads) / 8)
    local, glob = bd.shard_tables(t, 3, 8, rows=rows[3])
    assert local.n_elements == glob.shape[0] and len(np.unique(glob)) == glob.shape[0]


def _worker_jtj(rank, world, port, name, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    from oracle import oracle_np as onp
    c = Case(name)
    a = c.atoms[0]
    local, glob = bd.shard_tables(a["tables"], rank, world)
    rs = np.random.default_rng(0).uniform(0.5, 1.5, c.n_elements)      # same on every rank
    f = np.random.default_rng(1).standard_normal(c.n_elements)
    J = onp.dprobs_analytic(local, a["G"], a["rho"], a["E"], a["D"]) * rs[glob][:, None]
    jtj = torch.from_numpy(J.T @ J)
    jtf = torch.from_numpy(J.T @ f[glob])
    bd.allreduce_jtj(jtj, jtf)
    Jfull = c["dprobs_matrix"] * rs[:, None]
    e1 = float(np.max(np.abs(jtj.numpy() - Jfull.T @ Jfull)) / np.max(np.abs(Jfull.T @ Jfull)))
    e2 = float(np.max(np.abs(jtf.numpy() - Jfull.T @ f)) / np.max(np.abs(Jfull.T @ f)))
    q.put((rank, e1, e2))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_jtj_allreduce():
    """Sharded J^T J / J^T f: each rank reduces its element shard, one all-reduce adds the partial sums."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker_jtj, args=(r, 2, port, "c1_1q_full", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2 in res:
        assert e1 <= 1e-12 and e2 <= 1e-12, res


def _worker_plan(rank, world, port, name, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    from oracle import oracle_np as onp
    c = Case(name)
    a = c.atoms[0]
    plan = bd.ShardPlan(a["tables"], world)
    slot, n_loc = plan.slot, plan.n_local[rank]
    J = torch.zeros((plan.n_rows_padded, c.num_params), dtype=torch.float64)
    P = torch.zeros(plan.n_rows_padded, dtype=torch.float64)
    local = plan.tables[rank]
    J[rank * slot:rank * slot + n_loc] = torch.from_numpy(onp.dprobs_analytic(local, a["G"], a["rho"], a["E"], a["D"]))
    P[rank * slot:rank * slot + n_loc] = torch.from_numpy(onp.mapfill_probs(local, a["G"], a["rho"], a["E"]))
    bd.allgather_slots(J); bd.allgather_slots(P)
    pos = plan.position
    ok = (np.max(np.abs(P.numpy()[pos] - c["probs_map"])) <= 1e-14) and (np.max(np.abs(J.numpy()[pos] - c["dprobs_matrix"])) <= 1e-12)
    q.put((rank, bool(ok), plan.n_local, [t.num_state_propagations() for t in plan.tables]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c1_1q_full", "c2_2q_full_sub"])
def test_shard_plan_and_in_place_allgather(name):
    """The bench's multi-GPU step on CPU: contiguous prefix-ordered shards (ShardPlan), every rank fills its slot of the sharded
    element axis, ONE in-place all-gather over equal slots, results equal the un-sharded reference arrays
    (serial == N ranks, as test/unit/mpi/run_me_with_mpiexec.py:186-261 checks for the reference)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30700 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker_plan, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_shard_plan_full_layout_keeps_prefix_sharing():
    """8 shards of the C2-lite layout: every element exactly once, balanced, and the shards together need barely more
    propagations WITH prefix sharing than the un-sharded table (the cut costs ~nothing)."""
    sys.path.insert(0, REPO)
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    t = Case("c2_lite_layout").atoms[0]["tables"]
    plan = bd.ShardPlan(t, 8)
    assert sum(plan.n_local) == t.n_elements and len(np.unique(plan.position)) == t.n_elements
    assert plan.position.min() >= 0 and plan.position.max() < plan.n_rows_padded
    assert max(plan.n_local) <= 1.1 * t.n_elements / 8
    for r in range(8):
        sl = plan.element_slice(r)
        assert np.array_equal(np.sort(plan.position[plan.to_original[r]]), np.arange(sl.start, sl.stop))

    def shared_props(tab):                                   # distinct (prep, prefix) pairs = propagations with full prefix sharing
        seen = set()
        for k in range(tab.n_rows):
            ops = tab.row_ops[tab.row_ptr[k]:tab.row_ptr[k + 1]].tolist()
            key = (int(tab.row_prep[k]),)
            for o in ops:
                key = key + (o,)
                seen.add(hash(key))
        return len(seen)
    total = sum(shared_props(tb) for tb in plan.tables)
    assert total <= 1.02 * 38677 + 8 * 134, total            # 38 677 = the reference's own count for the un-sharded lite layout
