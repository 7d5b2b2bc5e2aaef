"""
Multi-GPU host logic: one process per GPU (torch.distributed; NCCL over NVLink on the B200 box, gloo on CPU).

The path shards over independent circuits (layout atoms / circuits are independent units that
need only the small model tensors), so there is NO collective on the data path.  The single exchange step
the reference has -- gathering element shards when the host optimizer wants whole arrays
(``gather_local_array`` -> ``Allgatherv``, pygsti/baseobjs/resourceallocation.py:323-329) -- is
``allgather_rows`` below.

``shard_tables`` cuts a layout atom into ``world`` independent sub-atoms balanced by propagation count.
Prefix-cache links are resolved (every row of a shard carries its full op sequence) so a shard never
depends on rows owned by another rank; the reference's equivalent is the prefix-tree cut of
``PrefixTable.find_splitting_new`` (pygsti/layouts/prefixtable.py:154-290), which duplicates shared
prefixes on both sides of a cut for the same reason.
"""
import numpy as np

from .packing import AtomTables


def expanded_lengths(t: AtomTables):
    """Depth of every row after following cache links (no sequences materialised)."""
    n = t.n_rows
    L = np.zeros(n, dtype=np.int64)
    slot_len = np.zeros(max(t.cache_size, 1), dtype=np.int64)
    rem = (t.row_ptr[1:] - t.row_ptr[:-1]).astype(np.int64)
    for k in range(n):
        l = rem[k] + (slot_len[t.row_istart[k]] if t.row_istart[k] >= 0 else 0)
        L[k] = l
        if t.row_icache[k] >= 0:
            slot_len[t.row_icache[k]] = l
    return L


def shard_rows(t: AtomTables, world: int):
    """Greedy balanced assignment of rows to ranks by (depth + 1) * n_outcomes work; returns list of row-index arrays."""
    L = expanded_lengths(t)
    nout = (t.out_ptr[1:] - t.out_ptr[:-1]).astype(np.int64)
    work = (L + 1) * (1 + 2 * nout)          # forward chain + per-outcome backward chain and accumulation
    order = np.argsort(-work, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.empty(t.n_rows, dtype=np.int64)
    # longest-processing-time-first in blocks (keeps it O(n log n) and deterministic)
    for idx in order:
        r = int(np.argmin(load))
        owner[idx] = r
        load[r] += work[idx]
    return [np.flatnonzero(owner == r) for r in range(world)]


def shard_tables(t: AtomTables, rank: int, world: int, rows=None):
    """Sub-atom of ``t`` owned by ``rank``: (AtomTables with locally renumbered elements, global element index
    of every local element).  Rows are independent (no cache links)."""
    if rows is None:
        rows = shard_rows(t, world)[rank]
    rows = np.asarray(rows, dtype=np.int64)
    # expand only what is needed: follow cache links through producing rows
    slot_row = {}
    prep = np.empty(t.n_rows, dtype=np.int32)
    parent = np.full(t.n_rows, -1, dtype=np.int64)
    for k in range(t.n_rows):
        if t.row_istart[k] >= 0:
            parent[k] = slot_row[int(t.row_istart[k])]
            prep[k] = prep[parent[k]]
        else:
            prep[k] = t.row_prep[k]
        if t.row_icache[k] >= 0:
            slot_row[int(t.row_icache[k])] = k
    ops, ptr = [], [0]
    out_eff, out_el, optr = [], [], [0]
    for k in rows:
        chain = []
        r = int(k)
        while r >= 0:
            chain.append(t.row_ops[t.row_ptr[r]:t.row_ptr[r + 1]])
            r = int(parent[r])
        seq = np.concatenate(chain[::-1]) if chain else np.zeros(0, np.int32)
        ops.append(seq)
        ptr.append(ptr[-1] + len(seq))
        a, b = t.out_ptr[k], t.out_ptr[k + 1]
        out_eff.append(t.out_eff[a:b]); out_el.append(t.out_el[a:b]); optr.append(optr[-1] + (b - a))
    glob = np.concatenate(out_el).astype(np.int64) if out_el else np.zeros(0, np.int64)
    n_local = glob.shape[0]
    local = AtomTables(
        dim=t.dim, n_ops=t.n_ops, n_rho=t.n_rho, n_eff=t.n_eff, n_elements=int(n_local), cache_size=0,
        row_dest=np.arange(len(rows), dtype=np.int32), row_istart=np.full(len(rows), -1, np.int32),
        row_icache=np.full(len(rows), -1, np.int32), row_prep=prep[rows].astype(np.int32),
        row_ptr=np.asarray(ptr, dtype=np.int32),
        row_ops=(np.concatenate(ops).astype(np.int32) if ops else np.zeros(0, np.int32)),
        out_ptr=np.asarray(optr, dtype=np.int32),
        out_eff=(np.concatenate(out_eff).astype(np.int32) if out_eff else np.zeros(0, np.int32)),
        out_el=np.arange(n_local, dtype=np.int32))
    return local, glob


def expanded_rows(t: AtomTables):
    """(prep [n_rows], list of full op sequences) with every cache link followed."""
    slot_row = {}
    prep = np.empty(t.n_rows, dtype=np.int32)
    seqs = [None] * t.n_rows
    for k in range(t.n_rows):
        rem = t.row_ops[t.row_ptr[k]:t.row_ptr[k + 1]]
        if t.row_istart[k] >= 0:
            src = slot_row[int(t.row_istart[k])]
            prep[k] = prep[src]
            seqs[k] = np.concatenate([seqs[src], rem]) if len(rem) else seqs[src]
        else:
            prep[k] = t.row_prep[k]
            seqs[k] = rem
        if t.row_icache[k] >= 0:
            slot_row[int(t.row_icache[k])] = k
    return prep, seqs


class ShardPlan:
    """A layout atom cut into ``world`` shards the way the reference cuts a layout into atoms
    (``MapCOPALayout`` with ``num_atoms = world``: pygsti/layouts/maplayout.py:296-303 -> ``PrefixTable.find_splitting_new``,
    prefixtable.py:154-290): rows are ordered by circuit prefix (prep, ops...) and cut into CONTIGUOUS blocks, so that the
    circuits of a shard keep sharing their prefixes, balanced by an HBM-write + gather work estimate.  Like the reference's
    atoms, every shard owns a contiguous ``element_slice`` of the sharded layout's element axis
    (distlayout.py:326-415): shard r owns global rows [r * slot, r * slot + n_local[r]) where ``slot`` = the largest shard
    rounded up (equal slots make the exchange ONE in-place ncclAllGather; the <= few padding rows at the end of a slot
    belong to no element).  ``to_original[r]`` maps a shard's local element index to the element index of the un-sharded
    atom, ``position`` is the inverse map (original element -> row of the sharded element axis)."""

    def __init__(self, t: AtomTables, world: int, work_per_outcome=77.0):
        self.world = int(world)
        prep, seqs = expanded_rows(t)
        order = sorted(range(t.n_rows), key=lambda k: (int(prep[k]), seqs[k].tolist()))
        nout = (t.out_ptr[1:] - t.out_ptr[:-1]).astype(np.float64)
        L = np.array([len(s) for s in seqs], dtype=np.float64)
        work = nout * (work_per_outcome + L)              # stores ~ outcomes, table gathers ~ outcomes x depth
        cw = np.cumsum(work[order]) if t.n_rows else np.zeros(0)
        total = cw[-1] if t.n_rows else 0.0
        cuts = [0]
        for r in range(1, self.world):
            cuts.append(int(np.searchsorted(cw, total * r / self.world, side="left")))
        cuts.append(t.n_rows)
        cuts = np.maximum.accumulate(np.asarray(cuts))
        self.rows = [np.asarray(order[cuts[r]:cuts[r + 1]], dtype=np.int64) for r in range(self.world)]
        self.tables, self.to_original = [], []
        for r in range(self.world):
            rows = self.rows[r]
            ops = [seqs[k] for k in rows]
            ptr = np.zeros(len(rows) + 1, np.int64)
            if len(rows):
                ptr[1:] = np.cumsum([len(o) for o in ops])
            oe = [t.out_eff[t.out_ptr[k]:t.out_ptr[k + 1]] for k in rows]
            ol = [t.out_el[t.out_ptr[k]:t.out_ptr[k + 1]] for k in rows]
            optr = np.zeros(len(rows) + 1, np.int64)
            if len(rows):
                optr[1:] = np.cumsum([len(o) for o in oe])
            glob = np.concatenate(ol).astype(np.int64) if len(rows) else np.zeros(0, np.int64)
            n_local = int(glob.shape[0])
            self.tables.append(AtomTables(
                dim=t.dim, n_ops=t.n_ops, n_rho=t.n_rho, n_eff=t.n_eff, n_elements=n_local, cache_size=0,
                row_dest=np.arange(len(rows), dtype=np.int32), row_istart=np.full(len(rows), -1, np.int32),
                row_icache=np.full(len(rows), -1, np.int32), row_prep=prep[rows].astype(np.int32),
                row_ptr=ptr.astype(np.int32),
                row_ops=(np.concatenate(ops).astype(np.int32) if len(rows) else np.zeros(0, np.int32)),
                out_ptr=optr.astype(np.int32),
                out_eff=(np.concatenate(oe).astype(np.int32) if len(rows) else np.zeros(0, np.int32)),
                out_el=np.arange(n_local, dtype=np.int32)))
            self.to_original.append(glob)
        self.n_local = [tb.n_elements for tb in self.tables]
        self.n_elements = int(t.n_elements)
        self.slot = max(1, -(-max(self.n_local) // 8) * 8)        # rows per slot (multiple of 8 rows)
        self.n_rows_padded = self.slot * self.world
        self.position = np.full(self.n_elements, -1, dtype=np.int64)
        for r in range(self.world):
            self.position[self.to_original[r]] = r * self.slot + np.arange(self.n_local[r])

    def element_slice(self, rank):
        return slice(rank * self.slot, rank * self.slot + self.n_local[rank])


def allgather_slots(full, group=None):
    """ONE in-place all-gather over equal slots: ``full`` is the (world * slot, ...) array of the sharded layout, of which
    this rank has filled its own slot ``full[rank * slot:(rank + 1) * slot]``; on return every rank holds every slot.
    NCCL: a single ncclAllGather with the send buffer inside the receive buffer (no staging, no copies); gloo on CPU.
    The analogue of ``gather_local_array`` -> ``Allgatherv`` (pygsti/baseobjs/resourceallocation.py:323-329)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    slot = full.shape[0] // world
    assert slot * world == full.shape[0] and full.is_contiguous()
    dist.all_gather_into_tensor(full, full[rank * slot:(rank + 1) * slot], group=group)
    return full


def allgather_rows(local, global_index, n_total, group=None):
    """All-gather row shards of a (n_local, ...) torch tensor into the full (n_total, ...) tensor on every rank.

    ``global_index`` (int64 tensor, n_local) gives the destination row of each local row.  Shards may be
    uneven: they are padded to the largest shard for the collective (one ``all_gather`` = one NCCL
    allgather on the GPU box; the analogue of the reference's Allgatherv)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes) if sizes else 0
    tail = tuple(local.shape[1:])
    pad = torch.zeros((nmax,) + tail, dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    ipad = torch.full((nmax,), -1, dtype=torch.int64, device=local.device)
    ipad[:local.shape[0]] = global_index.to(local.device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    ibufs = [torch.empty_like(ipad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    dist.all_gather(ibufs, ipad, group=group)
    out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
    for r in range(world):
        n = sizes[r]
        if n:
            out[ibufs[r][:n]] = bufs[r][:n]
    return out


def allreduce_jtj(jtj, jtf=None, group=None):
    """Sum the per-rank partial ``J^T J`` (and ``J^T f``) of element shards over all ranks, in place: ONE all-reduce of
    ``n_params^2 + n_params`` doubles (14.8 MB at BASELINE size) instead of gathering the 2.97 GB Jacobian -- what the
    reference does when every rank adds its block into the shared ``jtj`` (``fill_jtj`` / ``allreduce_sum``,
    pygsti/layouts/distlayout.py:1220-1359, pygsti/baseobjs/resourceallocation.py:331-376).  NCCL on the GPU box
    (tensors filled by ``Atom.jtj_dev``), gloo on CPU."""
    import torch
    import torch.distributed as dist
    if jtf is None:
        dist.all_reduce(jtj, op=dist.ReduceOp.SUM, group=group)
        return jtj, None
    n = jtj.shape[0]
    flat = torch.cat([jtj.reshape(-1), jtf.reshape(-1)])          # one collective for both
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    jtj.copy_(flat[:n * n].reshape(n, n))
    jtf.copy_(flat[n * n:])
    return jtj, jtf


class PeerArray:
    """A (n_rows, n_cols) float64 device array that exists on EVERY rank and that every rank can store into directly: each
    rank allocates its copy through the engine (``b200_peer_alloc``: cudaMalloc + CUDA IPC handle), the handles are exchanged
    once (``all_gather_object``) and every rank maps the copies of its peers (``b200_peer_open``, NVLink peer access).  With
    ``Atom.fill_dprobs_bcast_dev`` the Jacobian kernel of rank r then writes rank r's slot of the sharded element axis into all
    copies while it produces it -- the all-gather of ``allgather_slots`` without a separate pass.
    ``sync()`` = stream synchronisation + barrier: afterwards every copy is complete."""

    def __init__(self, ctx, n_rows, n_cols, group=None):
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nbytes = max(self.n_rows * self.n_cols * 8, 8)
        ptr, handle = ctx.peer_alloc(self.nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.ptrs = [ptr if r == self.rank else ctx.peer_open(handles[r]) for r in range(self.world)]
        dist.barrier(group=group)

    @property
    def __cuda_array_interface__(self):
        shape = (self.n_rows, self.n_cols) if self.n_cols > 1 else (self.n_rows,)
        return {"shape": shape, "typestr": "<f8", "data": (self.ptrs[self.rank], False), "version": 3, "strides": None}

    def tensor(self):
        """torch view of THIS rank's copy (no copy)."""
        import torch
        return torch.as_tensor(self, device="cuda:%d" % self.ctx.device)

    def row_ptr(self, owner, row):
        """Device pointer (valid in this process) of row ``row`` inside the copy held by rank ``owner``."""
        return self.ptrs[owner] + int(row) * self.n_cols * 8

    def sync(self):
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)

    def close(self):
        import torch.distributed as dist
        if self.ptrs is None:
            return
        self.ctx.sync()
        dist.barrier(group=self.group)                   # nobody stores into a copy that is about to go away
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.ctx.peer_close(p)
        dist.barrier(group=self.group)
        self.ctx.peer_free(self.ptrs[self.rank])
        self.ptrs = None
