#!/usr/bin/env python
"""
bench.py -- BASELINE.json metric: circuit-outcomes/sec of bulk_fill_dprobs (Jacobian + probabilities).

Headline workload (configs[1] of BASELINE.json, the configuration the metric is quoted on): smq2Q_XYCNOT `full`
model (d = 16, Np = 1360), long-sequence GST design maxL = 128 (`lite=False`): 68 335 circuits, 273 340 circuit
outcomes.  The layout tables and model tensors were produced by the reference itself (tests/golden/make_golden.py
c2_full_layout) and are read from tests/golden/c2_full_layout.npz: nothing here needs pyGSTi or /root/reference.

A "step" = one bulk_fill_dprobs over the WHOLE layout = one Jacobian (273 340 x 1360 f64 = 2.97 GB) + the probabilities.

N > 1 (torchrun, one rank per GPU): STRONG scaling of ONE layout.  The layout is cut into N shards the way the reference
cuts it into atoms (pygsti_b200.dist.ShardPlan; reference: maplayout.py:296-303, distlayout.py:326-415); each rank fills
its shard into its slot of the sharded element axis and the timed step ends after ONE in-place NCCL all-gather of the
device-resident Jacobian and probability shards (the reference's gather_local_array -> Allgatherv,
resourceallocation.py:323-329).  `value` = all outcomes / max-over-ranks device time.  The line also carries
  * `multi_gpu.fill_only` / `multi_gpu.allgather` : the two halves of the step, timed separately;
  * `jtj` : the exchange the optimizer actually needs (fill_jtj / fill_jtf, distlayout.py:1220-1359): per-rank J^T J, J^T f of
    the shard (hand-written DMMA kernel, no cuBLAS) + ONE all-reduce of Np^2 + Np doubles -- the Jacobian never moves;
  * `e2e` : the same step through the C ABI with HOST buffers: model upload (H2D) + shard fill + D2H of every rank's rows
    into ONE host array (POSIX shared memory, page-locked by every rank; N PCIe links in parallel);
  * `extra_configs` : BASELINE configs 3 (d = 64 Jacobian, sharded like the headline), 4 (CPTPLND Hessian rectangle) and
    5 (d = 256 probabilities) with their own roofline fraction and CPU baseline.

--impl reference : the reference's own CPU algorithm for the headline path -- forward-difference Jacobian, one prefix-table
pass per parameter (mapforwardsim_calc_densitymx.pyx:290-383) -- executed by the reference's own C++ reps (oracle/_ref)
on all host cores, on a bounded sample of parameters and extrapolated by (Np+1)/(n+1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOAD = "c2_full_layout"
WORKLOAD_DESC = ("smq2Q_XYCNOT full model (d=16, Np=1360), GST design maxL=128 lite=False: 68335 circuits, "
                 "273340 outcomes; bulk_fill_dprobs + probs")
METRIC = "circuit-outcomes/sec (bulk_fill_dprobs)"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel k_accum_trie_d16 (ncu --set full,
# profiles/r02_final_accum_ncu_raw.csv, the round-2 build): 0.176 GB read + 2.919 GB written
NCU_TRAFFIC_BYTES = 3.094e9
INT8_PEAK_POPS = 4.5        # nominal dense int8 tensor rate of a B200 (tcgen05.mma kind::i8)
DMMA_PEAK_TFLOPS = 37.2     # measured on this part with tools/ubench_fp64.cu (mma.sync m8n8k4 f64); DFMA 36.6


def _bf16_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f).get("bf16_tflops", 0.0))
    except Exception:
        return 0.0


def _peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk = float(f[1]); smax = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # timed region shorter than the sampling period: use every sample we have
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); smax = float(f[2])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baselines (the reference's algorithm on the host cores; oracle/ is executed ONLY here and in --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(case, threads, core_seconds=20.0):
    """Reference algorithm (FD Jacobian, pyx:290-383) on the host cores; returns dict for the JSON line.

    Sample = base pass (serial, as in the reference) + n FD parameter passes with n a multiple of the thread
    count (every thread gets the same number of identical full-table passes); the full Jacobian is
    base + Np passes, so  t_full = t_base + t_n * Np / n."""
    from oracle import oracle_c
    oracle_c.build()
    kind = "reference" if os.path.exists(oracle_c.LIB_REF) else "port"
    orc = oracle_c.Oracle(kind)
    a = case.atoms[0]
    t = a["tables"]
    csc = oracle_c.csc_of(a["D"])
    Np = a["D"].n_params
    t0 = time.time()
    orc.mapfill_probs(t, a["G"], a["rho"], a["E"])
    t_base = time.time() - t0
    per_thread = max(1, int(round(core_seconds / max(t_base, 1e-3) / threads)))
    n = min(Np - 80, per_thread * threads)
    lo = 80
    t0 = time.time()
    orc.dprobs_fd(t, a["G"], a["rho"], a["E"], a["D"], p_lo=lo, p_hi=lo + n, eps=1e-7, n_threads=threads, csc=csc)
    dt = time.time() - t0 - t_base           # dprobs_fd runs its own base pass first
    full = t_base + max(dt, 1e-6) * Np / n
    return {"value": case.n_elements / full, "unit": "circuit-outcomes/s", "cores": threads, "kind": kind,
            "sample": "base pass %.3f s + %d of %d FD parameter passes in %.2f s on %d threads over the full "
                      "68335-row prefix table; full Jacobian = base + Np passes (extrapolated x Np/n)"
                      % (t_base, n, Np, dt, threads),
            "seconds_full_jacobian_extrapolated": full}


def _oracle(prefer_ref=True):
    from oracle import oracle_c
    oracle_c.build()
    kind = "reference" if (prefer_ref and os.path.exists(oracle_c.LIB_REF)) else "port"
    return oracle_c.Oracle(kind), kind, oracle_c


def cpu_fd_jacobian_sample(t_sub, G, rho, E, D, threads, n_params_sample):
    """FD Jacobian of the reference on a sub-layout and a block of parameters; returns (outcomes/s extrapolated to all
    parameters, description)."""
    orc, kind, oc = _oracle()
    csc = oc.csc_of(D)
    Np = D.n_params
    t0 = time.time(); orc.mapfill_probs(t_sub, G, rho, E); t_base = time.time() - t0
    n = min(Np, max(threads, n_params_sample // threads * threads))
    t0 = time.time()
    orc.dprobs_fd(t_sub, G, rho, E, D, p_lo=0, p_hi=n, eps=1e-7, n_threads=threads, csc=csc)
    dt = max(time.time() - t0 - t_base, 1e-6)
    full = t_base + dt * Np / n
    return t_sub.n_elements / full, kind, ("%d-circuit sample: base pass %.3f s + %d of %d FD parameter passes in %.2f s on %d "
                                           "threads, extrapolated x Np/n" % (t_sub.n_rows, t_base, n, Np, dt, threads))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from pygsti_b200.fixtures import Case
    case = Case(WORKLOAD)
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(case, threads, core_seconds=2.0)
    t0 = time.time()
    for _ in range(args.steps):
        vals.append(cpu_reference_sample(case, threads, core_seconds=20.0))
    dt = time.time() - t0
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1]); cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "circuit-outcomes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * case.n_elements / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (reference-generated layout + depolarized target model)",
            "config": {"workload": WORKLOAD_DESC, "algorithm": "reference Map simulator: forward-difference Jacobian "
                       "(eps=1e-7), one prefix-table pass per parameter, prefix cache unlimited",
                       "parallelism": "%d host threads over parameters" % threads},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "circuit-outcomes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": dt}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
class Dist:
    """The little of torch.distributed the bench needs (identity at world size 1)."""

    def __init__(self, torch, dist, world, rank):
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def maxv(self, xs):
        t = self.torch.tensor([float(x) for x in xs], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def bcast_obj(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


def timed(D, stream, fn, steps, warmup):
    """ms per call of fn(): `warmup` untimed calls, then exactly `steps` calls bracketed by barrier + synchronize,
    CUDA events on the launching stream, max over ranks."""
    torch = D.torch
    for _ in range(warmup):
        fn()
    D.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    D.barrier()
    return D.max(e0.elapsed_time(e1)) / steps


class HostArray:
    """ONE host array for all ranks: POSIX shared memory created by rank 0, mapped by every rank, and page-locked
    (cudaHostRegister) so that every GPU DMAs its rows straight into it.  Falls back to a private pinned buffer per rank
    (the rows still land in page-locked host memory, but not in one array) if shared memory is unavailable."""

    def __init__(self, D, engine, n_rows, n_cols, tag):
        from multiprocessing import shared_memory
        self.D = D
        self.shm = None
        nbytes = max(int(n_rows) * int(n_cols) * 8, 8)
        name = None
        if D.world > 1:
            if D.rank == 0:
                try:
                    st = os.statvfs("/dev/shm")                      # a tmpfs smaller than the array would SIGBUS on first touch
                    if st.f_bavail * st.f_frsize > nbytes + (256 << 20):
                        self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
                        name = self.shm.name
                except Exception:
                    name = None
            name = D.bcast_obj(name)
            if name is not None and D.rank != 0:
                try:
                    self.shm = shared_memory.SharedMemory(name=name)
                except Exception:
                    self.shm = None
            ok = D.max(0.0 if self.shm is not None else 1.0) == 0.0
            if not ok and self.shm is not None:
                self._close()
        if self.shm is not None:
            self.arr = np.ndarray((n_rows, n_cols), dtype=np.float64, buffer=self.shm.buf)
            self.kind = "one POSIX shared-memory array mapped and page-locked by every rank"
            from pygsti_b200 import _lib
            import ctypes as C
            self._lib = _lib.load()
            self._ptr = C.c_void_p(self.arr.ctypes.data)
            _lib.check(self._lib.b200_host_register(self._ptr, nbytes))
        else:
            self.arr = engine.pinned_empty((n_rows, n_cols))
            self.kind = "pinned host array" if D.world == 1 else "private pinned buffer per rank (shared memory unavailable)"

    def _close(self):
        try:
            self.shm.close()
            if self.D.rank == 0:
                self.shm.unlink()
        except Exception:
            pass
        self.shm = None

    def close(self):
        if self.shm is not None:
            try:
                self._lib.b200_host_unregister(self._ptr)
            except Exception:
                pass
            self.arr = None
            self.D.barrier()
            self._close()


# ------------------------------------------------------------------------------------------------------------------
def bench_c2(args, D, engine, stream, ctx, sampler):
    """Headline: BASELINE config 2, sharded over the ranks."""
    torch, dist = D.torch, D.dist
    from pygsti_b200.fixtures import Case
    from pygsti_b200 import dist as bd
    world, rank = D.world, D.rank
    case = Case(WORKLOAD)
    a = case.atoms[0]
    nE, Np = case.n_elements, case.num_params
    plan = bd.ShardPlan(a["tables"], world)
    slot, n_loc = plan.slot, plan.n_local[rank]
    atom = ctx.upload_atom(plan.tables[rank])
    atom.set_model(a["G"], a["rho"], a["E"])
    atom.set_derivs(a["D"])
    info = atom.info()

    J = torch.empty((plan.n_rows_padded, Np), dtype=torch.float64, device="cuda")     # the sharded layout's element axis
    P = torch.empty(plan.n_rows_padded, dtype=torch.float64, device="cuda")
    J.zero_(); P.zero_()
    Jmine, Pmine = J[rank * slot:(rank + 1) * slot], P[rank * slot:(rank + 1) * slot]

    def fill():
        atom.fill_dprobs_dev(Jmine.data_ptr(), Np, Pmine.data_ptr())

    def gather():
        if world > 1:
            dist.all_gather_into_tensor(J, Jmine)
            dist.all_gather_into_tensor(P, Pmine)

    def step():
        fill(); gather()

    # ---------------- the timed region (value) ----------------
    for _ in range(args.warmup):
        step()
    D.barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = ctx.launch_count
    D.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    D.barrier()
    t_wall1 = time.time()
    launches = (ctx.launch_count - l0) + (2 * args.steps if world > 1 else 0)     # + the two NCCL all-gather kernels per step
    ms_per_step = D.max(e0.elapsed_time(e1)) / args.steps
    value = nE / (ms_per_step * 1e-3)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # parity of what was just timed: every rank holds the whole sharded Jacobian
    pos = plan.position
    st = int(case["probs_map_stride"])
    p_host = P.cpu().numpy()[pos]
    assert np.max(np.abs(p_host[::st] - case["probs_map_sample"])) <= 1e-10
    rows = torch.as_tensor(pos[case["dprobs_matrix_sample_elements"]], device="cuda")
    jerr = float(np.max(np.abs(J[rows].cpu().numpy() - case["dprobs_matrix_sample_rows"])))
    assert jerr <= 1e-10, jerr

    # ---------------- the two halves, separately ----------------
    short = max(5, min(args.steps, 20))
    ms_fill = timed(D, stream, fill, short, 1)
    ms_gather = timed(D, stream, gather, short, 1) if world > 1 else 0.0

    # ---------------- the exchange FUSED into the fill: peer stores over NVLink from the kernel epilogue ----------------
    fused = None
    if world > 1:
        try:
            JP = bd.PeerArray(ctx, plan.n_rows_padded, Np); PP = bd.PeerArray(ctx, plan.n_rows_padded, 1)
            Jt, Pt = JP.tensor(), PP.tensor()
            Jt.zero_(); Pt.zero_()
            D.barrier()
            peers = [r for r in range(world) if r != rank]
            jo = [JP.row_ptr(r, rank * slot) for r in peers]; po = [PP.row_ptr(r, rank * slot) for r in peers]
            jm, pm = JP.row_ptr(rank, rank * slot), PP.row_ptr(rank, rank * slot)
            tick = torch.zeros(1, device="cuda")

            def fused_step():
                atom.fill_dprobs_bcast_dev(jm, Np, pm, jo, po)
                dist.all_reduce(tick)                  # stream-ordered barrier: every rank's stores have been issued and completed

            ms_fused = timed(D, stream, fused_step, short, 2)
            same = bool(torch.equal(Jt, J)) and bool(torch.equal(Pt, P))
            fused = {"ms": ms_fused, "outcomes_per_s": nE / (ms_fused * 1e-3), "bitwise_equal_to_nccl_result": same,
                     "bytes_sent_per_rank": (world - 1) * plan.n_local[rank] * (Np + 1) * 8,
                     "what": "b200_fill_dprobs_bcast_dev: k_accum_trie_d16<PEERS> stores every finished Jacobian block into the local array "
                             "and into the CUDA-IPC-mapped arrays of the %d peers (the unit's 2 KB blocks staged in shared memory and sent as cp.async.bulk stores of the TMA engine over NVLink, one instruction per (outcome, destination) set), then one 4-byte all-reduce "
                             "as the barrier; no separate all-gather pass" % (world - 1)}
            del Jt, Pt
            JP.close(); PP.close()
        except Exception as e:            # CUDA IPC unavailable in this container, ...: the NCCL step above stands
            fused = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    # ---------------- dominant kernel alone (roofline.achieved): CUDA events around the phases ----------------
    ctx.phase_timing(True)
    for _ in range(short):
        fill()
    (ms_prep, ms_chains, ms_accum), n_ph = ctx.phase_ms()
    ctx.phase_timing(False)
    ms_prep, ms_chains, ms_accum = D.maxv([ms_prep / max(n_ph, 1), ms_chains / max(n_ph, 1), ms_accum / max(n_ph, 1)])

    # ---------------- J^T J + J^T f: per-rank DMMA SYRK of the shard + ONE all-reduce ----------------
    rs = torch.from_numpy(np.random.default_rng(0).uniform(0.5, 1.5, nE)[plan.to_original[rank]]).cuda()
    fv = torch.from_numpy(np.random.default_rng(1).standard_normal(nE)[plan.to_original[rank]]).cuda()
    jj = torch.empty(Np * Np + Np, dtype=torch.float64, device="cuda")          # [J^T J | J^T f]: one buffer, one collective

    def jtj_local():
        atom.jtj_dev(jj.data_ptr(), rs.data_ptr(), fv.data_ptr(), jj.data_ptr() + Np * Np * 8)

    def jtj_step():
        jtj_local()
        if world > 1:
            dist.all_reduce(jj)

    ms_jtj = timed(D, stream, jtj_step, max(3, min(args.steps, 10)), 2)
    ms_jtj_local = timed(D, stream, jtj_local, max(3, min(args.steps, 10)), 1)
    # the same contraction on the FP64 tensor cores (round 1/2's k_atb_dmma), for comparison and as a second parity reference
    ctx.set_jtj_mode(0)
    ms_jtj_local_dmma = timed(D, stream, jtj_local, max(3, min(args.steps, 5)), 1)
    jtj_step(); torch.cuda.synchronize()
    jj_dmma = jj.clone()
    ctx.set_jtj_mode(-1)
    # parity: against the same product formed from the gathered, scaled Jacobian by a library GEMM (outside any timed region)
    if world > 1:
        rs_all = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda"); fv_all = torch.zeros_like(rs_all)
        rs_all[rank * slot:rank * slot + n_loc] = rs; fv_all[rank * slot:rank * slot + n_loc] = fv
        dist.all_reduce(rs_all); dist.all_reduce(fv_all)
    else:
        rs_all = torch.zeros(plan.n_rows_padded, dtype=torch.float64, device="cuda"); fv_all = torch.zeros_like(rs_all)
        rs_all[:n_loc] = rs; fv_all[:n_loc] = fv
    step()
    jtj_step()
    torch.cuda.synchronize()
    blk = 32768
    ref = torch.zeros((Np, Np), dtype=torch.float64, device="cuda"); reff = torch.zeros(Np, dtype=torch.float64, device="cuda")
    for r0 in range(0, plan.n_rows_padded, blk):
        Js = J[r0:r0 + blk] * rs_all[r0:r0 + blk, None]
        ref += Js.T @ Js; reff += Js.T @ fv_all[r0:r0 + blk]
    jtj_err = float((jj[:Np * Np].reshape(Np, Np) - ref).abs().max() / ref.abs().max())
    jtf_err = float((jj[Np * Np:] - reff).abs().max() / reff.abs().max())
    assert jtj_err <= 1e-11 and jtf_err <= 1e-11, (jtj_err, jtf_err)
    jtj_err_dmma = float((jj_dmma[:Np * Np].reshape(Np, Np) - ref).abs().max() / ref.abs().max())
    jtj_vs_dmma = float((jj[:Np * Np] - jj_dmma[:Np * Np]).abs().max() / ref.abs().max())
    assert jtj_err_dmma <= 1e-11, jtj_err_dmma
    del jj_dmma
    del ref, reff, rs_all, fv_all
    jtj_flops = float(nE) * Np * (Np + 1)                      # the lower triangle of J^T J over all elements (2 flop per MAC)
    ms_syrk = max(ms_jtj_local - ms_fill, 1e-6)
    ms_syrk_dmma = max(ms_jtj_local_dmma - ms_fill, 1e-6)
    oz_pairs = 36                                               # digit pairs a + b < 8 of the 8-digit splitting

    # ---------------- end to end through the C ABI with HOST buffers (e2e) ----------------
    Jh = HostArray(D, engine, plan.n_rows_padded, Np, "J")
    Ph = HostArray(D, engine, plan.n_rows_padded, 1, "P")
    G_h, rho_h, E_h = (engine.pinned_empty(x.shape) for x in (a["G"], a["rho"], a["E"]))
    G_h[...] = a["G"]; rho_h[...] = a["rho"]; E_h[...] = a["E"]
    Jloc = Jh.arr[rank * slot:rank * slot + n_loc] if Jh.shm is not None else Jh.arr[:n_loc]
    Ploc = (Ph.arr[rank * slot:rank * slot + n_loc] if Ph.shm is not None else Ph.arr[:n_loc]).reshape(-1)

    def e2e_step():
        atom.set_model(G_h, rho_h, E_h)        # H2D of this step's inputs
        atom.fill_dprobs(Jloc, Ploc)           # kernels + D2H of this rank's Jacobian rows and probabilities into host memory

    for _ in range(2):
        e2e_step()
    D.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.time()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.sync()
    D.barrier()
    t_e2e = D.max((time.time() - t0) / e2e_steps)
    if Jh.shm is not None or world == 1:        # one array: check rows that other ranks wrote
        ph = Ph.arr.reshape(-1)[pos]
        assert np.max(np.abs(ph[::st] - case["probs_map_sample"])) <= 1e-10
        je = float(np.max(np.abs(Jh.arr[pos[case["dprobs_matrix_sample_elements"]]] - case["dprobs_matrix_sample_rows"])))
        assert je <= 1e-10, je
    h2d = int((a["G"].size + a["rho"].size + a["E"].size) * 8)
    d2h = int(n_loc * (Np + 1) * 8)

    # e2e of the J^T J path: model upload + shard J^T J + all-reduce + D2H of Np^2 + Np doubles on rank 0
    jj_h = engine.pinned_empty((Np * Np + Np,))
    jj_hm = torch.from_numpy(jj_h)

    def e2e_jtj_step():
        atom.set_model(G_h, rho_h, E_h)
        jtj_step()
        if rank == 0:
            jj_hm.copy_(jj, non_blocking=True)
        torch.cuda.synchronize()

    e2e_jtj_step()
    D.barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        e2e_jtj_step()
    D.barrier()
    t_e2e_jtj = D.max((time.time() - t0) / e2e_steps)
    host_kind = Jh.kind
    Jh.close(); Ph.close()
    atom.free()
    del J, P

    # headline step at N > 1: the exchange fused into the kernel epilogue when it ran and reproduced the NCCL result bit for bit,
    # else fill + ONE ncclAllGather (both are timed with the same rules; both stay in the line)
    ms_nccl_step = ms_per_step
    step_kind = "fill + ONE in-place ncclAllGather"
    if fused and fused.get("bitwise_equal_to_nccl_result") and fused["ms"] < ms_per_step:
        ms_per_step = fused["ms"]; value = nE / (ms_per_step * 1e-3)
        step_kind = "fill with the exchange fused into the kernel epilogue (peer stores over NVLink) + 4-byte all-reduce barrier"
        launches = launches - 2 * args.steps + args.steps            # one barrier all-reduce instead of two all-gathers per step
    peak, peak_src = _peaks()
    alg_bytes = nE * (Np + 1) * 8                                        # whole job
    shard_bytes = max(plan.n_local) * (Np + 1) * 8                       # what the slowest rank's kernel writes
    achieved = shard_bytes / (ms_accum * 1e-3) / 1e9                     # dominant kernel alone, per GPU
    achieved_step = alg_bytes / (ms_per_step * 1e-3) / 1e9 / world       # whole step, per GPU
    ag_bytes = (world - 1) * slot * (Np + 1) * 8                         # received by every rank
    line = {
        "metric": METRIC, "value": value, "unit": "circuit-outcomes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference-generated GST layout + depolarized target model; no dataset needed)",
        "config": {"workload": WORKLOAD_DESC, "derivative": "analytic adjoint (== reference MatrixForwardSimulator)",
                   "kernel": ("k_trie_prepare + k_trie_chains (prefix/suffix-trie chains, heavy-path decomposition) + "
                              "k_accum_trie_d16 (DMMA gather-accumulate, fused Jacobian store)") if info["fused_path"] else "general W.D path",
                   "parallelism": ("shard: ONE layout cut into %d contiguous prefix-ordered shards (ShardPlan), one per GPU; timed step = %s; "
                                   "every rank ends the step holding the whole Jacobian" % (world, step_kind)) if world > 1
                   else "1 GPU, whole layout (no collective)",
                   "l2": "each step writes a %.2f GB Jacobian shard (>> 126 MB L2); no explicit flush needed" % (shard_bytes / 1e9),
                   "dprobs_elements_per_s": value * Np},
        "multi_gpu": {"shard_outcomes": plan.n_local, "slot_rows": slot, "nccl_step_ms": ms_nccl_step, "headline_step": step_kind if world > 1 else None,
                      "fill_only": {"ms": ms_fill, "outcomes_per_s": nE / (ms_fill * 1e-3)},
                      "allgather": {"ms": ms_gather, "bytes_received_per_rank": ag_bytes,
                                    "GBps_per_rank": (ag_bytes / (ms_gather * 1e-3) / 1e9) if ms_gather > 0 else None,
                                    "collective": "ncclAllGather, in place (send buffer = own slot of the receive buffer)" if world > 1 else None},
                      "fused_fill_allgather": fused,
                      "note": "the all-gather moves (N-1)/N of a 2.97 GB Jacobian INTO every GPU over NVLink (<= 900 GB/s), "
                              "while one GPU produces it at ~4 TB/s: for d = 16 the full-Jacobian exchange is NVLink-bound and "
                              "the jtj exchange below is the one that scales"},
        "jtj": {"ms": ms_jtj, "ms_local": ms_jtj_local, "ms_contraction": ms_syrk,
                "fp64_equiv_tflops": jtj_flops / world / (ms_syrk * 1e-3) / 1e12,
                "int8_pops": oz_pairs * jtj_flops / world / (ms_syrk * 1e-3) / 1e15,
                "frac": oz_pairs * jtj_flops / world / (ms_syrk * 1e-3) / 1e15 / INT8_PEAK_POPS,
                "peak": INT8_PEAK_POPS, "peak_unit": "POP/s (int8, dense)",
                "peak_source": "nominal B200 dense int8 rate (MEASURED_PEAKS.json holds no int8 figure; 2 x its measured bf16 burst "
                               "of %.0f TFLOP/s would be %.2f)" % (_bf16_peak(), 2e-3 * _bf16_peak()),
                "fp64_dmma": {"ms_local": ms_jtj_local_dmma, "ms_contraction": ms_syrk_dmma,
                              "tflops": jtj_flops / world / (ms_syrk_dmma * 1e-3) / 1e12,
                              "frac": jtj_flops / world / (ms_syrk_dmma * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS, "peak": DMMA_PEAK_TFLOPS,
                              "peak_source": "measured DMMA rate, tools/ubench_fp64.cu", "parity_jtj_rel": jtj_err_dmma,
                              "what": "k_atb_dmma (hand-written FP64 tensor-core SYRK), b200_ctx_set_jtj_mode(0)"},
                "speedup_vs_fp64_dmma": ms_jtj_local_dmma / ms_jtj_local,
                "allreduce_bytes": (Np * Np + Np) * 8 if world > 1 else 0,
                "outcomes_per_s": nE / (ms_jtj * 1e-3), "e2e_ms": t_e2e_jtj * 1e3, "e2e_outcomes_per_s": nE / t_e2e_jtj,
                "parity": {"jtj_rel": jtj_err, "jtf_rel": jtf_err, "jtj_vs_fp64_dmma_rel": jtj_vs_dmma},
                "what": "scaled Jacobian fill + k_oz_colstats (column exponents + J^T f in one pass) + k_oz_slice (8 int8 digits per "
                        "entry, written as the tensor core's shared-memory image) + k_oz_syrk<8> (tcgen05.mma kind::i8, all 8 level "
                        "accumulators resident in TMEM, digit tiles staged by cp.async.bulk, 36 digit pairs per stage as 12 "
                        "instructions with the column digits stacked along N; exact int32 sums) + k_oz_reduce (FP64, deterministic, "
                        "bitwise symmetric)%s; ms_contraction = everything but the fill; flops = nE*Np*(Np+1), int8 ops = 36 x that" %
                        (" + ONE all-reduce of Np^2+Np doubles" if world > 1 else "")},
        "e2e": {"value": nE / t_e2e, "unit": "circuit-outcomes/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e * 1e3, "host_array": host_kind,
                "d2h_bytes_all_ranks": int(nE * (Np + 1) * 8),
                "path": "b200_atom_set_model + b200_fill_dprobs (C ABI), every rank lands its rows in the host array"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES if world == 1 else None, "peak_source": peak_src,
                     "kernel": "k_accum_trie_d16", "kernel_ms": ms_accum,
                     "algorithmic_bytes_per_launch": shard_bytes,
                     "traffic_source": "ncu --set full, one launch of k_accum_trie_d16 at N=1 (profiles/r02_final_accum_ncu_raw.csv): dram "
                                       "read 0.176 GB + write 2.919 GB = 1.04 x the algorithmic bytes",
                     "step": {"ms": ms_per_step, "achieved": achieved_step, "frac": achieved_step / peak,
                              "phases_ms": {"k_trie_prepare": ms_prep, "k_trie_chains": ms_chains,
                                            "k_accum_trie_d16": ms_accum}},
                     "note": "achieved = 8*nE_shard*(Np+1) bytes / average duration of the dominant kernel on the slowest rank, CUDA "
                             "events on the launching stream (b200_ctx_phase_timing, a separate loop after the timed region); "
                             "'step' is the same quantity over the whole timed step, per GPU"},
        "clocks": clocks,
        "parity_check": {"probs_vs_reference_map_sample": "<=1e-10", "dprobs_vs_reference_matrix_sample_max_abs": jerr},
    }
    return line, case


def bench_c3(args, D, engine, stream, ctx):
    """BASELINE config 3: 3-qubit crosstalk-free `full TP` model (d = 64, Np = 775; tensors from the reference,
    tests/golden/c3_3q_localnoise_sub.npz), 50 000 random circuits of depth U{1..256}, 8 outcomes each,
    bulk_fill_dprobs + probs, sharded over the ranks with one all-gather -- the configuration BASELINE names for 8 GPUs."""
    torch, dist = D.torch, D.dist
    from pygsti_b200 import fixtures as fx
    world, rank = D.world, D.rank
    n_circ = int(os.environ.get("B200_BENCH_C3_CIRCUITS", "50000"))
    c = fx.Case("c3_3q_localnoise_sub"); a = c.atoms[0]
    n_ops, n_eff, Np = a["tables"].n_ops, a["tables"].n_eff, a["D"].n_params
    _, circs = fx.random_layout(64, n_ops, n_eff, n_circ, 256, seed=0, rows=(0, 0))
    blocks = fx.balanced_blocks(circs, world, n_eff)
    t, _ = fx.random_layout(64, n_ops, n_eff, n_circ, 256, seed=0, rows=blocks[rank])
    n_loc_all = [(hi - lo) * n_eff for lo, hi in blocks]
    slot = -(-max(n_loc_all) // 8) * 8
    nE = n_circ * n_eff
    # the model as pyGSTi holds it: every layer = one operation embedded on 1-2 qubits (factor programs) + the derivative map in
    # factor space; the dense map is kept as well (Hessian paths) and serves the comparison run of the dense level-batched kernels
    atom = ctx.upload_atom(t); atom.set_model_factored(a["fm"]); atom.set_derivs(a["D"])
    info = atom.info()
    J = torch.empty((slot * world, Np), dtype=torch.float64, device="cuda"); P = torch.empty(slot * world, dtype=torch.float64, device="cuda")
    Jm, Pm = J[rank * slot:(rank + 1) * slot], P[rank * slot:(rank + 1) * slot]

    def fill():
        atom.fill_dprobs_dev(Jm.data_ptr(), Np, Pm.data_ptr())

    def step():
        fill()
        if world > 1:
            dist.all_gather_into_tensor(J, Jm); dist.all_gather_into_tensor(P, Pm)

    steps = max(3, min(args.steps, 5))
    ms_dense_fill = timed(D, stream, fill, 2, 1)             # dense d x d gates: k_level_gemm_rows sweeps + k_level_accum3
    J_dense_head = Jm[:4096].clone()
    atom.set_derivs_factored(a["Df"])                        # from here on the factored kernels run (k_fj64_forward / k_fj64_backward)
    ms = timed(D, stream, step, steps, 2)
    ms_fill = timed(D, stream, fill, steps, 0) if world > 1 else ms
    dense_vs_factored = float((Jm[:4096] - J_dense_head).abs().max())
    assert dense_vs_factored <= 1e-10, dense_vs_factored
    del J_dense_head
    # the same exchange fused into the kernel: k_fj64_backward stores every finished Jacobian row into the local array and into the
    # peers' arrays (CUDA IPC, NVLink) -- the transfer overlaps the fill instead of following it
    fused = None
    if world > 1:
        try:
            from pygsti_b200 import dist as bd
            JP = bd.PeerArray(ctx, slot * world, Np); PP = bd.PeerArray(ctx, slot * world, 1)
            Jt, Pt = JP.tensor(), PP.tensor()
            Jt.zero_(); Pt.zero_()
            D.barrier()
            peers = [r for r in range(world) if r != rank]
            jo = [JP.row_ptr(r, rank * slot) for r in peers]; po = [PP.row_ptr(r, rank * slot) for r in peers]
            jm, pm = JP.row_ptr(rank, rank * slot), PP.row_ptr(rank, rank * slot)
            tick = torch.zeros(1, device="cuda")

            def fused_step():
                atom.fill_dprobs_bcast_dev(jm, Np, pm, jo, po)
                dist.all_reduce(tick)

            ms_fused = timed(D, stream, fused_step, steps, 2)
            n_loc = n_loc_all[rank]
            same = all(bool(torch.equal(Jt[r * slot:r * slot + n_loc_all[r]], J[r * slot:r * slot + n_loc_all[r]])) for r in range(world))
            fused = {"ms": ms_fused, "bitwise_equal_to_nccl_result": same, "bytes_sent_per_rank": (world - 1) * n_loc * (Np + 1) * 8,
                     "what": "b200_fill_dprobs_bcast_dev: k_fj64_backward writes every Jacobian row to the local array and to the %d peers' "
                             "arrays (256-byte coalesced peer stores over NVLink), then a 4-byte all-reduce as the barrier" % (world - 1)}
            del Jt, Pt
            JP.close(); PP.close()
        except Exception as e:
            fused = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    ms_nccl = ms
    if fused and fused.get("bitwise_equal_to_nccl_result") and fused["ms"] < ms:
        ms = fused["ms"]
    # parity: first circuits of rank 0's shard against the C oracle (checker only)
    par = None
    if rank == 0:
        orc, kind, _ = _oracle(prefer_ref=False)
        sub, _ = fx.random_layout(64, n_ops, n_eff, n_circ, 256, seed=0, rows=(0, 6))
        Jo, po = orc.dprobs_analytic(sub, a["G"], a["rho"], a["E"], a["D"])
        par = {"dprobs_max_abs_vs_oracle_first_6_circuits": float(np.max(np.abs(J[:sub.n_elements].cpu().numpy() - Jo))),
               "probs_max_abs": float(np.max(np.abs(P[:sub.n_elements].cpu().numpy() - po)))}
        assert par["dprobs_max_abs_vs_oracle_first_6_circuits"] <= 1e-10 and par["probs_max_abs"] <= 1e-12, par
    n_prop = sum(len(x) for x in circs)
    fm = a["fm"]
    per_op = np.zeros(n_ops)                                             # multiply-adds of one application of op g through its factors
    for g in range(n_ops):
        for f in range(fm.op_fptr[g], fm.op_fptr[g + 1]):
            per_op[g] += 64 * (4 if fm.f_nq[f] == 1 else 16)
    macs = float(sum(per_op[x].sum() for x in circs))
    chain_flops = 2.0 * macs * (1 + n_eff)                               # forward chain + one backward chain per outcome
    accum_flops = 2.0 * macs * n_eff                                     # acc[a][b] += e(a, r) s(b, r), every outcome
    dense_flops = 2.0 * 64 * 64 * n_prop * (1 + 2 * n_eff)               # the same sweeps and outer products on dense 64 x 64 gates
    alg_bytes = nE * (Np + 1) * 8
    fs_bytes = (n_prop + n_circ) * 64 * 8                                # forward states of every step, written once
    peak, _ = _peaks()
    out = {"config": "BASELINE configs[2]: 3Q XYCNOT full TP (d=64, Np=775), %d random circuits depth U{1..256}, %d outcomes" % (n_circ, nE),
           "ms": ms, "ms_fill_only": ms_fill, "outcomes_per_s": nE / (ms * 1e-3), "dprobs_elements_per_s": nE * Np / (ms * 1e-3),
           "parallelism": ("shard x%d; step = %s; fill + ONE ncclAllGather of %.2f GB shards: %.3f ms" %
                           (world, "fill with the exchange fused into the kernel (peer stores)" if ms < ms_nccl else "fill + ncclAllGather",
                            slot * (Np + 1) * 8 / 1e9, ms_nccl)) if world > 1 else "1 GPU",
           "ms_fill_plus_allgather": ms_nccl, "fused_fill_allgather": fused,
           "kernel": "k_fj64_forward + k_fj64_backward: gates as factor programs (one 4x4 / 16x16 operation embedded on 1-2 of 3 qubits per layer), "
                     "DMMA chain + accumulate per factor, derivative map in factor space; no dense 64 x 64 product, no adjoint table",
           "algorithmic": {"chain_flops": chain_flops, "accumulate_flops": accum_flops, "jacobian_bytes": alg_bytes, "forward_state_bytes": fs_bytes,
                           "dense_equivalent_flops": dense_flops},
           "tflops": (chain_flops + accum_flops) / world / (ms_fill * 1e-3) / 1e12,
           "frac": (chain_flops + accum_flops) / world / (ms_fill * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
           "hbm_frac": (alg_bytes + fs_bytes) / world / (ms_fill * 1e-3) / 1e9 / peak,
           "bound": "latency / issue (dependent 64-vector chain steps of 6-12 DMMA each; ncu: issue-active ~45 %%, DMMA ~34 %%, L1 ~70 %%): "
                    "frac = algorithmic flops of the factored form / FP64 DMMA %.1f TFLOP/s measured, hbm_frac = (Jacobian + forward states) / "
                    "measured HBM peak; HBM floor %.2f ms" % (DMMA_PEAK_TFLOPS, (alg_bytes + fs_bytes) / world / peak / 1e6),
           "dense_level_path": {"ms_fill_only": ms_dense_fill, "tflops": dense_flops / world / (ms_dense_fill * 1e-3) / 1e12,
                                "frac": dense_flops / world / (ms_dense_fill * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
                                "what": "the same Jacobian with every layer as a dense 64 x 64 matrix (k_level_gemm_rows sweeps + k_level_accum3; round 2's "
                                        "first path, still used for models that are not factor programs)",
                                "max_abs_diff_first_4096_rows": dense_vs_factored},
           "speedup_vs_dense_level_path": ms_dense_fill / ms_fill,
           "atom": {k: int(v) for k, v in info.items()}, "parity": par}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sub, _ = fx.random_layout(64, n_ops, n_eff, n_circ, 256, seed=0, rows=(0, 200))
        threads = os.cpu_count() or 1
        v, kind, desc = cpu_fd_jacobian_sample(sub, a["G"], a["rho"], a["E"], a["D"], threads, 2 * threads)
        out["cpu_baseline"] = {"value": v, "unit": "circuit-outcomes/s", "cores": threads, "kind": kind, "sample": desc}
    atom.free()
    del J, P
    return out


def bench_c5(args, D, engine, stream, ctx):
    """BASELINE config 5: d = 256 dense model (14 layer labels, 16 outcomes), 5000 random circuits of depth U{1..128},
    probabilities only -- the FP64 tensor-core roofline run.  Replicated on every rank (2.9 ms of work)."""
    torch = D.torch
    from pygsti_b200 import fixtures as fx
    n_circ = 5000
    G, rho, E = fx.random_dense_model(256, 14, 1, 16, seed=1)
    t, circs = fx.random_layout(256, 14, 16, n_circ, 128, seed=0)
    atom = ctx.upload_atom(t); atom.set_model(G, rho, E)
    P = torch.empty(t.n_elements, dtype=torch.float64, device="cuda")
    ms = timed(D, stream, lambda: atom.fill_probs_dev(P.data_ptr()), max(5, min(args.steps, 20)), 3)
    n_prop = sum(len(x) for x in circs)
    flops = 2.0 * 256 * 256 * n_prop
    par = None
    if D.rank == 0:
        orc, kind, _ = _oracle(prefer_ref=False)
        sub, _ = fx.random_layout(256, 14, 16, n_circ, 128, seed=0, rows=(0, 40))
        po = orc.mapfill_probs(sub, G, rho, E)
        par = {"probs_max_abs_vs_oracle_first_40_circuits": float(np.max(np.abs(P[:po.size].cpu().numpy() - po)))}
        assert par["probs_max_abs_vs_oracle_first_40_circuits"] <= 1e-12, par
    out = {"config": "BASELINE configs[4]: d=256 dense model, 14 layer labels, 5000 random circuits depth U{1..128}, 80000 outcomes, probs only",
           "ms": ms, "outcomes_per_s": D.world * t.n_elements / (ms * 1e-3), "parallelism": "replica per GPU" if D.world > 1 else "1 GPU",
           "algorithmic": {"flops": flops}, "tflops": flops / (ms * 1e-3) / 1e12, "frac": flops / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
           "bound": "tensor (FP64 DMMA %.1f TFLOP/s measured)" % DMMA_PEAK_TFLOPS, "parity": par}
    # The same circuits on a model of the family BASELINE names (4-qubit crosstalk: every layer = small operations embedded on 1-2
    # qubits): the factor programs (device form of OpCRep_Composed / OpCRep_Embedded) against the dense products of the same model
    try:
        from pygsti_b200.packing import FactoredModel, factored_to_dense
        rng = np.random.default_rng(5)
        fptr, f_nq, f_t, f_off, mats, off = [0], [], [], [], [], 0
        targets = [(q,) for q in range(4)] * 2 + [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2)]         # 14 layer labels
        for tg in targets:
            for _ in range(2):                                                                           # ideal gate . local error factor
                k = len(tg)
                small = np.eye(4 ** k) * 0.9 + 0.2 * rng.standard_normal((4 ** k, 4 ** k)) / 2 ** k
                f_nq.append(k); f_t.append(list(tg) + [-1] * (4 - k)); f_off.append(off); mats.append(small.ravel()); off += small.size
            fptr.append(len(f_nq))
        fm = FactoredModel(n_qubits=4, op_fptr=np.asarray(fptr, np.int32), f_nq=np.asarray(f_nq, np.int32),
                           f_targets=np.asarray(f_t, np.int32).reshape(-1, 4), f_moff=np.asarray(f_off, np.int64),
                           mats=np.concatenate(mats), rho=rho, E=E)
        atom.set_model_factored(fm)
        Pf = torch.empty_like(P)
        ms_f = timed(D, stream, lambda: atom.fill_probs_dev(Pf.data_ptr()), max(5, min(args.steps, 20)), 3)
        atom.set_model(factored_to_dense(fm, 256), rho, E)
        ms_d = timed(D, stream, lambda: atom.fill_probs_dev(P.data_ptr()), max(5, min(args.steps, 20)), 3)
        err = float((Pf - P).abs().max())
        assert err <= 1e-11, err
        out["embedded_model"] = {"what": "same 5000 circuits, 14 layer labels each = 2 factors embedded on 1-2 of 4 qubits (28 factors): "
                                         "k_probs_fdmma (factor programs applied as DMMA chain steps, no dense matrices) vs the dense "
                                         "level-batched DMMA path on the densified model", "ms_factored": ms_f, "ms_dense": ms_d, "speedup": ms_d / ms_f,
                                 "outcomes_per_s_factored": D.world * t.n_elements / (ms_f * 1e-3), "max_abs_diff": err}
    except AssertionError:
        raise
    except Exception as e:
        out["embedded_model"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        orc, kind, _ = _oracle()
        sub, _ = fx.random_layout(256, 14, 16, n_circ, 128, seed=0, rows=(0, 250))
        t0 = time.time(); orc.mapfill_probs(sub, G, rho, E); dt = time.time() - t0
        out["cpu_baseline"] = {"value": sub.n_elements / dt, "unit": "circuit-outcomes/s", "cores": 1, "kind": kind,
                               "sample": "250 of the 5000 circuits, one bulk_fill_probs table pass, %.2f s (the reference's pass is serial)" % dt}
    atom.free()
    return out


def bench_c4(args, D, engine, stream, ctx):
    """BASELINE config 4: smq2Q_XYCNOT CPTPLND (d = 16, Np = 1680), GST design maxL = 16 (7860 circuits, 31 440 outcomes; tables,
    tensors, derivative and second-derivative maps from the reference, tests/golden/c4_gst16_layout.npz).  Unit of work = one
    64 x 64 rectangle of the MLE Hessian (`iter_hprobs_by_rectangle` + `_hessian_from_block`): the analytic hprobs block is
    evaluated and reduced on the device.  Rectangles are independent: every rank times the same rectangle (replicas)."""
    from pygsti_b200.fixtures import Case
    path = os.path.join(REPO, "tests", "golden", "c4_gst16_layout.npz")
    if not os.path.exists(path):
        return {"unavailable": "tests/golden/c4_gst16_layout.npz missing"}
    c = Case("c4_gst16_layout"); a = c.atoms[0]
    nE, Np = c.n_elements, c.num_params
    atom = ctx.upload_atom(a["tables"]); atom.set_model(a["G"], a["rho"], a["E"]); atom.set_derivs(a["D"])
    r = c["hess_rects"][0]
    p1, p2 = np.arange(r[0], r[1]), np.arange(r[2], r[3])
    H2 = c.hess_map("H2r0")
    rng = np.random.default_rng(0)
    w_h, w_d = rng.standard_normal(nE), rng.uniform(0.5, 1.5, nE)

    def wall(fn, n, warm=1):
        for _ in range(warm):
            fn()
        D.barrier()
        t0 = time.time()
        for _ in range(n):
            fn()
        ctx.sync()
        return D.max((time.time() - t0) / n)

    Jh = engine.pinned_empty((nE, Np)); ph = np.empty(nE)
    t_j = wall(lambda: atom.fill_dprobs(Jh, ph), 3)
    rows = c["dprobs_matrix_sample_elements"]
    jerr = float(np.max(np.abs(Jh[rows] - c["dprobs_matrix_sample_rows"])))
    hb = atom.hessian_block(p1, p2, w_h, w_d, H2)
    t_h = wall(lambda: atom.hessian_block(p1, p2, w_h, w_d, H2), 2, warm=0)
    # parity of the rectangle on the sampled elements (hprobs materialised only for the check)
    hp = np.empty((nE, p1.size, p2.size))
    atom.fill_hprobs(p1, p2, hp, H2)
    herr = float(np.max(np.abs(hp[rows] - c["hprobs_matrix_rect0_sample_rows"])))
    red = np.einsum("e,eab->ab", w_h, hp) + np.einsum("e,ea,eb->ab", w_d, Jh[:, p1], Jh[:, p2])
    rerr = float(np.max(np.abs(hb - red)) / np.max(np.abs(red)))
    assert jerr <= 1e-10 and herr <= 1e-9 and rerr <= 1e-10, (jerr, herr, rerr)
    n_rect = ((Np + 63) // 64) * ((Np + 63) // 64 + 1) // 2
    out = {"config": "BASELINE configs[3]: smq2Q_XYCNOT CPTPLND (d=16, Np=1680), GST maxL=16: 7860 circuits, %d outcomes" % nE,
           "dprobs": {"e2e_ms": t_j * 1e3, "outcomes_per_s": nE / t_j, "max_abs_vs_reference_matrix_sample": jerr,
                      "path": "general W.D path (D is not a permutation), C ABI with pinned host buffers"},
           "hessian_rectangle": {"block": [int(p1.size), int(p2.size)], "e2e_ms": t_h * 1e3,
                                 "hprobs_elements_per_s": nE * p1.size * p2.size / t_h,
                                 "rectangles_per_s_all_gpus": D.world / t_h,
                                 "full_hessian_s_extrapolated": n_rect * t_h / D.world, "n_rectangles_upper_triangle": n_rect,
                                 "hprobs_max_abs_vs_reference_matrix_sample": herr, "reduction_rel_err": rerr,
                                 "what": "b200_hessian_block: analytic hprobs rectangle (first + second derivative terms) evaluated and "
                                         "reduced with w_h / w_d on the device; only 64 x 64 doubles return"},
           "parallelism": "independent rectangles: replica per GPU" if D.world > 1 else "1 GPU"}
    # on-device model update of the Lindblad members (b200_lindblad_members): synthetic inputs of this config's size -- 7 error generators
    # (5 gates, prep, POVM) of 240 coefficients / 240 parameters each at d = 16, 10 members -- timed through the C ABI (host buffers)
    if D.rank == 0:
        try:
            from types import SimpleNamespace as NS
            rng = np.random.default_rng(11)
            egs = []
            for _ in range(7):
                B = (rng.standard_normal((240, 16, 16)) + 1j * rng.standard_normal((240, 16, 16))) / 16
                egs.append(NS(B_re=np.ascontiguousarray(B.real), B_im=np.ascontiguousarray(B.imag),
                              c=0.02 * (rng.standard_normal(240) + 1j * rng.standard_normal(240)),
                              dc=rng.standard_normal((240, 240)) + 1j * rng.standard_normal((240, 240))))
            mem = [NS(kind="op", errgen=g, static=rng.standard_normal((16, 16))) for g in range(5)] + [NS(kind="rho", errgen=5, static=rng.standard_normal(16))] + \
                  [NS(kind="eff", errgen=6, static=rng.standard_normal(16)) for _ in range(4)]
            ctx.lindblad_members(16, egs, mem)
            ts = []
            for _ in range(3):
                t0 = time.time(); ctx.lindblad_members(16, egs, mem); ts.append(time.time() - t0)
            out["lindblad_members"] = {"e2e_ms": min(ts) * 1e3, "what": "b200_lindblad_members: L, dL, exp(L), Frechet derivatives and composition of 10 members / 7 generators x 240 "
                                       "parameters (d = 16) incl. H2D of the 6.9 MB term tensors and D2H of values + derivatives; the host path "
                                       "(pack_model + pack_derivs through the members' own to_dense / deriv_wrt_params) takes 1.0-1.2 s in the build container"}
        except Exception as e:
            out["lindblad_members"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        orc, kind, _ = _oracle()
        t0 = time.time(); orc.mapfill_probs(a["tables"], a["G"], a["rho"], a["E"]); t_pass = time.time() - t0
        threads = os.cpu_count() or 1
        t_ref = t_pass * (p1.size + 1) * (p2.size + 1) / threads
        out["cpu_baseline"] = {"value": nE * p1.size * p2.size / t_ref, "unit": "hprobs-elements/s", "cores": threads, "kind": kind,
                               "sample": "one prefix-table pass %.3f s on 1 core; the reference's FD-of-FD rectangle (mapforwardsim.py:394-438) is "
                                         "(B1+1)(B2+1) = %d such passes, assumed perfectly parallel over %d threads; the members' expm per FD "
                                         "step is NOT included (conservative)" % (t_pass, (p1.size + 1) * (p2.size + 1), threads)}
    atom.free()
    return out


def bench_plugin(args):
    """`e2e_plugin`: the call a pyGSTi user makes -- `model.sim.bulk_fill_dprobs(array, layout)` with `model.sim =
    B200ForwardSimulator()` -- timed by wall clock next to the stock `MapForwardSimulator` (the reference's Cython path, one process,
    NOT extrapolated) on the same model, circuits and host: smq2Q_XYCNOT `full` (Np = 1360), GST design maxL = 16 (7860 circuits,
    31 440 outcomes; BASELINE.md's "C2-lite maxL=16" probe, 7.3 s in the build container).  Needs the reference install
    (baseline/_ref, which travels to the GPU box); rank 0 at N = 1 only."""
    ref = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "pygsti")):
        return {"unavailable": "baseline/_ref (reference install) not present"}
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore")
    from pygsti.modelpacks import smq2Q_XYCNOT as mp
    from pygsti.forwardsims import MapForwardSimulator
    from pygsti_b200.forwardsim import B200ForwardSimulator
    base = mp.target_model().depolarize(op_noise=0.01, spam_noise=0.01)
    circuits = list(mp.create_gst_experiment_design(16).all_circuits_needing_data)
    out = {"workload": "smq2Q_XYCNOT full (d=16, Np=%d), GST maxL=16: %d circuits; model.sim.bulk_fill_dprobs(J, layout) through pyGSTi, "
                       "host arrays, wall clock" % (base.num_params, len(circuits))}
    res = {}
    for name, sim, reps in (("b200", B200ForwardSimulator(), 4), ("reference", MapForwardSimulator(), 1)):
        m = base.copy(); m.sim = sim
        t0 = time.time(); layout = m.sim.create_layout(circuits, array_types=('e', 'ep')); t_layout = time.time() - t0
        J = np.empty((layout.num_elements, m.num_params))
        ts = []
        for _ in range(reps):
            t0 = time.time(); m.sim.bulk_fill_dprobs(J, layout); ts.append(time.time() - t0)
        # rows by circuit so that the two layouts can be compared
        rows = {}
        for i, c in enumerate(layout.circuits):
            idx, outc = layout.indices_and_outcomes_for_index(i)
            idx = np.arange(idx.start, idx.stop) if isinstance(idx, slice) else np.asarray(idx)
            rows[c] = dict(zip(outc, idx))
        res[name] = (J, rows)
        out[name] = {"create_layout_s": t_layout, "first_call_ms": ts[0] * 1e3, "ms": min(ts) * 1e3, "outcomes": int(layout.num_elements)}
    Jb, rb = res["b200"]; Jr, rr = res["reference"]
    err = 0.0
    for c in circuits[::97]:
        for o, k in rb[c].items():
            err = max(err, float(np.max(np.abs(Jb[k] - Jr[rr[c][o]]))))
    out["ratio"] = out["reference"]["ms"] / out["b200"]["ms"]
    out["max_abs_diff_sampled_circuits"] = err
    out["note"] = ("reference = the stock Cython MapForwardSimulator: forward differences, eps = 1e-7, one prefix-table pass per parameter on one core "
                   "(pyGSTi parallelises over MPI ranks only); the difference to the analytic device Jacobian is that FD's truncation error")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip BASELINE configs 3-5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # the contract is ONE JSON line on stdout: libraries that print there from C (NCCL prints its version line on the first
    # collective when NCCL_DEBUG=VERSION is set in the image) are sent to stderr; the JSON goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from pygsti_b200 import engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = Dist(torch, dist, world, rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = engine.Context(local_rank, stream=stream.cuda_stream)
    sampler = ClockSampler(local_rank)

    line, case = bench_c2(args, D, engine, stream, ctx, sampler)
    if not args.no_extra:
        extra = {}
        for name, fn in (("c3_d64_dprobs", bench_c3), ("c4_cptplnd_hessian", bench_c4), ("c5_d256_probs", bench_c5)):
            torch.cuda.empty_cache()
            try:
                extra[name] = fn(args, D, engine, stream, ctx)
            except AssertionError:
                raise
            except Exception as e:          # an extra workload must not take the headline line down with it
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        line["extra_configs"] = extra
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_sample(case, os.cpu_count() or 1, core_seconds=20.0)
            try:
                line["e2e_plugin"] = bench_plugin(args)
            except AssertionError:
                raise
            except Exception as e:
                line["e2e_plugin"] = {"error": "%s: %s" % (type(e).__name__, e)}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
