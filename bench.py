#!/usr/bin/env python
"""
bench.py -- BASELINE.json metric: circuit-outcomes/sec of bulk_fill_dprobs (Jacobian + probabilities).

Workload (configs[1] of BASELINE.json, the configuration the metric is quoted on): smq2Q_XYCNOT `full`
model (d = 16, Np = 1360), long-sequence GST design maxL = 128 (`lite=False`): 68 335 circuits,
273 340 circuit outcomes.  The layout tables and model tensors were produced by the reference itself
(tests/golden/make_golden.py c2_full_layout) and are read from tests/golden/c2_full_layout.npz, so nothing
here needs pyGSTi or /root/reference at run time.

A "step" = one bulk_fill_dprobs over the whole layout = one Jacobian (273 340 x 1360 f64 = 2.97 GB) plus
the probability vector.

  value   : outcomes/s with inputs resident in HBM and the Jacobian left in HBM (device-timed, CUDA events on
            the launching stream, max over ranks)
  e2e     : the same metric through the C ABI with HOST buffers: every step uploads the model tensors
            (b200_atom_set_model) and lands Jacobian + probs in (pinned) host memory (b200_fill_dprobs)
  --impl reference : the reference's own CPU algorithm for this path -- forward-difference Jacobian, one
            prefix-table pass per parameter (mapforwardsim_calc_densitymx.pyx:290-383) -- executed by the
            reference's own C++ reps (oracle/_ref) on all host cores, on a bounded sample of parameters and
            extrapolated by (Np+1)/(n+1) (every parameter is an identical full table pass).

Multi-GPU (torchrun): the path shards over independent circuits with no data-path collective
(SURVEY.md 8e); every rank simulates its own replica of the layout ("weak" scaling) and `value` is the
whole-job aggregate.
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
# profiles/r01_accum_final_ncu_raw.csv): 0.165 GB read + 2.929 GB written
NCU_TRAFFIC_BYTES = 3.094e9


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
            "ms_per_step": 1e3 * case.n_elements / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (reference-generated layout + depolarized target model)",
            "config": {"workload": WORKLOAD_DESC, "algorithm": "reference Map simulator: forward-difference Jacobian "
                       "(eps=1e-7), one prefix-table pass per parameter, prefix cache unlimited",
                       "parallelism": "%d host threads over parameters" % threads},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "circuit-outcomes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": dt}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from pygsti_b200.fixtures import Case

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    case = Case(WORKLOAD)
    a = case.atoms[0]
    nE, Np = case.n_elements, case.num_params
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = engine.Context(local_rank, stream=stream.cuda_stream)
    atom = ctx.upload_atom(a["tables"])
    atom.set_model(a["G"], a["rho"], a["E"])
    atom.set_derivs(a["D"])
    info = atom.info()

    J = torch.empty((nE, Np), dtype=torch.float64, device="cuda")
    P = torch.empty(nE, dtype=torch.float64, device="cuda")

    # ---------------- device-resident throughput (value) ----------------
    for _ in range(args.warmup):
        atom.fill_dprobs_dev(J.data_ptr(), Np, P.data_ptr())
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = ctx.launch_count
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        atom.fill_dprobs_dev(J.data_ptr(), Np, P.data_ptr())
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = world * nE / (ms_per_step * 1e-3)

    # ---------------- dominant kernel alone (roofline.achieved): CUDA events around the phases, on the launching stream ----
    ctx.phase_timing(True)
    for _ in range(max(5, min(args.steps, 20))):
        atom.fill_dprobs_dev(J.data_ptr(), Np, P.data_ptr())
    (ms_prep, ms_chains, ms_accum), n_ph = ctx.phase_ms()
    ctx.phase_timing(False)
    ph = torch.tensor([ms_prep / max(n_ph, 1), ms_chains / max(n_ph, 1), ms_accum / max(n_ph, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    ms_prep, ms_chains, ms_accum = (float(x) for x in ph.tolist())

    # quick on-device sanity of what was just timed (not part of the timed region)
    st = int(case["probs_map_stride"])
    p_host = P.cpu().numpy()
    assert np.max(np.abs(p_host[::st] - case["probs_map_sample"])) <= 1e-10
    rows = torch.as_tensor(case["dprobs_matrix_sample_elements"], device="cuda")
    jerr = float(np.max(np.abs(J[rows].cpu().numpy() - case["dprobs_matrix_sample_rows"])))
    assert jerr <= 1e-10, jerr

    # ---------------- end-to-end through the C ABI with host buffers (e2e) ----------------
    Jh = engine.pinned_empty((nE, Np))
    Ph = engine.pinned_empty((nE,))
    G_h, rho_h, E_h = (engine.pinned_empty(x.shape) for x in (a["G"], a["rho"], a["E"]))
    G_h[...] = a["G"]; rho_h[...] = a["rho"]; E_h[...] = a["E"]
    for _ in range(2):
        atom.set_model(G_h, rho_h, E_h); atom.fill_dprobs(Jh, Ph)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.time()
    for _ in range(e2e_steps):
        atom.set_model(G_h, rho_h, E_h)        # H2D of this step's inputs
        atom.fill_dprobs(Jh, Ph)               # kernel + D2H of Jacobian and probs into host memory
    ctx.sync()
    t_e2e = torch.tensor([(time.time() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    barrier()
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * nE / float(t_e2e.item())
    assert np.max(np.abs(Ph[::st] - case["probs_map_sample"])) <= 1e-10
    h2d = int((a["G"].size + a["rho"].size + a["E"].size) * 8)
    d2h = int(nE * (Np + 1) * 8)

    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = nE * (Np + 1) * 8
        achieved = alg_bytes / (ms_accum * 1e-3) / 1e9          # dominant kernel alone
        achieved_step = alg_bytes / (ms_per_step * 1e-3) / 1e9  # whole step (prepare + chains + accumulate)
        line = {
            "metric": METRIC, "value": value, "unit": "circuit-outcomes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (reference-generated GST layout + depolarized target model; no dataset needed)",
            "config": {"workload": WORKLOAD_DESC, "derivative": "analytic adjoint (== reference MatrixForwardSimulator)",
                       "kernel": ("k_trie_prepare + k_trie_chains (prefix/suffix-trie chains, heavy-path decomposition) + "
                                  "k_accum_trie_d16 (DMMA gather-accumulate, fused Jacobian store)") if info["fused_path"] else "general W.D path",
                       "parallelism": "dp%d: one replica of the layout per GPU, no data-path collective" % world,
                       "l2": "each step writes a 2.97 GB Jacobian (>> 126 MB L2); no explicit flush needed",
                       "dprobs_elements_per_s": value * Np},
            "e2e": {"value": e2e_value, "unit": "circuit-outcomes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": float(t_e2e.item()) * 1e3,
                    "path": "b200_atom_set_model + b200_fill_dprobs (C ABI), pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES, "peak_source": peak_src,
                         "kernel": "k_accum_trie_d16", "kernel_ms": ms_accum,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "traffic_source": "ncu --set full, one launch of k_accum_trie_d16 (profiles/r01_accum_final_ncu_raw.csv): dram "
                                           "read 0.165 GB + write 2.929 GB = 1.04 x the algorithmic bytes",
                         "step": {"ms": ms_per_step, "achieved": achieved_step, "frac": achieved_step / peak,
                                  "phases_ms": {"k_trie_prepare": ms_prep, "k_trie_chains": ms_chains,
                                                "k_accum_trie_d16": ms_accum}},
                         "note": "achieved = 8*nE*(Np+1) bytes / average duration of the dominant kernel, CUDA events on the "
                                 "launching stream (b200_ctx_phase_timing, a separate loop after the timed region); 'step' "
                                 "is the same quantity over the whole timed step (all three kernels), i.e. value * 10888 B"},
            "clocks": clocks,
            "parity_check": {"probs_vs_reference_map_sample": "<=1e-10", "dprobs_vs_reference_matrix_sample_max_abs": jerr},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_sample(case, os.cpu_count() or 1, core_seconds=20.0)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
