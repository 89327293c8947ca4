#!/usr/bin/env python
"""bench.py -- trajectories/sec of the batched grasp-trajectory solver (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl b200|reference]

A "step" is one pass of the hot path over one batch of synthetic problems: all problems of the batch are solved
to convergence (LM iterations of the fused linearise kernel + the block-tridiagonal step kernel).
N=1 workload: BASELINE config C2 (Panda tabletop, 256 candidate grasps x 30 knots, 2000 surface points, 128^3 SDF).
N>1 (launched by torchrun, one rank per GPU): every rank solves its own C2-sized shard (weak scaling; problems are
independent, no collective inside the solve) and the converged trajectories are exchanged with ONE NCCL all-gather.

`value`  = converged trajectories / device time of the solves, inputs already resident in HBM (CUDA events on the
           library's stream, max over ranks).
`e2e`    = the same through the public C-ABI call gto_solve_batch with pinned HOST buffers: H2D of the per-problem
           inputs, solve, D2H of Q/dQ/cost inside the timed region (wall clock between device synchronisations).
`--impl reference` times the CPU oracle port of the same path on the host cores (the reference's CasADi/IPOPT path
cannot be installed offline -- see DESIGN.md).
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
sys.path.insert(0, REPO)

METRIC = "trajectories/sec (batched grasp NLPs to convergence)"
UNIT = "trajectories/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=str, default="C2")
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the batch (debug only; makes the number invalid)")
    ap.add_argument("--cpu-sample", type=int, default=-1, help="problems in the cpu_baseline sample (-1: auto, 0: skip)")
    ap.add_argument("--no-jrows", action="store_true", help="secondary mode: do not materialise Jacobian rows")
    ap.add_argument("--slow-window", type=int, default=0,
                    help="secondary mode: stop a problem (status 'no progress', NOT counted as converged) once its cost fell by <= "
                         "--slow-ftol * f over this many iterations (gto_options.slow_window, at most 16; 0 = off, the contract setting)")
    ap.add_argument("--slow-ftol", type=float, default=1e-3)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum of one full k_linearize_cull launch (256 problems x 28 knots, C2), from the
# ncu --set full capture summarised in profiles/ (bytes per launch; algorithmic bytes of that launch: 413.4 MB)
TRAFFIC_NCU = 445.6e6


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(w, prof, n_fields_per_launch=2):
    """Jacobian rows + residual written once (exact count from the device), both fields read once
    per launch, trajectory read + step written per problem-iteration."""
    t = w.table
    field_bytes = sum(int(np.prod(w.fields[s].cost.shape)) * 4 for s in sorted(w.fields)[:n_fields_per_launch])
    units = prof["problem_iterations"]
    return prof["jrow_bytes"] + prof["linearize_launches_with_work"] * field_bytes + units * 8 * t.nopt * w.batch.T


# --------------------------------------------------------------------------------------------------------------
def time_cpu(w, nsample, nthreads):
    """C restatement of the oracle (oracle/gto_oracle.c, float64, one pthread per core) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import c_oracle

    B = w.batch.B
    idx = np.unique(np.linspace(0, B - 1, min(nsample, B)).astype(int))
    c_oracle.load()
    t0 = time.perf_counter()
    res = c_oracle.solve_workload(w, indices=idx, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return dt, int(np.sum(res["status"] == 0)), len(idx), int(res["threads"])


def run_reference(args):
    """`--impl reference`: the reference's CPU path.  CasADi/IPOPT are not installable offline, so this is the oracle port
    (same reduced problem, projected LM, float64 C) on all host cores; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from grasptrajopt_b200 import workloads as W

    ncpu = os.cpu_count() or 1
    w = W.make_workload(args.config, scale=args.scale)
    B = w.batch.B
    nsample = max(1, min(B, 8 * ncpu))
    for _ in range(max(0, min(args.warmup, 1))):
        time_cpu(w, nsample, ncpu)
    tot_t, tot_conv, threads = 0.0, 0, ncpu
    steps = max(1, args.steps)
    for _ in range(steps):
        dt, conv, n, threads = time_cpu(w, nsample, ncpu)
        tot_t += dt
        tot_conv += conv
    value = tot_conv / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": w.description, "config": args.config.upper()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{nsample} of {B} problems per step, oracle/gto_oracle.c (projected LM, float64), {threads} pthreads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CasADi/IPOPT path not installable offline; its published wall-clock is ~0.1 trajectories/s (BASELINE.md)",
    }
    _emit(line)


# --------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from grasptrajopt_b200 import capi, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    # weak scaling: every rank owns a C2-sized shard generated from its own seed stream
    idx = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}[args.config.upper()]
    scale = args.scale / world if args.config.upper() == "C5" else args.scale  # C5 is the fixed-size (strong) sweep
    w = W.make_workload(args.config, scale=scale, seed=idx + 1000 * rank if world > 1 else None)
    b = w.batch
    if args.no_jrows:
        b.flags |= capi.FLAG_NO_JROWS
    B = b.B
    ctx = capi.GtoContext(local)
    ctx.set_robot(w.table)
    for slot, cf in w.fields.items():
        ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
    opts = capi.default_options(slow_window=min(16, max(0, args.slow_window)), slow_ftol=args.slow_ftol) if args.slow_window > 0 else capi.default_options()

    # pinned host staging for the end-to-end arm
    def pin(a):
        tns = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tns.numpy(), tns

    keep = []
    for name in ("qc", "q_seed", "goal_tf", "base_position"):
        arr, tns = pin(getattr(b, name))
        setattr(b, name, arr)
        keep.append(tns)

    nfl = w.table.nopt * b.T + 2
    gathered = torch.empty((world * B, nfl), dtype=torch.float32, device=f"cuda:{local}") if world > 1 else None

    from grasptrajopt_b200.distributed import DeviceArray

    def exchange():  # ONE all-gather of the converged trajectories, straight from the library's device buffer
        if world > 1:
            ptr, n = ctx.result_device_ptr()
            local_res = torch.as_tensor(DeviceArray(ptr, (B, n)), device=f"cuda:{local}")
            dist.all_gather_into_tensor(gathered, local_res)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----
    ctx.upload_batch(b)
    for _ in range(args.warmup):
        ctx.solve_resident(opts)
        exchange()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, lin_ms, step_ms, launches, conv = 0.0, 0.0, 0.0, 0, 0
    prof_acc = {"jrow_bytes": 0, "problem_iterations": 0, "linearize_launches_with_work": 0, "linearize_launches": 0}
    t0 = time.perf_counter()
    xch_ms = 0.0
    for _ in range(args.steps):
        ctx.solve_resident(opts)
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            exchange()
            e1.record()
            e1.synchronize()
            xch_ms += e0.elapsed_time(e1)
        p = ctx.profile()
        dev_ms += p["solve_ms"]
        lin_ms += p["linearize_ms"]
        step_ms += p["step_ms"]
        launches += p["kernel_launches"]
        for k in prof_acc:
            prof_acc[k] += p[k]
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    last_prof = ctx.profile()
    # cross-check pass (not part of `value`): the same steps with CUDA events recorded around every launch on the library's
    # stream.  The timed region above measures the kernels with in-kernel %globaltimer stamps instead, because an event
    # record between two launches costs ~3 us of stream serialisation (x ~300 launches per solve).
    ev = {"lin_ms": 0.0, "step_ms": 0.0, "solve_ms": 0.0}
    ctx.configure(launch_events=1)
    for _ in range(args.steps):
        ctx.solve_resident(opts)
        p = ctx.profile()
        ev["lin_ms"] += p["linearize_ms"]; ev["step_ms"] += p["step_ms"]; ev["solve_ms"] += p["solve_ms"]
    ctx.configure(launch_events=0)
    sync_all()
    res = ctx.download_batch()
    conv = int(np.sum(res["status"] == capi.STATUS_CONVERGED))
    iters_hist = np.bincount(res["iters"], minlength=1).tolist()
    sc = np.bincount(res["status"], minlength=5)
    status_counts = {"converged": int(sc[0]), "max_iter": int(sc[1]), "nan": int(sc[2]), "stalled": int(sc[3]), "no_progress_at_field_kink": int(sc[4])}
    step_dev_ms = dev_ms + xch_ms  # device time of this rank: solves (library events) + all-gather (torch events)

    # ---- e2e: public API with host buffers ----
    nd = w.table.ndof
    out_pinned = {}
    for k, (shp, dt) in dict(Q=((B, b.T, nd), np.float64), dQ=((B, b.T - 1, nd), np.float64), cost=((B,), np.float64), iters=((B,), np.int32),
                             status=((B,), np.int32)).items():
        tns = torch.empty(shp, dtype=torch.float64 if dt == np.float64 else torch.int32).pin_memory()
        keep.append(tns)
        out_pinned[k] = tns.numpy()
    for _ in range(min(args.warmup, 2)):
        ctx.solve_batch(b, opts, out=out_pinned)
    sync_all()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        r2 = ctx.solve_batch(b, opts, out=out_pinned)
        exchange()
        p = ctx.profile()
        h2d, d2h = p["h2d_bytes"], p["d2h_bytes"]
    sync_all()
    e2e_ms = 1e3 * (time.perf_counter() - t1)
    clocks = sampler.stop()
    conv2 = int(np.sum(r2["status"] == capi.STATUS_CONVERGED))

    # max over ranks
    if world > 1:
        tt = torch.tensor([step_dev_ms, e2e_ms, float(conv), float(conv2), wall_ms], dtype=torch.float64, device=f"cuda:{local}")
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_dev_ms, e2e_ms, wall_ms = float(mx[0]), float(mx[1]), float(mx[4])
        conv_tot, conv2_tot = int(sm[2]), int(sm[3])
    else:
        conv_tot, conv2_tot = conv, conv2

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = algorithmic_bytes(w, prof_acc)
        achieved = alg / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else 0.0
        value = conv_tot * args.steps / (step_dev_ms * 1e-3)
        e2e_value = conv2_tot * args.steps / (e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.config.upper() == "C5" else "weak",
            "vs_baseline": None, "dtype": "f32 (point kernel; TF32 tensor-core J^T J) + f64 (trajectory, step solve)", "data": "synthetic",
            "config": {"workload": w.description, "config": args.config.upper(), "problems_per_gpu": B, "knots": b.T, "surface_points": w.table.npoints,
                       "field": list(next(iter(w.fields.values())).cost.shape), "materialize_jacobian_rows": not args.no_jrows,
                       "l2": "working set per iteration (Jacobian rows, >=0.4 GB) exceeds the 126 MB L2; no explicit flush",
                       "convergence": f"|dq|inf<={opts.tol_step:g} or |proj grad|inf<={opts.tol_grad:g}, max_iter={opts.max_iter}",
                       "converged": conv_tot, "problems": B * world, "iterations_histogram": iters_hist, "status_counts_rank0": status_counts,
                       **({"secondary_mode": f"slow_window={args.slow_window}, slow_ftol={args.slow_ftol}"} if args.slow_window > 0 else {}), "wall_ms_per_step": wall_ms / args.steps,
                       "scale": args.scale},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "timing": "wall clock between device synchronisations around gto_solve_batch with pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_linearize_cull (+ k_item_fk, its per-item pre-pass)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": TRAFFIC_NCU,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": alg / args.steps, "kernel_ms_per_step": lin_ms / args.steps,
                         "kernel_share_of_step": lin_ms / dev_ms if dev_ms else None, "step_kernel_ms_per_step": step_ms / args.steps,
                         "launches_per_step": prof_acc["linearize_launches"] / args.steps,
                         "timing": "sum over the linearise launches of (last warp end - first CTA start), %globaltimer stamps written by the kernels "
                                   "on the library's stream, inside the timed region",
                         "cuda_events_cross_check": {"achieved": alg / (ev["lin_ms"] * 1e-3) / 1e9 if ev["lin_ms"] > 0 else None,
                                                     "kernel_ms_per_step": ev["lin_ms"] / args.steps, "step_kernel_ms_per_step": ev["step_ms"] / args.steps,
                                                     "ms_per_step": ev["solve_ms"] / args.steps,
                                                     "note": "separate pass of the same steps with a CUDA event before/after every launch"},
                         "links_culled_frac": 1.0 - last_prof["links_active"] / max(1, last_prof["links_tested"])},
            "clocks": clocks,
        }
        # CPU baseline on rank 0 at N=1 only: bounded sample of the same workload
        if world == 1 and args.cpu_sample != 0:
            ncpu = os.cpu_count() or 1
            ns = args.cpu_sample if args.cpu_sample > 0 else min(B, 16 * ncpu)
            dt, cconv, n, threads = time_cpu(w, ns, ncpu)
            line["cpu_baseline"] = {"value": cconv / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} of {B} problems of the same workload, oracle/gto_oracle.c (projected LM, float64), {threads} pthreads, {dt:.1f} s"}
        _emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _emit(line: dict) -> None:
    """The contract's ONE JSON line goes to the process's real stdout; everything else written to file descriptor 1 while the
    benchmark runs (e.g. NCCL's version banner when the box exports NCCL_DEBUG) has been diverted to stderr by main()."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # native libraries that print to stdout now land on stderr
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
