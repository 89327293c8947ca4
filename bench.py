#!/usr/bin/env python
"""bench.py -- trajectories/sec of the batched grasp-trajectory solver (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl b200|reference]

A "step" is one pass of the hot path over one batch of synthetic problems: all problems of the batch are solved
to convergence (LM iterations of the fused linearise kernel + the block-tridiagonal step kernel).
N=1 workload: BASELINE config C2 (Panda tabletop, 256 candidate grasps x 30 knots, 2000 surface points, 128^3 SDF).
N>1 (launched by torchrun, one rank per GPU): every rank solves a C2-sized batch (weak scaling with equal work per GPU: the
same batch on every rank; --distinct-shards 1 for differently seeded ones; problems are independent, no collective inside the
solve) and the converged trajectories are exchanged with ONE NCCL all-gather per batch.

`value`  = converged trajectories / device time of the K solves, inputs already resident in HBM (CUDA events around the
           timed region, max over ranks).  The K steps are issued with `--in-flight` batches at a time (continuous batching at
           batch granularity: every batch is an independent planning call, as in the reference's per-scene loop).
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
    ap.add_argument("--extras", type=int, default=1,
                    help="also time BASELINE configs[4] (C5, 16384 problems, strong sweep: 16384/N per rank) and configs[3] (C4, 4096 problems, "
                         "4096/N per rank) and report them under config.extra (the headline value stays the C2 line); 0 = skip")
    ap.add_argument("--fused", type=int, default=0, help="secondary mode: k_solve_fused (one persistent CTA per problem)")
    ap.add_argument("--distinct-shards", type=int, default=0,
                    help="N > 1: 1 = every rank generates its own C2-sized shard from its own seed (round-1 behaviour).  Default 0: every rank "
                         "solves the SAME batch, the N = 1 workload -- weak scaling with exactly equal work per GPU.  The differently seeded "
                         "shards differ in difficulty (on one GPU alone they run at 30.7 .. 40.0 k traj/s, tools/exp_rank_shards.py; rank 0's is "
                         "the easiest), so with them the step time of an N-GPU run is that of the hardest shard and says nothing about scaling")
    ap.add_argument("--shard-seed-of-rank", type=int, default=-1,
                    help="diagnostic (1 GPU): solve the shard that this rank of a multi-GPU run generates (seed = config index + 1000 * rank), to "
                         "separate the difficulty of the shards from the cost of running N ranks")
    ap.add_argument("--in-flight", type=int, default=6,
                    help="batches in flight: the K steps are issued from this many host threads, each with its own context (own stream, own "
                         "resident copy of the batch), so that the latency-bound tail of one batch (a handful of problems still iterating) "
                         "overlaps the head of the next; 1 = one batch at a time (its latency is reported as solve_ms_per_step either way)")
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


def ncu_traffic(config, no_jrows, scale):
    """dram__bytes_read.sum + dram__bytes_write.sum of one full k_linearize_cull launch, from the kept `ncu --set full` capture of
    THIS build (profiles/r2_traffic.json, written by tools/ncu_summary.py from the committed CSV).  Only valid for the
    configuration it was captured on; anything else reports null."""
    p = os.path.join(REPO, "profiles", "r2_traffic.json")
    if not os.path.exists(p) or no_jrows or scale != 1.0:
        return None, None
    with open(p) as fh:
        d = json.load(fh)
    if d.get("config") != config.upper():
        return None, None
    return float(d["dram_bytes_per_launch"]), d


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(w, prof):
    """SURVEY.md section 8(d), per problem-iteration: 4(nopt+1)(T*P + 6Pg) Jacobian rows + residual written once (exact row count
    from the device: knots 0,1 are linearised only once) + 8*N^3 / B_scene (both fields of a scene read once per batch-iteration,
    amortised over the B_scene problems that share the scene) + 8*nopt*T (trajectory read, step written)."""
    t = w.table
    b = w.batch
    slots, counts = np.unique(np.concatenate([b.field_all, b.field_obs]), return_counts=True)
    per_scene = {}
    for s_, c_ in zip(slots, counts):
        if s_ >= 0:
            per_scene[int(s_)] = int(np.prod(w.fields[int(s_)].cost.shape)) * 4 / c_  # bytes of this field per problem that reads it
    field_per_problem = float(np.mean([per_scene.get(int(b.field_all[i]), 0.0) + per_scene.get(int(b.field_obs[i]), 0.0) for i in range(b.B)]))
    units = prof["problem_iterations"]
    return prof["jrow_bytes"] + units * (field_per_problem + 8 * t.nopt * b.T)


# --------------------------------------------------------------------------------------------------------------
def config_block(w, args, opts, world, **more):
    """The `config` object of the JSON line -- identical keys for the b200 and the reference arm."""
    b = w.batch
    d = {"workload": w.description, "config": args.config.upper(), "problems_per_gpu": b.B, "knots": int(b.T), "surface_points": int(w.table.npoints),
         "field": list(next(iter(w.fields.values())).cost.shape), "materialize_jacobian_rows": not args.no_jrows,
         "l2": "working set per iteration (Jacobian rows, >=0.4 GB) exceeds the 126 MB L2; no explicit flush",
         "convergence": f"|dq|inf<={opts.tol_step:g} under damping<={opts.lambda_conv:g} or |proj grad|inf<={opts.tol_grad:g}, max_iter={opts.max_iter}",
         "problems": b.B * world, "scale": args.scale}
    d.update(more)
    return d


def status_dict(status):
    sc = np.bincount(status, minlength=5)
    return {"converged": int(sc[0]), "max_iter": int(sc[1]), "nan": int(sc[2]), "stalled": int(sc[3]), "rests_on_field_kink": int(sc[4])}


def time_cpu(w, idx, nthreads, options=None):
    """C restatement of the oracle (oracle/gto_oracle.c, float64, one pthread per core) on problems `idx` of the workload."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import c_oracle

    c_oracle.load()
    t0 = time.perf_counter()
    res = c_oracle.solve_workload(w, indices=idx, nthreads=nthreads, options=options)
    dt = time.perf_counter() - t0
    return dt, res


def run_reference(args):
    """`--impl reference`: the reference's CPU path.  CasADi/IPOPT are not installable offline, so this is the oracle
    port (same reduced problem, same projected LM, float64 C) on all host cores.  Same configuration and the same FULL batch
    per step as the b200 arm for C1/C2/C5; for the Fetch configs (seconds per problem-batch on the host) a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import c_oracle
    from grasptrajopt_b200 import workloads as W

    ncpu = os.cpu_count() or 1
    w = W.make_workload(args.config, scale=args.scale)
    B = w.batch.B
    full = args.config.upper() in ("C1", "C2")
    idx = np.arange(B) if full else np.unique(np.linspace(0, B - 1, min(B, 8 * ncpu)).astype(int))
    opts = c_oracle.default_options()
    for _ in range(max(0, min(args.warmup, 1))):
        time_cpu(w, idx, ncpu, opts)
    tot_t, tot_conv, threads = 0.0, 0, ncpu
    steps = max(1, args.steps)
    for _ in range(steps):
        dt, res = time_cpu(w, idx, ncpu, opts)
        tot_t += dt
        tot_conv += int(np.sum(res["status"] == 0))
        threads = int(res["threads"])
    value = tot_conv / tot_t
    sample = (f"all {B} problems per step" if full else f"{len(idx)} of {B} problems per step") + \
        f", oracle/gto_oracle.c (projected LM, float64), {threads} pthreads on {ncpu} host cores"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_block(w, args, opts, 1, converged=int(np.sum(res["status"] == 0)), iterations_histogram=np.bincount(res["iters"], minlength=1).tolist(),
                               status_counts_rank0=status_dict(res["status"]), wall_ms_per_step=1e3 * tot_t / steps,
                               batches_in_flight=1, host_blocking_sync=False, shards="one batch", solve_ms_per_step=1e3 * tot_t / steps,
                               solve_ms_note="wall clock of one batch on the host cores", xch_ms_per_step=0.0, exchange="none (rank 0 only)",
                               host_cores=ncpu, problems_per_step=int(len(idx))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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
    from grasptrajopt_b200.distributed import DeviceArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE is {world}: launch N > 1 with "
                         f"`python -m torch.distributed.run --nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...`")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    cfg = args.config.upper()
    idx = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}[cfg]
    # headline: weak scaling, every rank solves a full-size batch of the configuration (the same batch on every rank by default:
    # equal work per GPU; --distinct-shards 1: a batch generated from the rank's own seed stream)
    shard_rank = args.shard_seed_of_rank if (world == 1 and args.shard_seed_of_rank >= 0) else rank
    distinct = (world > 1 and args.distinct_shards) or args.shard_seed_of_rank >= 0
    w = W.make_workload(cfg, scale=args.scale, seed=idx + 1000 * shard_rank if distinct else None)
    NF = max(1, args.in_flight)
    ctxs = [capi.GtoContext(local) for _ in range(NF)]
    # more solving host threads on the box than cores (8 ranks x 4 in flight on 32 cores): sleep in the polls instead of spinning
    blocking = int(NF * world >= (os.cpu_count() or 1))
    for c in ctxs:
        c.configure(fused=args.fused, blocking_sync=blocking)
    ctx = ctxs[0]
    xch_lock = threading.Lock()
    opts = capi.default_options(slow_window=min(15, max(0, args.slow_window)), slow_ftol=args.slow_ftol) if args.slow_window > 0 else capi.default_options()
    keep = []

    def pin(a):
        tns = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(tns)
        return tns.numpy()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(w, steps, warmup, e2e=True, events_pass=True):
        """Times `steps` solves of workload `w` on this rank (inputs resident), the all-gather of the results, and -- optionally --
        the end-to-end arm through gto_solve_batch with pinned host buffers.  Returns a dict of per-rank measurements."""
        b = w.batch
        if args.no_jrows:
            b.flags |= capi.FLAG_NO_JROWS
        B = b.B
        nf = max(1, min(NF, steps))
        for c in ctxs[:nf]:
            c.set_robot(w.table)
            for slot, cf in w.fields.items():
                c.set_field(slot, cf.cost, cf.origin, cf.pitch)
        for name in ("qc", "q_seed", "goal_tf", "base_position"):
            setattr(b, name, pin(getattr(b, name)))
        nfl = w.table.nopt * b.T + 2
        gathered = [torch.empty((world * B, nfl), dtype=torch.float32, device=dev) for _ in range(nf)] if world > 1 else None
        xs = torch.cuda.Stream(device=dev) if world > 1 else None

        staging = [torch.empty((B, nfl), dtype=torch.float32, device=dev) for _ in range(nf)] if world > 1 else None
        pending = [None] * nf
        xch_events = []

        def exchange(i):
            """ONE all-gather of the packed trajectories of the solve that just finished on context i (gto_solve_resident is
            synchronous), issued on a side stream and NOT waited for: the worker goes on with its next solve while NCCL moves the
            data.  The library's result buffer is copied (device to device, 0.2 MB) into a per-context staging tensor first and
            only that copy is waited for, so the next solve may overwrite the buffer.  Collectives of the in-flight batches are
            serialised by a lock (identical calls, so any order matches across ranks); timed with events on the side stream."""
            if world == 1:
                return
            with xch_lock:
                if pending[i] is not None:
                    pending[i].synchronize()  # the previous all-gather out of this staging tensor has finished
                ptr, n = ctxs[i].result_device_ptr()
                local_res = torch.as_tensor(DeviceArray(ptr, (B, n)), device=dev)
                e0, e1, ec = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
                with torch.cuda.stream(xs):
                    staging[i].copy_(local_res)
                    ec.record()
                    e0.record()
                    dist.all_gather_into_tensor(gathered[i], staging[i])
                    e1.record()
                ec.synchronize()
                pending[i] = e1
                xch_events.append((e0, e1))

        def xch_ms_total():
            t = sum(e0.elapsed_time(e1) for e0, e1 in xch_events)
            xch_events.clear()
            return t

        def in_flight(fn, nsteps):
            """Runs fn(i) nsteps times in total from nf host threads (thread i owns context i); returns the CUDA-event time of
            the region: both events are recorded on an otherwise idle stream after a device synchronisation, and every solve
            has completed on the device when its call returns, so the interval covers all of the work."""
            share = [nsteps // nf + (1 if i < nsteps % nf else 0) for i in range(nf)]
            errs = []

            def work(i):
                try:
                    torch.cuda.set_device(local)
                    for _ in range(share[i]):
                        fn(i)
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)

            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            th = [threading.Thread(target=work, args=(i,)) for i in range(nf)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            e1.record()
            sync_all()
            wall = 1e3 * (time.perf_counter() - t0)
            if errs:
                raise errs[0]
            return e0.elapsed_time(e1), wall

        for c in ctxs[:nf]:
            c.upload_batch(b)
        for i in range(nf):
            for _ in range(warmup):
                ctxs[i].solve_resident(opts)
                exchange(i)
        sync_all()
        xch_ms_total()
        m = {"dev_ms": 0.0, "lin_ms": 0.0, "step_ms": 0.0, "xch_ms": 0.0, "launches": 0, "solve_ms": 0.0, "in_flight": nf,
             "prof": {"jrow_bytes": 0, "problem_iterations": 0, "linearize_launches_with_work": 0, "linearize_launches": 0}}
        acc_lock = threading.Lock()

        def one_step(i):
            ctxs[i].solve_resident(opts)
            exchange(i)
            p = ctxs[i].profile()
            with acc_lock:
                m["solve_ms"] += p["solve_ms"]; m["lin_ms"] += p["linearize_ms"]; m["step_ms"] += p["step_ms"]; m["launches"] += p["kernel_launches"]
                for k in m["prof"]:
                    m["prof"][k] += p[k]

        m["dev_ms"], m["wall_ms"] = in_flight(one_step, steps)
        m["xch_ms"] = xch_ms_total()
        m["last_prof"] = ctx.profile()
        res = ctx.download_batch()  # the result of the last timed solve of context 0
        m["conv"] = int(np.sum(res["status"] == capi.STATUS_CONVERGED))
        m["status"] = res["status"]; m["iters"] = res["iters"]
        m["ev"] = None
        if events_pass and not args.fused:
            # kernel durations with CUDA events recorded around every launch on the library's stream (separate pass of the same
            # steps: an event record between two launches costs ~3 us of stream serialisation, x ~300 launches per solve, so the
            # pass is slower than the timed region; the in-kernel %globaltimer stamps of the timed region are reported beside it)
            ev = {"lin_ms": 0.0, "step_ms": 0.0, "solve_ms": 0.0}
            ctx.configure(launch_events=1)
            for _ in range(steps):
                ctx.solve_resident(opts)
                p = ctx.profile()
                ev["lin_ms"] += p["linearize_ms"]; ev["step_ms"] += p["step_ms"]; ev["solve_ms"] += p["solve_ms"]
            ctx.configure(launch_events=0)
            sync_all()
            m["ev"] = ev
        m["e2e_ms"] = None
        if e2e:
            nd = w.table.ndof
            out_pinned = {}
            for k, (shp, dt) in dict(Q=((B, b.T, nd), np.float64), dQ=((B, b.T - 1, nd), np.float64), cost=((B,), np.float64), iters=((B,), np.int32),
                                     status=((B,), np.int32)).items():
                tns = torch.empty(shp, dtype=torch.float64 if dt == np.float64 else torch.int32).pin_memory()
                keep.append(tns)
                out_pinned[k] = tns.numpy()
            outs = [out_pinned]
            for _ in range(1, nf):
                o2 = {}
                for k, v in out_pinned.items():
                    tns = torch.empty(v.shape, dtype=torch.float64 if v.dtype == np.float64 else torch.int32).pin_memory()
                    keep.append(tns)
                    o2[k] = tns.numpy()
                outs.append(o2)
            for i in range(nf):
                for _ in range(min(warmup, 2)):
                    ctxs[i].solve_batch(b, opts, out=outs[i])
            last = {}

            def one_e2e(i):
                last[i] = ctxs[i].solve_batch(b, opts, out=outs[i])
                exchange(i)

            _, m["e2e_ms"] = in_flight(one_e2e, steps)
            xch_ms_total()
            p = ctx.profile()
            m["h2d"], m["d2h"] = p["h2d_bytes"], p["d2h_bytes"]
            m["conv2"] = int(np.sum(last[0]["status"] == capi.STATUS_CONVERGED))
        return m

    def over_ranks(vals_max, vals_sum):
        if world == 1:
            return vals_max, vals_sum
        tmx = torch.tensor(vals_max, dtype=torch.float64, device=dev)
        tsm = torch.tensor(vals_sum, dtype=torch.float64, device=dev)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
        return tmx.tolist(), tsm.tolist()

    sampler = ClockSampler(local)
    sampler.start()
    m = measure(w, args.steps, args.warmup)
    clocks = sampler.stop()
    step_dev_ms = m["dev_ms"]  # device time of this rank for the K steps (CUDA events around the region: solves + all-gathers)
    (step_dev_ms_mx, e2e_ms_mx, wall_ms_mx, xch_ms_mx, solve_ms_mx), (conv_tot, conv2_tot) = over_ranks(
        [step_dev_ms, m["e2e_ms"], m["wall_ms"], m["xch_ms"], m["solve_ms"]], [float(m["conv"]), float(m["conv2"])])

    # ---- BASELINE configs[4] (C5 strong sweep) and configs[3] (C4 sharded over the ranks): secondary lines under config.extra ----
    extra = {}
    if args.extras and cfg == "C2" and args.scale == 1.0:
        for xcfg, total in (("C5", 16384), ("C4", 4096)):
            wx = W.make_workload(xcfg)  # the SAME problems on every rank ...
            lo, hi = W.shard_range(wx.batch.B, rank, world)  # ... of which this rank solves a contiguous shard (strong scaling)
            wx.batch = W.slice_batch(wx.batch, lo, hi)
            used = set(int(v) for v in np.concatenate([wx.batch.field_all, wx.batch.field_obs]) if v >= 0)
            wx.fields = {s_: f for s_, f in wx.fields.items() if s_ in used}
            xsteps = min(4, max(2, NF))  # as many steps as batches in flight (at most 4 contexts: each holds a row buffer of up to 24 GiB)
            mx_ = measure(wx, xsteps, 1, e2e=False, events_pass=False)
            (t_mx, x_mx), (c_tot,) = over_ranks([mx_["dev_ms"], mx_["xch_ms"]], [float(mx_["conv"])])
            pk, _ = measured_peaks()
            algx = algorithmic_bytes(wx, mx_["prof"])
            extra[xcfg] = {"workload": wx.description, "problems_total": total, "problems_per_gpu": int(wx.batch.B), "scaling": "strong", "batches_in_flight": mx_["in_flight"],
                           "converged_total": int(c_tot), "status_counts_rank0": status_dict(mx_["status"]),
                           "steps": xsteps, "ms_per_step": t_mx / xsteps, "xch_ms_per_step": x_mx / xsteps, "value": c_tot * xsteps / (t_mx * 1e-3), "unit": UNIT,
                           "linearize_ms_per_step_rank0": mx_["lin_ms"] / xsteps, "step_ms_per_step_rank0": mx_["step_ms"] / xsteps,
                           "roofline_frac_rank0": (algx / (mx_["lin_ms"] * 1e-3) / 1e9 / pk) if mx_["lin_ms"] > 0 else None,
                           "iterations_max": int(mx_["iters"].max()), "iterations_mean": float(mx_["iters"].mean())}

    if rank == 0:
        b = w.batch
        B = b.B
        peak, peak_src = measured_peaks()
        alg = algorithmic_bytes(w, m["prof"])
        ev = m["ev"]
        kernel_ms_events = ev["lin_ms"] if ev else None
        kernel_ms = kernel_ms_events if kernel_ms_events else m["lin_ms"]
        achieved = alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        value = conv_tot * args.steps / (step_dev_ms_mx * 1e-3)
        e2e_value = conv2_tot * args.steps / (e2e_ms_mx * 1e-3)
        traffic, traffic_src = ncu_traffic(cfg, args.no_jrows, args.scale)
        nlaunch = m["prof"]["linearize_launches"] / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_dev_ms_mx / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 point kernel (f64 voxel coordinates, f64 sums of J^T r and cost, 3xTF32 tensor-core J^T J) + f64 trajectory / step solve",
            "data": "synthetic",
            "config": config_block(w, args, opts, world, converged=int(conv_tot), iterations_histogram=np.bincount(m["iters"], minlength=1).tolist(),
                                   status_counts_rank0=status_dict(m["status"]), wall_ms_per_step=wall_ms_mx / args.steps,
                                   batches_in_flight=m["in_flight"], host_blocking_sync=bool(blocking),
                                   shards=("a differently seeded batch per rank" if distinct else "the same batch (the N = 1 workload) on every rank: equal work per GPU") if world > 1 else "one batch",
                                   solve_ms_per_step=solve_ms_mx / args.steps, solve_ms_note="mean latency of one batch solve (library events on its stream) with "
                                   f"{m['in_flight']} batches in flight; ms_per_step is the device time of the region / steps",
                                   xch_ms_per_step=xch_ms_mx / args.steps,
                                   exchange="one NCCL all-gather of [B][nopt*T+2] f32 per batch on a side stream, overlapped with the next solves (the library's result buffer is staged by one device-to-device copy)",
                                   **({"secondary_mode": f"slow_window={args.slow_window}, slow_ftol={args.slow_ftol}"} if args.slow_window > 0 else {}),
                                   **({"secondary_mode_fused": "k_solve_fused"} if args.fused else {}),
                                   **({"extra": extra} if extra else {})),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": int(m["d2h"]),
                    "timing": "wall clock between device synchronisations around the K gto_solve_batch calls with pinned host buffers, "
                              f"{m['in_flight']} calls in flight (one host thread + context each)"},
            "gpu_launches": int(m["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_linearize_cull (+ k_item_fk, its per-item pre-pass)" if not args.fused else "k_solve_fused (FK + linearise phases)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": alg / args.steps,
                         "algorithmic_bytes_per_launch": alg / max(1.0, m["prof"]["linearize_launches_with_work"]),
                         "kernel_ms_per_step": kernel_ms / args.steps, "avg_launch_us": 1e3 * kernel_ms / max(1.0, m["prof"]["linearize_launches_with_work"]),
                         "launches_per_step": nlaunch,
                         "timing": "CUDA events around every linearise launch (k_item_fk + k_linearize_cull) on the library's stream, separate pass of the same "
                                   "steps" if ev else "in-kernel %globaltimer stamps",
                         "kernel_share_of_step": (ev["lin_ms"] / ev["solve_ms"]) if ev else (m["lin_ms"] / m["dev_ms"] if m["dev_ms"] else None),
                         "whole_step_frac": alg / (m["dev_ms"] * 1e-3) / 1e9 / peak if m["dev_ms"] else None,
                         "kernel_timing_note": "linearise / step kernel times are taken with ONE batch in flight (events pass on context 0); the stamps of the "
                                               "timed region overlap between the in-flight batches",
                         "globaltimer_stamps": {"achieved": alg / (m["lin_ms"] * 1e-3) / 1e9 if m["lin_ms"] > 0 else None, "kernel_ms_per_step": m["lin_ms"] / args.steps,
                                                "step_kernel_ms_per_step": m["step_ms"] / args.steps,
                                                "note": "first CTA start .. last warp end per launch, written by the kernels inside the timed region (no event serialisation)"},
                         "step_kernel_ms_per_step": (ev["step_ms"] if ev else m["step_ms"]) / args.steps,
                         "links_culled_frac": 1.0 - m["last_prof"]["links_active"] / max(1, m["last_prof"]["links_tested"])},
            "clocks": clocks,
        }
        # CPU baseline on rank 0 at N=1 only: the full batch of the same workload (C2: ~0.6 s on 16 threads), else a bounded sample
        if world == 1 and args.cpu_sample != 0:
            ncpu = os.cpu_count() or 1
            ns = args.cpu_sample if args.cpu_sample > 0 else (B if cfg in ("C1", "C2") else min(B, 16 * ncpu))
            idxs = np.arange(B) if ns >= B else np.unique(np.linspace(0, B - 1, ns).astype(int))
            dt, cres = time_cpu(w, idxs, ncpu)
            line["cpu_baseline"] = {"value": int(np.sum(cres["status"] == 0)) / dt, "unit": UNIT, "cores": int(cres["threads"]), "kind": "port",
                                    "sample": f"{len(idxs)} of {B} problems of the same workload, oracle/gto_oracle.c (projected LM, float64), "
                                              f"{int(cres['threads'])} pthreads on {ncpu} host cores, {dt:.1f} s"}
        _emit(line)
    for c in ctxs:
        c.close()
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
