"""GPU vs C-oracle solve parity on the BASELINE configurations: C2 at full size (256 problems), 64-problem shards of C3
(Fetch-8 shelf, 256^3), C4 (Fetch-10 mobile) and C5 (Panda clutter).  `compare()` is what tests/test_gpu_parity_configs.py
asserts on; run as a script it prints one line per config, writes gpurun_out/r2_parity_configs.json and exits non-zero
on a mismatch.  (reference: gto/gto_planner.py:42-142 with data/configs/fetch.yaml:15-39)"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np

FULL = {"C2": 256, "C3": 1024, "C4": 4096, "C5": 16384}
CASES = (("C2", 256), ("C3", 64), ("C4", 64), ("C5", 64))
TOL_Q = 1e-4  # rad, north-star tolerance
# fraction of the both-converged problems that may end in a neighbouring kink minimiser (tests/test_gpu_parity_configs.py)
OUT_FRAC = {"C2": 0.02, "C3": 0.20, "C4": 0.20, "C5": 0.02}


def shard_workload(cfg, B):
    """The first B problems of the full-size configuration (same fields, same point sets, same seeds)."""
    from grasptrajopt_b200 import workloads as W

    if cfg == "C5":  # scenes come in blocks of 256 problems: one scene, then the first B of it
        w = W.make_workload(cfg, scale=1.0 / 64.0)
    else:
        w = W.make_workload(cfg, scale=B / FULL[cfg])
    if w.batch.B > B:
        w.batch = W.slice_batch(w.batch, 0, B)
        w.RT = w.RT[:B]
        w.q_star = w.q_star[:B]
    return w


def compare(ctx, cfg, B, options=None):
    import c_oracle as CO
    from grasptrajopt_b200 import capi
    from helpers import upload_fields

    w = shard_workload(cfg, B)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    t0 = time.time()
    res = ctx.solve_batch(w.batch, options)
    t1 = time.time()
    ora = CO.solve_workload(w)
    t2 = time.time()
    dq = np.abs(res["Q"] - ora["Q"]).max(axis=(1, 2))
    both = (res["status"] == capi.STATUS_CONVERGED) & (ora["status"] == capi.STATUS_CONVERGED)
    rel_cost = np.abs(res["cost"] - ora["cost"]) / np.maximum(ora["cost"], 1e-12)
    pf = ctx.profile()
    out = dict(
        config=cfg, B=int(w.batch.B), T=int(w.batch.T), nopt=int(w.table.nopt), npoints=int(w.table.npoints),
        gpu_status=np.bincount(res["status"], minlength=5).tolist(), oracle_status=np.bincount(ora["status"], minlength=5).tolist(),
        both_converged=int(both.sum()), status_equal=int((res["status"] == ora["status"]).sum()),
        iters_equal=int((res["iters"] == ora["iters"]).sum()), iters_equal_both=int((res["iters"] == ora["iters"])[both].sum()),
        max_iter_diff_both=int(np.abs(res["iters"] - ora["iters"])[both].max()) if both.any() else 0,
        max_dq_both=float(dq[both].max()) if both.any() else 0.0, n_dq_both_over_tol=int((dq[both] > TOL_Q).sum()),
        dq_both_sorted_top=np.sort(dq[both])[::-1][:8].tolist(), median_dq_both=float(np.median(dq[both])) if both.any() else 0.0,
        max_rel_cost_both=float(rel_cost[both].max()) if both.any() else 0.0, max_rel_cost_all=float(rel_cost.max()),
        max_dq_all=float(dq.max()), gpu_iters_mean=float(res["iters"].mean()), oracle_iters_mean=float(ora["iters"].mean()),
        gpu_solve_ms=float(pf["solve_ms"]), gpu_wall_s=t1 - t0, oracle_wall_s=t2 - t1, oracle_threads=int(ora["threads"]),
    )
    out["_arrays"] = dict(res=res, ora=ora, dq=dq, both=both, workload=w)
    return out


def main():
    from grasptrajopt_b200 import capi

    ctx = capi.GtoContext(0)
    report, bad = [], 0
    for cfg, B in CASES:
        r = compare(ctx, cfg, B)
        a = r.pop("_arrays")
        report.append(r)
        print(json.dumps(r), flush=True)
        if r["n_dq_both_over_tol"] > 0:
            bad += r["n_dq_both_over_tol"] > max(1, int(OUT_FRAC[cfg] * r["both_converged"]))
            idx = np.nonzero(a["both"] & (a["dq"] > TOL_Q))[0]
            print(f"  {cfg}: problems over tolerance {idx.tolist()} dq {a['dq'][idx]} gpu iters {a['res']['iters'][idx]} oracle iters {a['ora']['iters'][idx]}")
    ctx.close()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "r2_parity_configs.json"), "w") as f:
        json.dump(report, f, indent=1)
    if bad:
        raise SystemExit(f"parity mismatch on {bad} configuration(s)")


if __name__ == "__main__":
    main()
