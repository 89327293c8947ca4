"""GPU vs C-oracle check on scaled-down BASELINE configs (prints, never raises on mismatch)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import c_oracle as CO
from grasptrajopt_b200 import capi, workloads as W
from helpers import upload_fields

ctx = capi.GtoContext(0)
for cfg, B, nf in (("C2", 16, None), ("C3", 8, 128), ("C4", 8, None), ("C5", 16, None)):
    full = {"C2": 256, "C3": 1024, "C4": 4096, "C5": 16384}[cfg]
    w = W.make_workload(cfg, scale=B / full, n_field=nf)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    t0 = time.time(); res = ctx.solve_batch(w.batch); t1 = time.time()
    ora = CO.solve_workload(w)
    dq = np.abs(res["Q"] - ora["Q"]).max(axis=(1, 2))
    both = (res["status"] == 0) & (ora["status"] == 0)
    pf = ctx.profile()
    print(f"{cfg} B={w.batch.B} T={w.batch.T} nopt={w.table.nopt}: gpu status {np.bincount(res['status'], minlength=5)} oracle status {np.bincount(ora['status'], minlength=5)} "
          f"iters gpu {res['iters'][:8]} ora {ora['iters'][:8]} max|dQ| both-converged {dq[both].max() if both.any() else -1:.2e} (n={both.sum()}) all {np.round(dq,6)[:8]} "
          f"links active {pf['links_active']}/{pf['links_tested']} solve {pf['solve_ms']:.1f} ms lin {pf['linearize_ms']:.1f} step {pf['step_ms']:.1f} wall {t1-t0:.2f}s", flush=True)
ctx.close()
