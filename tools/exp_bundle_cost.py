"""Experiment: step-kernel time of a C2 solve against the bundle options (pieces kept, radius) -- what the bookkeeping costs
(radius 1e-12: pieces are kept and re-based but never used) and what the extra right-hand sides / dual QP cost."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W

w = W.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C2")
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
for slot, cf in w.fields.items():
    ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
ctx.upload_batch(w.batch)
for bundle, radius in ((0, 3e-3), (3, 1e-12), (1, 3e-3), (2, 3e-3), (3, 3e-3), (4, 3e-3)):
    o = capi.default_options()
    o.bundle, o.bundle_radius = bundle, radius
    best = None
    for _ in range(3):
        ctx.solve_resident(o)
        p = ctx.profile()
        if best is None or p["solve_ms"] < best["solve_ms"]:
            best = p
    res = ctx.download_batch()
    print(f"bundle {bundle} radius {radius:g}: solve {best['solve_ms']:.3f} ms, lin {best['linearize_ms']:.3f}, step {best['step_ms']:.3f} ms over {best['step_launches']} launches, "
          f"conv {int(np.sum(res['status'] == 0))}, mean iterations {res['iters'].mean():.1f}", flush=True)
ctx.close()
