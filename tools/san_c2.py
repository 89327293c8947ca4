"""compute-sanitizer target: one small multi-item solve of the C2 workload (more items than CTAs)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W
from helpers import upload_fields

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
full = {"C2": 256, "C3": 1024, "C4": 4096, "C5": 256}[cfg]
w = W.make_workload(cfg, scale=B / full, n_field=int(sys.argv[3]) if len(sys.argv) > 3 else None)
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
upload_fields(ctx, w)
o = capi.Options.defaults() if hasattr(capi.Options, "defaults") else None
res = ctx.solve_batch(w.batch)
print("status", np.bincount(res["status"], minlength=5), "iters max", res["iters"].max(), ctx.profile())
