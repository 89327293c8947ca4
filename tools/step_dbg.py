"""Phase clocks of k_step_cr (CTA 0 of the last launch with work) on a BASELINE configuration: gto_configure("step_dbg", 1)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = W.make_workload(cfg, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
for slot, cf in w.fields.items():
    ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
ctx.upload_batch(w.batch)
for bundle in (3, 1, 0):
    o = capi.default_options()
    o.bundle = bundle
    ctx.configure(step_dbg=0)
    ctx.solve_resident(o)
    ctx.configure(step_dbg=int(os.environ.get("STEP_DBG_ITER", "30")), step_dbg_cta=int(os.environ.get("STEP_DBG_CTA", "0")))
    ctx.solve_resident(o)
    p = ctx.profile()
    res = ctx.download_batch()
    print(f"bundle {bundle}: {p['solve_ms']:.3f} ms, {p['iterations']} iterations, lin {p['linearize_ms']:.3f} ms, step {p['step_ms']:.3f} ms, status {np.bincount(res['status'], minlength=5)}", flush=True)
ctx.close()
