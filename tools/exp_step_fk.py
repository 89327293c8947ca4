"""Experiment: FK item records written by the step kernel (gto_configure("step_fk", n): when at most n problems are active) instead
of a k_item_fk launch -- one launch less per tail iteration."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W

w = W.make_workload("C2")
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
for slot, cf in w.fields.items():
    ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
ctx.upload_batch(w.batch)
for n in (0, 8, 16, 32, 64, 1 << 30):
    ctx.configure(step_fk=n)
    best = None
    for _ in range(4):
        ctx.solve_resident()
        p = ctx.profile()
        if best is None or p["solve_ms"] < best["solve_ms"]:
            best = p
    res = ctx.download_batch()
    print(f"step_fk {n}: solve {best['solve_ms']:.3f} ms, lin {best['linearize_ms']:.3f}, step {best['step_ms']:.3f}, launches {best['kernel_launches']}, conv {int(np.sum(res['status'] == 0))}", flush=True)
ctx.close()
