"""One warm-up solve + one solve of a BASELINE configuration with inputs resident -- the command wrapped by ncu for profiles/:
  ncu --metrics gpu__time_duration.sum --clock-control none -s <launches of the warm-up solve> -c 320 --csv --log-file gpurun_out/launches.csv python tools/prof_c2.py
  ncu --set full --clock-control none --import-source on -k regex:k_linearize_cull -s 104 -c 1 -o gpurun_out/prof_full python tools/prof_c2.py
Prints the per-solve profile (launch counts) so that -s can be chosen."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
nsolve = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = W.make_workload(cfg)
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
for slot, cf in w.fields.items():
    ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
ctx.upload_batch(w.batch)
for i in range(nsolve):
    ctx.solve_resident()
    p = ctx.profile()
    print(f"solve {i}: {p['solve_ms']:.3f} ms, {p['kernel_launches']} launches, {p['iterations']} iterations, lin {p['linearize_ms']:.3f} ms, step {p['step_ms']:.3f} ms")
res = ctx.download_batch()
print("status", np.bincount(res["status"], minlength=5), "iters hist", np.bincount(res["iters"]).tolist())
ctx.close()
