"""Experiment: linearise / step time of a C2 solve under the run-time knobs of k_linearize_cull (ring slots, consumer warps)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = W.make_workload(cfg, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
SETS = {"a": ({}, {"cull_nslot": 2}, {"cull_nslot": 3}, {"cull_nslot": 6}, {"cull_nslot": 8}, {"cons_warps": 4}, {"cons_warps": 6}, {"cons_warps": 4, "cull_nslot": 6},
              {"slot_floats": 2048}, {"slot_floats": 8192}),
        "b": ({}, {"cons_warps": 4, "cull_nslot": 2, "slot_floats": 2048}, {"cons_warps": 4, "cull_nslot": 3, "slot_floats": 2048}, {"cons_warps": 5, "cull_nslot": 2, "slot_floats": 2048},
              {"cons_warps": 6, "cull_nslot": 2, "slot_floats": 2048}, {"cons_warps": 4, "cull_nslot": 2, "slot_floats": 3072}, {"cons_warps": 3, "cull_nslot": 2, "slot_floats": 2048})}
for knobs in SETS[sys.argv[3] if len(sys.argv) > 3 else "a"]:
    ctx = capi.GtoContext(0)
    try:
        ctx.configure(**knobs)
        ctx.set_robot(w.table)
        for slot, cf in w.fields.items():
            ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
        ctx.upload_batch(w.batch)
        best = None
        for _ in range(4):
            ctx.solve_resident()
            p = ctx.profile()
            if best is None or p["solve_ms"] < best["solve_ms"]:
                best = p
        res = ctx.download_batch()
        print(f"{cfg} {knobs}: solve {best['solve_ms']:.3f} ms, lin {best['linearize_ms']:.3f} ms, step {best['step_ms']:.3f} ms, conv {int(np.sum(res['status'] == 0))}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{cfg} {knobs}: {e}", flush=True)
    ctx.close()
