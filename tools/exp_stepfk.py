import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W
from helpers import upload_fields
w = W.make_workload("C2")
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
upload_fields(ctx, w)
os.environ.pop("GTO_STEP_FK", None)
ref = ctx.solve_batch(w.batch); pr = ctx.profile()
for mode in ("tail", "all"):
    os.environ["GTO_STEP_FK"] = mode
    for rep in range(2):
        res = ctx.solve_batch(w.batch); pf = ctx.profile()
        same = all(np.array_equal(res[k], ref[k]) for k in ("Q", "cost", "iters", "status"))
        print(mode, rep, "identical", same, "launches", pf["linearize_launches"], pf["kernel_launches"], "solve_ms", round(pf["solve_ms"], 3), "iters max", res["iters"].max(),
              "status", np.bincount(res["status"], minlength=5), "max|dQ|", np.abs(res["Q"] - ref["Q"]).max(), "ref launches", pr["linearize_launches"], round(pr["solve_ms"], 3))
