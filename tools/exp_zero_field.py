"""Experiment: C2 batch with the cost fields switched off (every link culled) -- isolates the producer path of k_linearize_cull."""
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
mode = sys.argv[1] if len(sys.argv) > 1 else "zero"
if mode == "zero":
    w.batch.field_all[:] = -1
    w.batch.field_obs[:] = -1
if len(sys.argv) > 2:
    w.batch.flags |= capi.FLAG_NO_JROWS
res = ctx.solve_batch(w.batch, capi.default_options(max_iter=6))
print(ctx.profile())
