import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = W.make_workload(cfg, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
ctx = capi.GtoContext(0)
ctx.set_robot(w.table)
for s, cf in w.fields.items():
    ctx.set_field(s, cf.cost, cf.origin, cf.pitch)
kw = {k[4:].lower(): (int(v) if k[4:].lower() in ('max_iter', 'check_every', 'slow_window') else float(v)) for k, v in os.environ.items() if k.startswith('OPT_')}
opts = capi.default_options(**kw)
print('options', kw)
for rep in range(3):
    t0 = time.time(); res = ctx.solve_batch(w.batch, opts); dt = time.time() - t0
    pr = ctx.profile()
    print(cfg, f"rep {rep} wall {dt*1e3:.1f} ms solve {pr['solve_ms']:.2f} lin {pr['linearize_ms']:.2f} step {pr['step_ms']:.2f} launches {pr['linearize_launches']} conv {np.mean(res['status']==0):.3f} iters mean {res['iters'].mean():.1f} p50 {np.median(res['iters'])} p90 {np.percentile(res['iters'],90)} max {res['iters'].max()} status {np.bincount(res['status'],minlength=4)}", flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(REPO, "gpurun_out", f"{cfg.lower()}_solution.npz"), Q=res["Q"], cost=res["cost"], iters=res["iters"], status=res["status"])
ctx.close()
