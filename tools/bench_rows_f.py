"""Measurements for the SURVEY.md section 8(f) rows built so far (not part of bench.py's contract line):
  row 1  batched IK seeding   -- Panda, B goals in one batch (T = 3 layout, goal rows only), vs the C oracle on the host cores
  row 2  cost-field build     -- DepthPointCloud.get_sdf_cost on a 640x480 depth image, 128^3 grid, GPU kernel vs scikit-learn KD-tree
  row 3  seed ranking         -- gto_plan_cost over 256 candidate plans x 30 knots x 2000 points vs the NumPy oracle
  row 4  base placement       -- gto_base_place, 4096 problems x 10 goals in one launch vs oracle/base_oracle.py on one host core
Prints one JSON object."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests"), os.path.join(REPO, "grasptrajopt_b200", "compat"), os.path.join(REPO, "tools")):
    sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi, workloads as W, scenes as S

out = {}
# ---- row 1: batched IK ----
w = W.make_workload("C2", scale=16.0)  # 4096 grasps around the target
t, b = w.table, w.batch
B = b.B
ik = capi.Batch(T=3, dt=1.0, qc=b.qc, q_seed=np.repeat(b.qc[:, None, :], 3, axis=1), goal_tf=b.goal_tf, base_position=np.zeros((B, 3)),
                field_all=None, field_obs=None, standoff_offset=-1, use_standoff=False, collision_avoidance=False, w_vel=0.0)
ctx = capi.GtoContext(0)
ctx.set_robot(t)
opts = capi.default_options(max_iter=50)
opts.bundle = 0  # as gto.IKSolver without collision avoidance: goal rows only, a smooth problem (no field kinks)
for _ in range(3):
    res = ctx.solve_batch(ik, opts)
t0 = time.perf_counter()
for _ in range(5):
    res = ctx.solve_batch(ik, opts)
dt = (time.perf_counter() - t0) / 5
pf = ctx.profile()
out["ik_batch"] = {"goals": int(B), "converged": int((res["status"] == 0).sum()), "ms_per_batch_e2e": 1e3 * dt, "ik_solves_per_s": B / dt,
                   "device_ms": pf["solve_ms"], "iterations_max": int(res["iters"].max())}
try:
    import c_oracle as CO
    wk = W.Workload(w.name, t, {}, ik, w.RT, 0.0, "z", w.q_star)
    idx = np.arange(0, B, 8)
    t0 = time.perf_counter(); ro = CO.solve_workload(wk, indices=idx, options=CO.default_options(max_iter=50, bundle=0)); dtc = time.perf_counter() - t0
    dq = np.abs(ro["Q"][:, 2] - res["Q"][idx, 2]).max(axis=1)
    both = (ro["status"] == 0) & (res["status"][idx] == 0)
    out["ik_batch"]["cpu_port"] = {"ik_solves_per_s": len(idx) / dtc, "threads": int(ro["threads"]), "sample": int(len(idx)),
                                   "median_abs_dq_both_converged": float(np.median(dq[both])) if both.any() else None,
                                   "frac_within_1e-4_rad": float((dq[both] < 1e-4).mean()) if both.any() else None, "n_both": int(both.sum()),
                                   "note": "a 7-DoF arm has a one-parameter family of IK solutions per pose, so the joint vectors of two float paths need not agree; "
                                           "both reach the pose (see converged counts)"}
except Exception as e:  # pragma: no cover
    out["ik_batch"]["cpu_port"] = {"error": repr(e)}

# ---- row 2: cost field from a depth image ----
from mesh_to_sdf.depth_point_cloud import DepthPointCloud
from dpc_oracle import KDTreeDepthPointCloud
H, Wd, f = 480, 640, 550.0
K = np.array([[f, 0, Wd / 2], [0, f, H / 2], [0, 0, 1.0]])
cam = np.eye(4); cam[:3, :3] = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]]); cam[:3, 3] = [0.5, 0.0, 1.2]
v, u = np.mgrid[0:H, 0:Wd]
depth = np.full((H, Wd), 1.2, np.float32)
depth[180:300, 250:390] = 1.05
depth[60:140, 80:560] = (1.15 - 0.0002 * (u[60:140, 80:560] - 80)).astype(np.float32)
for n in (64, 128):
    g = np.stack(np.meshgrid(np.linspace(-0.4, 1.4, n), np.linspace(-1.4, 1.4, n), np.linspace(-0.4, 1.4, n), indexing="ij"), axis=-1).reshape(-1, 3)
    gpu = DepthPointCloud(depth, K, cam, threshold=1.5)
    gpu.get_sdf_cost(g[:4096])
    t0 = time.perf_counter(); c_gpu = gpu.get_sdf_cost(g); t1 = time.perf_counter()
    rec = {"cloud_points": int(gpu.points.shape[0]), "queries": int(g.shape[0]), "gpu_kernel_ms": gpu.last_kernel_ms, "gpu_call_ms": 1e3 * (t1 - t0),
           "pairs_per_s": gpu.points.shape[0] * g.shape[0] / (gpu.last_kernel_ms * 1e-3)}
    if n == 64:
        cpu = KDTreeDepthPointCloud(depth, K, cam, threshold=1.5)
        t0 = time.perf_counter(); c_cpu = cpu.get_sdf_cost(g); t1 = time.perf_counter()
        rec["kdtree_cpu_ms"] = 1e3 * (t1 - t0)
        rec["max_abs_cost_diff"] = float(np.abs(c_gpu - c_cpu).max())
        rec["mismatch_frac_gt_1e-5"] = float((np.abs(c_gpu - c_cpu) > 1e-5).mean())
    out[f"cost_field_{n}"] = rec

# ---- row 3: seed ranking (value-only pass over whole plans) ----
w2 = W.make_workload("C2")
t2, b2 = w2.table, w2.batch
ctx2 = capi.GtoContext(0)
ctx2.set_robot(t2)
cf = w2.fields[int(b2.field_obs[0])]
ctx2.set_field(0, cf.cost, cf.origin, cf.pitch)
plans = b2.q_seed  # 256 candidate seeds x 30 knots
ctx2.plan_cost(plans, 0)
t0 = time.perf_counter()
for _ in range(5):
    c_gpu, _ = ctx2.plan_cost(plans, 0)
dtg = (time.perf_counter() - t0) / 5
import gto_oracle as O
fld = O.Field(cf.cost, cf.origin, cf.pitch)
t0 = time.perf_counter()
c_cpu = np.array([O.plan_cost_nearest(t2, plans[i], fld, np.zeros(3))[0] for i in range(16)])
dtc = (time.perf_counter() - t0) / 16
out["seed_ranking"] = {"plans": int(plans.shape[0]), "knots": int(plans.shape[1]), "points": int(t2.npoints), "gpu_call_ms": 1e3 * dtg,
                       "plans_per_s_gpu": plans.shape[0] / dtg, "plans_per_s_numpy_1core": 1.0 / dtc, "cpu_sample": 16,
                       "max_rel_diff_vs_oracle": float(np.abs(c_gpu[:16] - c_cpu).max() / max(1e-12, np.abs(c_cpu).max())),
                       "note": "float32 FK on the GPU vs float64 in the oracle: a point within rounding of a cell face may read the neighbouring node"}
ctx2.close()

# ---- row 4: mobile-base placement ----
import bench_base
out.update(bench_base.run())
print(json.dumps(out))
