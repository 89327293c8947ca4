"""Shared test helpers: turn a Workload / Batch into oracle Problems (identical bits on both sides)."""
import numpy as np

import gto_oracle as O


def oracle_field(cf):
    return O.Field(cf.cost, cf.origin, cf.pitch)


def problems_from_workload(w, indices=None, sdf_mode="trilinear"):
    b = w.batch
    fields = {slot: oracle_field(cf) for slot, cf in w.fields.items()}
    out = []
    for i in (range(b.B) if indices is None else indices):
        fa = fields.get(int(b.field_all[i])) if b.field_all is not None and b.field_all[i] >= 0 else None
        fo = fields.get(int(b.field_obs[i])) if b.field_obs is not None and b.field_obs[i] >= 0 else None
        out.append(
            O.Problem(
                table=w.table, T=b.T, dt=b.dt, qc=b.qc[i], RT=w.RT[i], q_seed=b.q_seed[i],
                base_position=np.zeros(3) if b.base_position is None else b.base_position[i],
                field_all=fa, field_obs=fo, standoff_offset=b.standoff_offset, standoff_distance=w.standoff_distance,
                axis_standoff=w.axis_standoff, use_standoff=b.use_standoff, collision_avoidance=b.collision_avoidance,
                w_goal=b.w_goal, w_obs=b.w_obs, w_vel=b.w_vel, sdf_mode=sdf_mode, obs_linear=bool(int(b.flags) & 2),
            )
        )
    return out


def upload_fields(ctx, w):
    for slot, cf in w.fields.items():
        ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)


def small_workload(config="C2", table_name=None, B=4, n_field=64, seed=7):
    """A BASELINE config shrunk for oracle-speed parity tests: same generator, fewer problems, coarser field and
    (optionally) the small point-set table."""
    from grasptrajopt_b200 import workloads as W
    from grasptrajopt_b200.robot_table import RobotTable
    import os

    full = {"C1": 1, "C2": 256, "C3": 1024, "C4": 4096, "C5": 256}[config]
    w = W.make_workload(config, scale=B / full, seed=seed, n_field=n_field)
    if table_name is not None:
        t = RobotTable.load(os.path.join(W.ASSETS, table_name + ".npz"))
        assert t.nopt == w.table.nopt and t.ndof == w.table.ndof
        w.table = t
    return w


def base_placement_case(table_name="panda_small", B=4, n=4, seed=3, spread=1.5, with_grid=True):
    """Seeded base-placement problems: goals are gripper poses of random arm configurations seen from a
    base displaced by a random planar motion, so every problem has a zero-residual placement the effort term pulls on."""
    import os
    import base_oracle as BO
    from grasptrajopt_b200 import workloads as W
    from grasptrajopt_b200.robot_table import RobotTable

    t = RobotTable.load(os.path.join(W.ASSETS, table_name + ".npz"))
    qc = {9: W.PANDA_QC, 15: W.FETCH_QC, 18: np.concatenate([[0.0, 0.0, 0.0], W.FETCH_QC])}[t.ndof].copy()
    rng = np.random.default_rng(seed)
    Ginv = np.linalg.inv(O.hom(t.G))
    RTs = np.zeros((B, n, 4, 4))
    for b in range(B):
        ystar = np.array([rng.uniform(-spread, spread), rng.uniform(-spread, spread), rng.uniform(-1.0, 1.0)])
        Tbi = np.linalg.inv(BO.base_tf(ystar))
        for i in range(n):
            q = qc.copy()
            q[t.opt_qidx] = np.clip(qc[t.opt_qidx] + rng.normal(0, 0.4, t.nopt), t.lo, t.hi)
            RTs[b, i] = Tbi @ O.gripper_frame(t, q) @ Ginv
    grid = origin = None
    res = 0.05
    if with_grid:
        grid = (rng.random((60, 70)) < 0.05).astype(np.float32)
        origin = np.array([-1.0, -1.7])
    return t, qc, RTs, grid, origin, res
