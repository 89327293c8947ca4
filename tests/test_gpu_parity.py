"""Parity of the CUDA path (through the C-ABI) against the float64 oracle.  Needs a B200: run with -m gpu."""
import os

import numpy as np
import pytest

import gto_oracle as O
from grasptrajopt_b200 import capi, workloads as W
from helpers import problems_from_workload, small_workload, upload_fields

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.GtoContext(0)
    yield c
    c.close()


def _eval_oracle(w):
    rows, H, g, cost = [], [], [], []
    for p in problems_from_workload(w):
        lin = O.linearize(p, p.q_seed)
        rows.append(O.pack_rows(p, lin))
        H.append(lin.H); g.append(lin.g); cost.append(lin.cost_pts)
    return np.stack(rows), np.stack(H), np.stack(g), np.stack(cost)


def _check_eval(ctx, w, frac_bad_rows=2e-4):
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    out = ctx.eval_batch(w.batch)
    rows, H, g, cost = _eval_oracle(w)
    n = w.table.nopt
    # residual column: continuous in the inputs -> tight everywhere
    np.testing.assert_allclose(out["rows"][..., n], rows[..., n], rtol=0, atol=2e-5)
    # Jacobian columns: the trilinear gradient is discontinuous across cell faces, so a float32 point that lands
    # on the other side of a face than its float64 twin legitimately differs; everything else must agree
    err = np.abs(out["rows"][..., :n] - rows[..., :n]).max(axis=-1)
    scale = 1.0 + np.abs(rows[..., :n]).max(axis=-1)
    bad = err > 1e-4 * scale
    assert bad.mean() <= frac_bad_rows, f"{bad.sum()} of {bad.size} Jacobian rows differ"
    if not bad.any():
        # tensor-core J^T J (error-compensated TF32, fp32 accumulate); J^T r and cost: float32 products, float64 sums
        Hs = np.abs(H).max(axis=(-1, -2), keepdims=True) + 1e-6
        assert np.max(np.abs(out["H"] - H) / Hs) < 2e-5
        gs = np.abs(g).max(axis=-1, keepdims=True) + 1e-4
        assert np.max(np.abs(out["g"] - g) / gs) < 2e-5
        np.testing.assert_allclose(out["cost"], cost, rtol=5e-6, atol=1e-8)
    assert np.abs(rows[..., :n]).max() > 0.1 and cost.max() > 0
    return out


KERNELS = [pytest.param(0, id="cull"), pytest.param(capi.FLAG_NO_CULL, id="nocull")]


@pytest.mark.parametrize("kflag", KERNELS)
def test_eval_parity_panda_tabletop(ctx, kflag):
    """Rows A1-A9: FK -> points -> trilinear SDF -> residual + Jacobian rows -> J^T J / J^T r, Panda (row stride 8)."""
    w = small_workload("C2", "panda_small", B=4, n_field=64)
    w.batch.flags = kflag
    out = _check_eval(ctx, w)
    assert np.abs(out["rows"][:, : 30 * w.table.npoints, :7]).max() > 0  # the obstacle term is active in this scene


@pytest.mark.parametrize("kflag", KERNELS)
def test_eval_parity_full_point_set(ctx, kflag):
    w = small_workload("C2", None, B=2, n_field=96)
    w.batch.flags = kflag
    _check_eval(ctx, w)


@pytest.mark.parametrize("kflag", KERNELS)
def test_eval_parity_fetch8_shelf(ctx, kflag):
    """nopt = 8 -> row stride 9 (unaligned store path), prismatic torso joint in the chain."""
    w = small_workload("C3", None, B=2, n_field=96)
    w.batch.flags = kflag
    _check_eval(ctx, w)


@pytest.mark.parametrize("kflag", KERNELS)
def test_eval_parity_fetch10_mobile(ctx, kflag):
    """nopt = 10 -> 16-wide tensor-core tile (two n-tiles), virtual planar base joints."""
    w = small_workload("C4", None, B=2, n_field=64)
    w.batch.flags = kflag
    _check_eval(ctx, w)


def test_eval_parity_points_outside_the_field(ctx):
    """Clamped lookups (gto_models.py:176-183): a base offset pushes part of the robot outside the voxel field, so the
    pipelined kernel must take its generic path (index clamping, partial bricks) and still match the oracle."""
    w = small_workload("C2", "panda_small", B=2, n_field=48)
    w.batch.base_position = np.array([[0.0, 1.9, 0.0], [-0.7, 0.0, -0.6]])
    _check_eval(ctx, w, frac_bad_rows=1e-3)


def test_culling_is_exact(ctx):
    """Links whose node box holds only zero cost nodes are culled (zero rows written by bulk copies): the result must be
    identical to the same kernel with the test disabled (rows bit for bit, the per-knot sums up to summation order), and
    links must actually be culled in this scene."""
    for cfg, tab, nf in (("C2", "panda_small", 64), ("C3", None, 96), ("C4", None, 64)):
        w = small_workload(cfg, tab, B=3, n_field=nf)
        ctx.set_robot(w.table)
        upload_fields(ctx, w)
        outs = []
        for flags in (0, capi.FLAG_NO_CULL):
            w.batch.flags = flags
            outs.append(ctx.eval_batch(w.batch))
        w.batch.flags = 0
        for o in outs[1:]:
            np.testing.assert_array_equal(o["rows"], outs[0]["rows"], err_msg=cfg)
            for k in ("H", "g", "cost"):  # same terms, summed in a different order over the warps
                np.testing.assert_allclose(o[k], outs[0][k], rtol=2e-5, atol=1e-7, err_msg=f"{cfg} {k}")
        assert np.abs(outs[0]["rows"]).max() > 0
    # the solver reports how many (problem, knot, link) triples survived the test
    w = small_workload("C2", "panda_small", B=4, n_field=64)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    ctx.solve_batch(w.batch)
    pf = ctx.profile()
    assert 0 < pf["links_active"] < 0.6 * pf["links_tested"]


def test_eval_options_no_collision_no_standoff(ctx):
    w = small_workload("C2", "panda_small", B=2, n_field=64)
    w.batch.collision_avoidance = False
    w.batch.use_standoff = False
    ctx.set_robot(w.table)
    out = ctx.eval_batch(w.batch)
    rows, H, g, cost = _eval_oracle(w)
    assert out["rows"].shape == rows.shape == (2, 3 * w.table.grip_pt_count, 8)
    np.testing.assert_allclose(out["rows"], rows, rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["g"], g, rtol=2e-4, atol=1e-5)


def _solve_both(ctx, w, opts=None):
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    res = ctx.solve_batch(w.batch, opts)
    ora = [O.solve_lm(p) for p in problems_from_workload(w)]
    return res, ora


def test_solve_parity_zero_field(ctx):
    """The reference-equivalent case: zero cost field, so nearest-node, central-difference
    and trilinear lookups coincide.  Final joint trajectories within 1e-4 rad of the oracle (north-star tolerance)."""
    w = small_workload("C2", "panda_small", B=6, n_field=64)
    w.batch.field_all[:] = -1
    w.batch.field_obs[:] = -1
    res, ora = _solve_both(ctx, w)
    for i, r in enumerate(ora):
        assert r.status == O.STATUS_CONVERGED and res["status"][i] == capi.STATUS_CONVERGED
        assert np.abs(res["Q"][i] - r.Q).max() < 1e-4, (i, np.abs(res["Q"][i] - r.Q).max())
        assert np.abs(res["dQ"][i] - r.dQ).max() < 1e-3
        assert res["cost"][i] == pytest.approx(r.cost, rel=1e-4)
    # independent solver on the same residuals (SciPy TRF) agrees as well
    p0 = problems_from_workload(w, [0])[0]
    Qs, cs, _ = O.solve_scipy(p0)
    assert np.abs(res["Q"][0] - Qs).max() < 1e-4


def test_solve_parity_tabletop_field(ctx):
    """Active obstacle term (trilinear value + gradient): CUDA float32 path vs the same algorithm in float64."""
    w = small_workload("C2", "panda_small", B=8, n_field=64)
    res, ora = _solve_both(ctx, w)
    dev = np.array([np.abs(res["Q"][i] - r.Q).max() for i, r in enumerate(ora)])
    rel = np.array([abs(res["cost"][i] - r.cost) / r.cost for i, r in enumerate(ora)])
    print("max |dQ| per problem", dev, "rel cost", rel, "iters", res["iters"], [r.iters for r in ora])
    assert np.all(rel < 1e-3)
    assert np.all(dev < 1e-4), dev


def test_solve_properties_full_c2(ctx):
    """BASELINE C2 at full size (256 x 30 knots x 2000 points x 128^3): size-independent properties."""
    w = W.make_workload("C2")
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    res = ctx.solve_batch(w.batch)
    t, b = w.table, w.batch
    Q = res["Q"]
    oi = t.opt_qidx
    st = res["status"]
    assert np.mean(st == capi.STATUS_CONVERGED) > 0.85  # step (under light damping) / gradient criterion
    assert not np.any((st == capi.STATUS_NAN) | (st == capi.STATUS_STALLED))
    np.testing.assert_array_equal(Q[:, 0, oi], b.qc[:, oi])  # initial configuration
    np.testing.assert_array_equal(Q[:, 1, oi], b.qc[:, oi])  # zero initial velocity
    assert np.all(Q[:, :, oi] >= t.lo - 1e-12) and np.all(Q[:, :, oi] <= t.hi + 1e-12)
    np.testing.assert_array_equal(Q[:, :, t.par_qidx], b.q_seed[:, :, t.par_qidx])  # parameter joints untouched
    np.testing.assert_allclose(res["dQ"][:, :, oi], np.diff(Q[:, :, oi], axis=1) / b.dt, atol=1e-12)
    assert np.all(res["dQ"][:, :, t.par_qidx] == 0)
    # reported cost == objective re-evaluated by the oracle at the returned trajectory (sample)
    for i in (0, 17, 255):
        p = problems_from_workload(w, [i])[0]
        assert res["cost"][i] == pytest.approx(O.total_cost(p, Q[i]), rel=2e-4)
    # solving twice gives identical bits (deterministic reduction order)
    res2 = ctx.solve_batch(w.batch)
    np.testing.assert_array_equal(res2["Q"], Q)
    # the packed float32 result that feeds the all-gather matches
    assert ctx.profile()["linearize_launches"] > 0


def test_chunked_solve_is_identical(ctx, monkeypatch):
    """A Jacobian-row budget smaller than the batch splits the solve into chunks (C5 at full size needs two): same bits, and
    the per-launch time stamps of all chunks add up to less than the solve."""
    w = small_workload("C2", "panda_small", B=8, n_field=64)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    ref = ctx.solve_batch(w.batch)
    per_problem_mb = (w.batch.T * w.table.npoints + 6 * w.table.grip_pt_count) * (w.table.nopt + 1) * 4 / 2**20
    ctx.configure(jrows_budget_mb=3.5 * per_problem_mb)  # 3 problems per chunk -> chunks of 3, 3, 2
    try:
        res = ctx.solve_batch(w.batch)
        pf = ctx.profile()
    finally:
        ctx.configure(jrows_budget_mb=24576)
    for k in ("Q", "dQ", "cost", "iters", "status"):
        np.testing.assert_array_equal(res[k], ref[k], err_msg=k)
    assert 0 < pf["linearize_ms"] + pf["step_ms"] <= pf["solve_ms"] * 1.05


@pytest.mark.parametrize("knobs", [{"fused": 1}, {"step_fk": 1 << 30}, {"step_fk": 4}, {"pdl": 0}, {"launch_events": 1}, {"cull_nslot": 2},
                                   {"cons_warps": 5}])
def test_launch_options_do_not_change_the_result(ctx, knobs):
    """The persistent one-CTA-per-problem solver (k_solve_fused, "fused" = 1), the optional launch structures of the default path
    (FK records written by the step kernel, plain stream-ordered launches, CUDA events between the launches) and a shorter
    brick ring must reproduce the default path (k_item_fk -> k_linearize_cull -> k_step_cr per iteration) bit for bit; a different
    number of consumer warps changes the summation order of the per-knot blocks only."""
    w = small_workload("C2", "panda_small", B=12, n_field=64)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    ref = ctx.solve_batch(w.batch)
    default = dict(fused=0, step_fk=0, pdl=1, launch_events=0, cull_nslot=4, cons_warps=0)
    ctx.configure(**knobs)
    try:
        res = ctx.solve_batch(w.batch)
    finally:
        ctx.configure(**default)
    if "cons_warps" in knobs:
        ok = (ref["status"] == 0) & (res["status"] == 0)
        assert ok.sum() >= 8 and np.abs(res["Q"][ok] - ref["Q"][ok]).max() < 1e-4
    else:
        for k in ("Q", "dQ", "cost", "iters", "status"):
            np.testing.assert_array_equal(res[k], ref[k], err_msg=f"{knobs} {k}")


def test_contexts_in_flight_give_identical_results(ctx):
    """Several contexts solving concurrently from their own host threads (what bench.py --in-flight does: the tail of one batch
    overlaps the head of the next) return the bits of a context running alone, and the bundle option 0 / radius 0 are the plain
    Levenberg-Marquardt path."""
    import threading

    w = small_workload("C2", "panda_small", B=24, n_field=64)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    ref = ctx.solve_batch(w.batch)
    others = [capi.GtoContext(0) for _ in range(3)]
    out, errs = {}, []

    def work(i, c):
        try:
            c.set_robot(w.table)
            upload_fields(c, w)
            for _ in range(3):
                out[i] = c.solve_batch(w.batch)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(others)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in others:
        c.close()
    assert not errs, errs
    for i in range(3):
        for k in ("Q", "dQ", "cost", "iters", "status"):
            np.testing.assert_array_equal(out[i][k], ref[k], err_msg=f"context {i} {k}")
    o0 = capi.default_options()
    o0.bundle = 0
    o1 = capi.default_options()
    o1.bundle_radius = 0.0
    r0, r1 = ctx.solve_batch(w.batch, o0), ctx.solve_batch(w.batch, o1)
    for k in ("Q", "iters", "status"):
        np.testing.assert_array_equal(r0[k], r1[k])
    assert np.any(r0["iters"] != ref["iters"])  # (the bundle does change the path of some problems)


def _sub_batch(b, idx):
    import dataclasses
    idx = np.asarray(idx, dtype=np.int64)
    return dataclasses.replace(b, qc=b.qc[idx], q_seed=b.q_seed[idx], goal_tf=b.goal_tf[idx], base_position=b.base_position[idx],
                               field_all=b.field_all[idx], field_obs=b.field_obs[idx])


def test_single_and_ragged_batches_equal_the_full_batch(ctx):
    """Edge cases of the batch dimension: one problem, and an odd-sized subset in a different order, give the same bits as
    the same problems inside a larger batch (problems never interact)."""
    w = small_workload("C2", "panda_small", B=7, n_field=48)
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    full = ctx.solve_batch(w.batch)
    one = ctx.solve_batch(_sub_batch(w.batch, [3]))
    assert np.array_equal(one["Q"][0], full["Q"][3]) and one["iters"][0] == full["iters"][3] and one["cost"][0] == full["cost"][3]
    idx = [6, 0, 4]
    rag = ctx.solve_batch(_sub_batch(w.batch, idx))
    for k, i in enumerate(idx):
        assert np.array_equal(rag["Q"][k], full["Q"][i]) and rag["status"][k] == full["status"][i]


def test_error_behaviour_of_the_boundary(ctx):
    """Every misuse returns a negative code with a message (raised as GtoError by the binding); nothing crashes or falls back."""
    import dataclasses
    w = small_workload("C2", "panda_small", B=2, n_field=48)
    fresh = capi.GtoContext(0)
    with pytest.raises(capi.GtoError) as e:  # robot not set
        fresh.solve_batch(w.batch)
    assert e.value.code == -4
    fresh.close()
    ctx.set_robot(w.table)
    upload_fields(ctx, w)
    with pytest.raises(capi.GtoError) as e:  # field slot never uploaded
        ctx.solve_batch(dataclasses.replace(w.batch, field_obs=np.full(2, 4000, np.int32)))
    assert e.value.code == -1
    with pytest.raises(capi.GtoError) as e:  # stand-off knot outside the trajectory
        ctx.solve_batch(dataclasses.replace(w.batch, standoff_offset=-(w.batch.T + 1)))
    assert e.value.code == -1
    with pytest.raises(capi.GtoError) as e:  # too few knots: the first two are pinned to qc
        ctx.solve_batch(dataclasses.replace(_sub_batch(w.batch, [0, 1]), T=2, q_seed=w.batch.q_seed[:, :2], standoff_offset=-1))
    assert e.value.code == -1
    with pytest.raises(capi.GtoError):  # empty batch
        ctx.solve_batch(_sub_batch(w.batch, []))
    res = ctx.solve_batch(w.batch)  # the context is still usable afterwards
    assert res["Q"].shape == (2, w.batch.T, w.table.ndof) and np.all(np.isfinite(res["Q"]))
