"""GPU: gto_base_place (k_base_place) against the float64 oracle of the reference's BasePlanner."""
import numpy as np
import pytest

import base_oracle as BO
from grasptrajopt_b200 import capi
from helpers import base_placement_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,B,n", [("panda_small", 6, 4), ("fetch_small", 4, 5), ("panda_small", 3, 1), ("panda_small", 2, 32),
                                       ("fetch10_c4", 3, 3)])  # the last one: 10 optimised joints -> k_base_place_sm<12>
def test_base_place_matches_oracle(name, B, n):
    t, qc, RTs, grid, origin, res = base_placement_case(name, B=B, n=n, spread=0.5)
    ctx = capi.GtoContext(0)
    ctx.set_robot(t)
    out = ctx.base_place(qc, RTs, 0.01, grid, origin, res)
    for b in range(B):
        r = BO.solve_base(BO.BaseProblem(t, qc, RTs[b], 0.01, grid, origin, res))
        print(name, b, "gpu", out["status"][b], out["iters"][b], out["cost"][b], out["y"][b], "oracle", r.status, r.iters, r.cost, r.y,
              "max|dQ|", np.abs(out["Q"][b] - r.Q).max())
        assert out["status"][b] == r.status and out["iters"][b] == r.iters, (out["iters"][b], r.iters)
        # same algorithm, float64 on both sides: metres / radians.  A problem that is still moving at the iteration cap
        # (status 1) is compared at its 100th iterate, where rounding differences have been amplified along the flat valley.
        tol = 1e-6 if r.status == 0 else 1e-4
        assert np.abs(out["y"][b] - r.y).max() < tol
        assert np.abs(out["Q"][b] - r.Q).max() < tol
        assert abs(out["cost"][b] - r.cost) <= 1e-7 * max(1.0, r.cost)
        if np.abs(out["y"][b] - r.y).max() < 1e-9:
            assert out["collision"][b] == r.collision
        else:
            assert abs(out["collision"][b] - r.collision) <= 2  # a point within 1e-6 m of a cell edge may fall either side
        assert np.all(out["Q"][b][:, t.par_qidx] == qc[t.par_qidx])
    ctx.close()


def test_base_place_large_batch_properties():
    """Size-independent properties at a batch the oracle would take minutes for: bounds hold, the reported cost is the
    cost of the returned point and not above the seed's, identical problems give identical bits."""
    t, qc, RTs, grid, origin, res = base_placement_case("panda_small", B=512, n=8, seed=11)
    RTs[256:] = RTs[:256]
    ctx = capi.GtoContext(0)
    ctx.set_robot(t)
    out = ctx.base_place(qc, RTs, 0.01, grid, origin, res)
    assert np.array_equal(out["y"][:256], out["y"][256:]) and np.array_equal(out["Q"][:256], out["Q"][256:])
    assert np.all(np.abs(out["y"][:, 2]) <= np.pi)
    Qx = out["Q"][:, :, t.opt_qidx]
    assert np.all(Qx >= t.lo - 1e-12) and np.all(Qx <= t.hi + 1e-12)
    assert np.all((out["status"] == 0) | (out["status"] == 1))
    seed_x = np.tile(qc[t.opt_qidx], (8, 1))
    for b in range(0, 256, 8):  # the returned iterate never costs more than the seed
        assert out["cost"][b] <= BO.cost_of(BO.BaseProblem(t, qc, RTs[b], 0.01), np.zeros(3), seed_x)
        assert abs(out["cost"][b] - BO.cost_of(BO.BaseProblem(t, qc, RTs[b], 0.01), out["y"][b], out["Q"][b][:, t.opt_qidx])) <= 1e-9 * max(1.0, out["cost"][b])
    assert np.all(out["collision"] >= 0) and np.all(out["collision"] <= t.npoints)
    ctx.close()


def test_base_place_rejects_bad_sizes():
    t, qc, RTs, grid, origin, res = base_placement_case("panda_small", B=1, n=2)
    ctx = capi.GtoContext(0)
    ctx.set_robot(t)
    with pytest.raises(capi.GtoError):
        ctx.base_place(qc, np.tile(np.eye(4), (1, 33, 1, 1)))
    ctx.close()


@pytest.mark.parametrize("name", ["panda_small", "fetch10_c4"])
def test_local_memory_kernel_agrees(name, monkeypatch):
    """k_base_place (one problem per warp, per-thread arrays; GTO_BASE_V1=1, also the path for > 12 optimised joints) and the
    default shared-memory kernel run the same iteration; only the order of the warp / group sums differs."""
    t, qc, RTs, grid, origin, res = base_placement_case(name, B=8, n=4, spread=0.5)
    ctx = capi.GtoContext(0)
    ctx.set_robot(t)
    a = ctx.base_place(qc, RTs, 0.01, grid, origin, res)
    monkeypatch.setenv("GTO_BASE_V1", "1")
    b = ctx.base_place(qc, RTs, 0.01, grid, origin, res)
    monkeypatch.delenv("GTO_BASE_V1")
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"])
    conv = a["status"] == 0
    assert np.abs(a["y"][conv] - b["y"][conv]).max(initial=0.0) < 1e-6 and np.abs(a["y"] - b["y"]).max() < 1e-4
    assert np.abs(a["Q"] - b["Q"]).max() < 1e-4
    assert np.allclose(a["cost"], b["cost"], rtol=1e-7, atol=1e-12)
    ctx.close()
