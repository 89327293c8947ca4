"""SURVEY.md section 8(f) row 4, measured: gto_base_place (k_base_place) -- 4096 base-placement problems x 10 goals in one launch,
against oracle/base_oracle.c (the same algorithm in C, all host threads, bounded sample) and oracle/base_oracle.py (NumPy, 8 problems).  `python tools/bench_base.py` prints one JSON
object; `--once NAME` runs a single launch (for ncu)."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
from grasptrajopt_b200 import capi


def run(cases=(("fetch_small", 4096, 10), ("panda_small", 4096, 10)), cpu_sample=8, reps=3):
    import base_oracle as BO
    from helpers import base_placement_case
    out = {}
    for name, Bn, n in cases:
        tb, qcb, RTb, grid, origin, res = base_placement_case(name, B=Bn, n=n, seed=21, spread=0.5)
        ctx = capi.GtoContext(0)
        ctx.set_robot(tb)
        ctx.base_place(qcb, RTb[:64], 0.01, grid, origin, res)
        t0 = time.perf_counter()
        for _ in range(reps):
            ob = ctx.base_place(qcb, RTb, 0.01, grid, origin, res)
        dtb = (time.perf_counter() - t0) / reps
        rec = {"problems": Bn, "goals_per_problem": n, "kernel_ms": ob["kernel_ms"], "call_ms_e2e": 1e3 * dtb, "problems_per_s_e2e": Bn / dtb,
               "goal_iterations_per_s_kernel": float(ob["iters"].sum()) * n / (ob["kernel_ms"] * 1e-3),
               "converged": int((ob["status"] == 0).sum()), "at_max_iter": int((ob["status"] == 1).sum()), "iters_mean": float(ob["iters"].mean()),
               "collision_free": int((ob["collision"] == 0).sum())}
        if cpu_sample:
            import c_oracle as CO
            ncs = min(Bn, 1024)  # bounded sample for the C port on all host threads
            CO.base_place(tb, qcb, RTb[:8], 0.01, grid, origin, res)
            t0 = time.perf_counter()
            rc = CO.base_place(tb, qcb, RTb[:ncs], 0.01, grid, origin, res)
            dtc = time.perf_counter() - t0
            t0 = time.perf_counter()
            ro = [BO.solve_base(BO.BaseProblem(tb, qcb, RTb[i], 0.01, grid, origin, res)) for i in range(cpu_sample)]
            dto = (time.perf_counter() - t0) / cpu_sample
            rec["cpu_baseline"] = {"value": ncs / dtc, "unit": "placements/s", "cores": int(rc["threads"]), "kind": "port",
                                   "sample": f"{ncs} of {Bn} problems, oracle/base_oracle.c (same projected LM, float64), {int(rc['threads'])} pthreads, {dtc:.2f} s",
                                   "max_abs_dy_vs_gpu": float(np.abs(rc["y"] - ob["y"][:ncs]).max()),
                                   "iters_equal_frac": float((rc["iters"] == ob["iters"][:ncs]).mean())}
            rec["numpy_oracle"] = {"problems_per_s": 1.0 / dto, "cores": 1, "sample": cpu_sample,
                                   "max_abs_dy_vs_gpu": float(max(np.abs(ro[i].y - ob["y"][i]).max() for i in range(cpu_sample))),
                                   "iters_equal": bool(all(ro[i].iters == ob["iters"][i] for i in range(cpu_sample)))}
        out[f"base_placement_{name}"] = rec
        ctx.close()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--once":
        print(json.dumps(run(cases=((sys.argv[2], 4096, 10),), cpu_sample=0, reps=1)))
    else:
        print(json.dumps(run()))
