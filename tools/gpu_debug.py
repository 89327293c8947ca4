"""First-contact diagnostics on the GPU box: runs each stage separately and prints what it sees (never raises on a
numerical mismatch), so that ONE gpurun call tells as much as possible."""
import os, sys, time, traceback
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import gto_oracle as O
from grasptrajopt_b200 import capi, workloads as W
from helpers import problems_from_workload, small_workload, upload_fields


def stage(name):
    print(f"\n===== {name} =====", flush=True)


def eval_compare(ctx, w, flags, label):
    w.batch.flags = flags
    t0 = time.time()
    out = ctx.eval_batch(w.batch)
    t1 = time.time()
    n = w.table.nopt
    ps = problems_from_workload(w)
    for i, p in enumerate(ps[:2]):
        lin = O.linearize(p, p.q_seed)
        rows = O.pack_rows(p, lin)
        er = np.abs(out["rows"][i][:, n] - rows[:, n]).max()
        ej = np.abs(out["rows"][i][:, :n] - rows[:, :n])
        bad = (ej.max(axis=1) > 1e-4 * (1 + np.abs(rows[:, :n]).max(axis=1)))
        eH = np.abs(out["H"][i] - lin.H).max() / (np.abs(lin.H).max() + 1e-9)
        eg = np.abs(out["g"][i] - lin.g).max() / (np.abs(lin.g).max() + 1e-9)
        ec = np.abs(out["cost"][i] - lin.cost_pts).max() / (np.abs(lin.cost_pts).max() + 1e-9)
        print(f"[{label}] prob {i}: |dr|max {er:.2e}  |dJ|max {ej.max():.2e} bad rows {bad.sum()}/{bad.size} first bad {np.nonzero(bad)[0][:5]}  relH {eH:.2e} relg {eg:.2e} relcost {ec:.2e}  (eval {t1-t0:.2f}s)", flush=True)
    return out


def main():
    ctx = capi.GtoContext(0)
    for cfg, tab, B, nf in [("C2", "panda_small", 3, 64), ("C3", None, 2, 96), ("C4", None, 2, 64)]:
        stage(f"eval parity {cfg} {tab}")
        try:
            w = small_workload(cfg, tab, B=B, n_field=nf)
            ctx.set_robot(w.table)
            upload_fields(ctx, w)
            outs = {}
            for flags, label in [(capi.FLAG_NO_BRICK, "global"), (capi.FLAG_V1_KERNEL, "tma-brick"), (0, "pipe")]:
                try:
                    outs[label] = eval_compare(ctx, w, flags, label)
                except Exception:
                    traceback.print_exc()
            if "global" in outs:
                for k in ("tma-brick", "pipe"):
                    if k in outs:
                        print(k, "identical to global:", np.array_equal(outs[k]["rows"], outs["global"]["rows"]),
                              "max diff", np.abs(outs[k]["rows"] - outs["global"]["rows"]).max(), flush=True)
            w.batch.flags = 0
        except Exception:
            traceback.print_exc()

    stage("solve parity small (zero field, then tabletop field)")
    try:
        for zero in (True, False):
            w = small_workload("C2", "panda_small", B=4, n_field=64)
            if zero:
                w.batch.field_all[:] = -1
                w.batch.field_obs[:] = -1
            ctx.set_robot(w.table)
            upload_fields(ctx, w)
            t0 = time.time()
            res = ctx.solve_batch(w.batch)
            print("gpu solve %.3fs" % (time.time() - t0), "status", res["status"], "iters", res["iters"], "cost", res["cost"], flush=True)
            print("profile", ctx.profile(), flush=True)
            for i, p in enumerate(problems_from_workload(w)):
                r = O.solve_lm(p)
                print(f"  prob {i}: oracle iters {r.iters} status {r.status} cost {r.cost:.8f} | gpu cost {res['cost'][i]:.8f} | max|dQ| {np.abs(res['Q'][i]-r.Q).max():.3e}", flush=True)
    except Exception:
        traceback.print_exc()

    stage("full C2 solve")
    try:
        w = W.make_workload("C2")
        ctx.set_robot(w.table)
        upload_fields(ctx, w)
        for rep in range(3):
            t0 = time.time()
            res = ctx.solve_batch(w.batch)
            dt = time.time() - t0
            pr = ctx.profile()
            print(f"rep {rep}: wall {dt*1e3:.1f} ms  conv {np.mean(res['status']==0):.3f} iters mean {res['iters'].mean():.1f} max {res['iters'].max()}  status hist {np.bincount(res['status'], minlength=4)}", flush=True)
            print("   profile", pr, flush=True)
        w.batch.flags = capi.FLAG_NO_JROWS
        res = ctx.solve_batch(w.batch)
        print("no-jrows profile", ctx.profile(), flush=True)
        w.batch.flags = capi.FLAG_V1_KERNEL
        res = ctx.solve_batch(w.batch)
        print("v1-kernel profile", ctx.profile(), flush=True)
        w.batch.flags = capi.FLAG_NO_BRICK
        res = ctx.solve_batch(w.batch)
        print("no-brick profile", ctx.profile(), flush=True)
    except Exception:
        traceback.print_exc()
    ctx.close()


if __name__ == "__main__":
    main()
