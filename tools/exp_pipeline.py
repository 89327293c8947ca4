"""Experiment: throughput of K solves of a configuration with 1..4 batches in flight (one context + stream + host thread each)."""
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from grasptrajopt_b200 import capi, workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
w = W.make_workload(cfg)
ctxs = []
for i in range(8):
    c = capi.GtoContext(0)
    c.set_robot(w.table)
    for slot, cf in w.fields.items():
        c.set_field(slot, cf.cost, cf.origin, cf.pitch)
    c.upload_batch(w.batch)
    c.solve_resident()
    ctxs.append(c)
for nf in (1, 2, 3, 4, 6, 8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lat = []

    def work(c, n):
        for _ in range(n):
            c.solve_resident()
            lat.append(c.profile()["solve_ms"])

    th = [threading.Thread(target=work, args=(ctxs[i], steps // nf)) for i in range(nf)]
    e0.record()
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n = (steps // nf) * nf
    print(f"{cfg} in flight {nf}: {n} solves, device {e0.elapsed_time(e1):.2f} ms, wall {1e3 * wall:.2f} ms, per solve {e0.elapsed_time(e1) / n:.3f} ms, "
          f"solve latency mean {np.mean(lat):.2f} ms", flush=True)
for c in ctxs:
    c.close()
