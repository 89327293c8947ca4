"""Experiment (1 GPU): bench.py on the shard of every rank of an 8-rank weak-scaling run, one after the other -- separates the
difficulty of the differently seeded shards from the cost of running N ranks."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--extras", "0", "--cpu-sample", "0", "--steps", "12", "--shard-seed-of-rank", str(r)],
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    c = d["config"]
    print(f"shard of rank {r}: value {d['value']:.0f} traj/s, {d['ms_per_step']:.3f} ms per batch ({c['batches_in_flight']} in flight), converged {c['converged']}, "
          f"one batch alone {c['solve_ms_per_step']:.2f} ms, problems with >= 60 iterations: {sum(c['iterations_histogram'][60:])}", flush=True)
