"""Two (or more) GPUs: a goal set sharded over ranks, each rank solves its goals through the C-ABI, the winner is found with
distributed.goalset_argmin (one NCCL all-gather) and compared with the single-GPU solve of the whole set on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_goalset_nccl.py"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist
from grasptrajopt_b200 import capi, workloads as W
from grasptrajopt_b200.distributed import goalset_argmin, shard_range

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
w = W.make_workload("C2", scale=24 / 256, n_field=64)  # 24 candidate grasps = one goal set
ctx = capi.GtoContext(local)
ctx.set_robot(w.table)
for slot, cf in w.fields.items():
    ctx.set_field(slot, cf.cost, cf.origin, cf.pitch)
lo, hi = shard_range(w.batch.B, rank, world)
res = ctx.solve_batch(W.slice_batch(w.batch, lo, hi))
idx, cost, owner = goalset_argmin(res["cost"], lo, world, res["status"])
out = {"rank": rank, "winner": idx, "cost": cost, "owner": owner}
if rank == 0:
    full = ctx.solve_batch(w.batch)
    c = np.where(full["status"] == 0, full["cost"], np.inf)
    out["single_gpu_winner"] = int(np.argmin(c))
    out["single_gpu_cost"] = float(full["cost"][int(np.argmin(c))])
    out["match"] = bool(out["single_gpu_winner"] == idx and out["single_gpu_cost"] == cost)
    print(json.dumps(out))
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
