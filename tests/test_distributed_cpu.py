"""N > 1 path on CPU: contiguous sharding + one all-gather of the packed results over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

from grasptrajopt_b200 import distributed as D
from grasptrajopt_b200.workloads import slice_batch


def test_shard_ranges_partition_the_batch():
    for B, world in ((256, 8), (10, 4), (3, 8), (16384, 8)):
        r = [D.shard_range(B, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == B
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    Q = rng.normal(size=(5, 30, 9))
    opt = np.arange(7)
    p = D.pack_result(Q, rng.random(5), np.array([0, 1, 0, 4, 0]), opt)
    assert p.shape == (5, 7 * 30 + 2) and p.dtype == np.float32
    Qx, cost, st = D.unpack_result(p, 30, 7)
    np.testing.assert_allclose(Qx, Q[:, :, :7].astype(np.float32))
    assert st.tolist() == [0, 1, 0, 4, 0]


def _worker(rank, world, port, tmp):
    import sys
    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here]
    import c_oracle
    from helpers import small_workload

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = small_workload("C2", "panda_small", B=5, n_field=32)  # ragged: 3 + 2
    lo, hi = D.shard_range(w.batch.B, rank, world)
    counts = [D.shard_range(w.batch.B, r, world)[1] - D.shard_range(w.batch.B, r, world)[0] for r in range(world)]
    w_local = type(w)(**{**w.__dict__, "batch": slice_batch(w.batch, lo, hi), "RT": w.RT[lo:hi], "q_star": w.q_star[lo:hi]})
    res = c_oracle.solve_workload(w_local, nthreads=1)  # stands in for the GPU solve of this shard
    packed = torch.from_numpy(D.pack_result(res["Q"], res["cost"], res["status"], w.table.opt_qidx))
    allres = D.all_gather_results(packed, world, counts)
    if rank == 0:
        full = c_oracle.solve_workload(w, nthreads=1)
        ref = D.pack_result(full["Q"], full["cost"], full["status"], w.table.opt_qidx)
        np.save(os.path.join(tmp, "ok.npy"), np.array([np.array_equal(allres.numpy(), ref)]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_plus_all_gather_equals_full_batch(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy")[0]


def _argmin_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from grasptrajopt_b200.distributed import goalset_argmin, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    cost = rng.uniform(1.0, 2.0, 11)
    status = np.zeros(11, np.int32)
    cost[7] = 0.25  # global minimum, on rank 1 ...
    status[7] = 4   # ... but not converged
    cost[2] = 0.5   # best converged one, on rank 0
    lo, hi = shard_range(11, rank, world)
    got = goalset_argmin(torch.from_numpy(cost[lo:hi]), lo, world, torch.from_numpy(status[lo:hi]))
    got_all = goalset_argmin(torch.from_numpy(cost[lo:hi]), lo, world)  # without statuses: plain arg-min
    none = goalset_argmin(torch.from_numpy(cost[lo:hi]), lo, world, torch.ones(hi - lo, dtype=torch.int32))  # nobody converged
    np.save(os.path.join(tmp, f"argmin{rank}.npy"), np.array([got[0], got[2], got_all[0], got_all[2], none[0]]))
    dist.barrier()
    dist.destroy_process_group()


def test_goalset_argmin_across_ranks(tmp_path):
    """plan_goalset over a sharded goal set: local arg-min + one all-gather of (cost, index) pairs; every rank gets the same winner."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_argmin_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "argmin0.npy"), np.load(tmp_path / "argmin1.npy")
    assert np.array_equal(a, b)
    assert list(a) == [2, 0, 7, 1, 7]
