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
