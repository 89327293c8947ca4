"""Multi-GPU plumbing: problems are independent, so a batch is split into contiguous shards (one process per GPU,
``torch.distributed``), each rank solves its shard with no communication, and the converged trajectories are exchanged
with ONE all-gather of the packed float32 results ``[B_local, nopt*T + 2]`` = (optimised rows of Q knot-major, cost,
status) -- SURVEY.md section 8(e).  On GPUs the payload is the library's own device buffer (``gto_result_device_ptr``) and the
backend is NCCL over NVLink; the same code runs with ``gloo`` on CPU tensors in the tests."""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous equal split of [0, B) (remainder to the first ranks)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_result(Q: np.ndarray, cost: np.ndarray, status: np.ndarray, opt_qidx) -> np.ndarray:
    """Host-side twin of the packing k_finalize does on the device."""
    B, T = Q.shape[0], Q.shape[1]
    out = np.zeros((B, len(opt_qidx) * T + 2), dtype=np.float32)
    out[:, : len(opt_qidx) * T] = Q[:, :, opt_qidx].reshape(B, -1)
    out[:, -2] = cost
    out[:, -1] = status
    return out


def unpack_result(packed: np.ndarray, T: int, nopt: int):
    packed = np.asarray(packed)
    B = packed.shape[0]
    return packed[:, : nopt * T].reshape(B, T, nopt), packed[:, -2], packed[:, -1].astype(np.int32)


class DeviceArray:
    """Zero-copy view of a raw device pointer for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def all_gather_results(local, world: int, counts=None):
    """All-gather of per-rank packed results (torch tensors, equal or ragged shard sizes).  Returns [sum(B_r), nfl]."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    mx = max(counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx : r * mx + counts[r]] for r in range(world)], dim=0)


def goalset_argmin(local_cost, lo: int, world: int, local_status=None):
    """Goal-set semantics across ranks (the reference's ``mmin`` over per-goal costs,
    ``gto/gto_planner.py:105``, is an arg-min after the solve): every rank reduces its own shard to one ``(cost, global index, flag)``
    triple, the triples are exchanged with one all-gather of ``[world, 3]`` float64 (cost, index, not-converged flag), and every rank returns the same winner
    ``(global index, cost, owner rank)``.  ``lo`` is the global index of this rank's first problem (``shard_range``).  Problems
    whose ``local_status`` is not 0 (converged) only win when no rank has a converged one; ties go to the lowest global index."""
    import torch
    import torch.distributed as dist

    from .goalset import goalset_best, status_class

    cost = torch.as_tensor(local_cost, dtype=torch.float64)
    dev = cost.device
    n = cost.shape[0]
    big = torch.finfo(torch.float64).max
    if n == 0:
        pair = torch.tensor([big, -1.0, 3.0], dtype=torch.float64, device=dev)
    else:  # the same ranking as the single-GPU planner (goalset.py): converged first, never a NaN solve
        c_np = cost.detach().cpu().numpy()
        s_np = np.zeros(n, np.int64) if local_status is None else torch.as_tensor(local_status).detach().cpu().numpy()
        i = goalset_best(c_np, s_np)
        ci = float(c_np[i]) if np.isfinite(c_np[i]) else big
        pair = torch.tensor([ci, float(lo + i), float(status_class(c_np[i : i + 1], s_np[i : i + 1])[0])], dtype=torch.float64, device=dev)
    if world > 1:
        if dist.get_backend() == "nccl" and not pair.is_cuda:  # host-side costs (the C-ABI returns NumPy arrays): NCCL needs device memory
            pair = pair.to(torch.device("cuda", torch.cuda.current_device()))
        allp = torch.empty((world, 3), dtype=torch.float64, device=pair.device)
        dist.all_gather_into_tensor(allp, pair.reshape(1, 3).contiguous())
    else:
        allp = pair.reshape(1, 3)
    allp = allp.cpu().numpy()
    # status class first (converged, finite but not converged, NaN), then cost, then index
    order = sorted(range(world), key=lambda r: (allp[r, 2], allp[r, 0], allp[r, 1] if allp[r, 1] >= 0 else np.inf))
    r = order[0]
    return int(allp[r, 1]), float(allp[r, 0]), int(r)
