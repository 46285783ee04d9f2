"""Chain-parallel sharding over GPUs (one process per GPU, torch.distributed).

Chains are independent (they share only the read-only model), so there is no collective in the hop
loop: rank ``k`` runs a contiguous block of global chain ids, and because the Philox streams are keyed
by the *global* chain id the results do not depend on the world size.  The only exchange is the final
gather of acceptance counts and Hamming histograms ("K7").
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_chains: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the chains owned by ``rank``: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n_chains), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_visit_histogram(visit_hist, group=None):
    """Sum of the ranks' visit histograms (``visit_hist(2^n)``, ncclAllReduce).  ``visit_hist`` is
    this rank's int64[2^n] -- a CUDA tensor straight from the device accumulator (nccl: reduced in place over NVLink,
    no host copy) or a numpy array / CPU tensor (gloo).  Returns the same kind of object."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    if isinstance(visit_hist, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(visit_hist, dtype=np.int64))
        if backend == "nccl":
            t = t.to(torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    t = visit_hist
    if backend == "nccl" and not t.is_cuda:
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_statistics(acc_count: np.ndarray, hamming_hist: np.ndarray, n_chains: int, group=None, visit_hist=None):
    """All-gather the per-chain acceptance counts [C_local] and Hamming histograms [C_local, n+1]
    into global arrays ordered by chain id.  Works on CPU tensors (gloo) and CUDA tensors (nccl).
    With ``visit_hist`` (this rank's int64[2^n]) a third result is returned: its sum over the ranks (one all-reduce)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = hamming_hist.shape[1]
    cap = max(shard_range(n_chains, r, world)[1] - shard_range(n_chains, r, world)[0] for r in range(world))
    pad = np.zeros((cap, width + 1), dtype=np.int64)
    pad[: len(acc_count), 0] = acc_count
    pad[: len(acc_count), 1:] = hamming_hist
    mine = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    acc = np.empty(n_chains, dtype=np.int64)
    hist = np.empty((n_chains, width), dtype=np.int64)
    for r in range(world):
        a, b = shard_range(n_chains, r, world)
        blk = out[r].cpu().numpy()
        acc[a:b] = blk[: b - a, 0]
        hist[a:b] = blk[: b - a, 1:]
    if visit_hist is not None:
        return acc, hist, reduce_visit_histogram(visit_hist, group)
    return acc, hist


def run_chains_sharded(run_local: Callable[[int, int], tuple], n_chains: int, group=None):
    """Run ``run_local(chain_id0, count) -> (acc_count, hamming_hist, payload)`` on this rank's shard and
    gather the statistics.  Returns (acc_count[C], hamming_hist[C, n+1], local payload)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(n_chains, rank, world)
    acc, hist, payload = run_local(a, b - a)
    g_acc, g_hist = gather_statistics(np.asarray(acc), np.asarray(hist), n_chains, group)
    return g_acc, g_hist, payload


def combine_logsumexp(pairs: np.ndarray) -> float:
    """log sum_k s_k exp(m_k) for (m_k, s_k) partial results of disjoint index ranges."""
    pairs = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
    keep = pairs[:, 1] > 0
    if not keep.any():
        return float("-inf")
    m = pairs[keep, 0].max()
    return float(m + np.log(np.sum(pairs[keep, 1] * np.exp(pairs[keep, 0] - m))))


def exact_logz_sharded(model, beta: float, device: int = 0, group=None, partial_fn=None) -> float:
    """log Z of the Boltzmann distribution with the 2^n indices split evenly over the ranks: every rank
    enumerates its range on its GPU (qmc_exact_partial), the (max, sumexp) pairs are all-gathered and
    combined on every rank.  ``partial_fn(begin, count) -> (max, sumexp)`` replaces the CUDA call in the
    CPU (gloo) tests."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, end = shard_range(1 << model.num_spins, rank, world)
    if partial_fn is None:
        from . import _device, _lib
        dm = model.device_model(device)
        pair = torch.empty(2, dtype=torch.float64, device=torch.device("cuda", device))
        _lib.check(_lib.lib().qmc_exact_partial(dm.handle, float(beta), int(begin), int(end - begin), pair.data_ptr(),
                                                None, _device.current_stream_ptr(device)))
        mine = pair
    else:
        mine = torch.tensor(partial_fn(begin, end - begin), dtype=torch.float64)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return combine_logsumexp(np.stack([o.cpu().numpy() for o in out]))
