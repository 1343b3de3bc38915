"""Multi-GPU plumbing: the path shards by independent units (a video for the grid filter, a (video, expression) for
the gdino filter and for J&F), exactly like the reference's `video_idx % n_pid == pid` process sharding
(generate_prompts_gdino.py:114, generate_prompts_grid.py:72).  There is no data-path collective; the only exchange is
ONE all-reduce of the J/F accumulators at the end of a sweep (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_units: int, rank: int, world: int) -> List[int]:
    """Round-robin: unit i belongs to rank i % world (the reference's modulo rule)."""
    return list(range(rank, n_units, world))


def shard_balanced(costs: Sequence[float], rank: int, world: int) -> List[int]:
    """Longest-processing-time assignment by cost (e.g. T*H*W) for sweeps with very uneven units; deterministic, so
    every rank computes the same partition without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    mine: List[int] = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += costs[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def init_process_group_from_env(device: torch.device | None = None) -> Tuple[int, int]:
    """torchrun-style init: NCCL when a CUDA device is given, gloo otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = "nccl" if (device is not None and device.type == "cuda") else "gloo"
        kw = {"device_id": device} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


def allreduce_jf(sum_J: float, sum_F: float, sum_JF: float, n_units: int, int_totals: np.ndarray, device=None):
    """Sum the per-rank accumulators: float64 [ΣJ, ΣF, ΣJF] and int64 [n_units, Σinter, Σ|pred|, Σ|gt|].
    Integer sums are order-independent, so the integer audit is bit-identical for any world size; the float64 sums
    agree to ~1e-12."""
    f = torch.tensor([sum_J, sum_F, sum_JF], dtype=torch.float64)
    i = torch.tensor([n_units, *[int(x) for x in np.asarray(int_totals).tolist()]], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            f, i = f.to(dev), i.to(dev)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        dist.all_reduce(i, op=dist.ReduceOp.SUM)
        f, i = f.cpu(), i.cpu()
    n = int(i[0])
    means = (f / max(n, 1)).tolist()
    return {"mean_J": means[0], "mean_F": means[1], "mean_JF": means[2], "n_units": n,
            "int_totals": i[1:].numpy().copy()}
