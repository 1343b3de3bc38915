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


def pairwise_inter_matrix_sharded(local_tracks, group=None) -> torch.Tensor:
    """BASELINE config 5 (one video with more candidate tracks than one GPU should binarise): every rank holds the packed
    planes of ITS tracks (n_local, T, H, Wp) — equal n_local on every rank.  The path has a real exchange step here:
      1. NCCL all-gather of the packed tracks (1/32 of the mask bytes; 13.3 GB in total for 256 x 200 x 1080p),
      2. each rank computes its share of the 64 x 64 pair tiles of the upper triangle (sola_pair_iou_st_part),
      3. one all-reduce(SUM) of the int64 N x N matrix (512 KB at N = 256).
    Returns the full symmetric matrix on every rank; integer sums, so the result is identical for any world size."""
    from . import packed as P
    w = local_tracks.words.contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return P.pairwise_inter_matrix(local_tracks)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    full = torch.empty((world * w.shape[0], *w.shape[1:]), dtype=w.dtype, device=w.device)
    dist.all_gather_into_tensor(full, w, group=group)
    inter = P.pairwise_inter_matrix_part(P.PackedMasks(full, local_tracks.H, local_tracks.W), rank, world)
    dist.all_reduce(inter, op=dist.ReduceOp.SUM, group=group)
    return inter


def pair_tile_owner(n_tracks: int, world: int, tile: int = 64):
    """Host mirror of the kernel's tile assignment: list of (ti, tj, owner rank) over the upper triangle."""
    nt = (n_tracks + tile - 1) // tile
    out, idx = [], 0
    for ti in range(nt):
        for tj in range(ti, nt):
            out.append((ti, tj, idx % world))
            idx += 1
    return out
