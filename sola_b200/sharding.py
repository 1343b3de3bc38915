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


def bind_to_gpu_cpus(device_index: int) -> List[int]:
    """Pin this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root), intersected with the CPUs the
    process may use.  One process per GPU: pinned host staging buffers allocated afterwards are first-touched on the GPU's own
    node, so H2D copies of several ranks do not queue on one socket's memory controllers.  Returns the CPU list now in force
    ([] = left unchanged: NVML unavailable, or no overlap with the allowed set)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = sorted(local & allowed)
        if not pick:
            return []
        os.sched_setaffinity(0, pick)
        return pick
    except Exception:
        return []


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


def word_slices(words: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Partition of the word axis used by every K-split path (host mirror of sola_pair_iou_st_rows): rank r owns the 32-word
    stages [stages*r/world, stages*(r+1)/world), i.e. words [32*lo, min(32*hi, words))."""
    stages = (words + align - 1) // align
    out = []
    for r in range(world):
        lo, hi = stages * r // world, stages * (r + 1) // world
        out.append((min(lo * align, words), min(hi * align, words)))
    return out


def _all_to_all(recv: List[torch.Tensor], send: List[torch.Tensor], group=None) -> None:
    """dist.all_to_all where the backend has it (NCCL); gloo (the CPU tests) has no all-to-all, so there it is point-to-point
    isend / irecv pairs (slices of different ranks differ in length, which rules out scatter)."""
    if dist.get_backend(group) != "gloo":
        dist.all_to_all(recv, send, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    recv[rank].copy_(send[rank])
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if send[peer].numel():
            reqs.append(dist.isend(send[peer].contiguous(), dst=g(peer), group=group))
        if recv[peer].numel():
            reqs.append(dist.irecv(recv[peer], src=g(peer), group=group))
    for r in reqs:
        r.wait()


def pairwise_inter_matrix_sharded(local_tracks, group=None, split: str = "words") -> torch.Tensor:
    """BASELINE config 5 (one video with more candidate tracks than one GPU should binarise): every rank holds the packed
    planes of ITS tracks (n_local, T, H, Wp) — equal n_local on every rank.  The path has a real exchange step here.
    split="words" (default): intersections are sums over words, so the WORD axis is partitioned —
      1. NCCL all-to-all: rank r receives words [lo_r, hi_r) of every track (each rank moves 1/world of an all-gather's bytes),
      2. each rank computes the full N x N matrix over its slice (sola_pair_iou_st; balanced for any N),
      3. one all-reduce(SUM) of the int64 N x N matrix (512 KB at N = 256).
    split="tiles": all-gather every track (13.3 GB for 256 x 200 x 1080p native planes), then 64 x 64 pair tiles dealt round-robin.
    Returns the full symmetric matrix on every rank; integer sums, so the result is identical for any world size and split.
    PeerPlanes (below) is the fused version of split="words": no NCCL exchange, the K2 kernel reads the peers over NVLink."""
    from . import packed as P
    w = local_tracks.words.contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return P.pairwise_inter_matrix(local_tracks)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if split == "words":
        n_local = int(w.shape[0])
        flat = w.view(n_local, -1)
        sl = word_slices(int(flat.shape[1]), world)
        send = [flat[:, lo:hi].contiguous() for lo, hi in sl]
        mine = sl[rank][1] - sl[rank][0]
        recv = torch.empty((world, n_local, mine), dtype=w.dtype, device=w.device)
        _all_to_all(list(recv.unbind(0)), send, group=group)
        inter = P.pairwise_inter_matrix_words(recv.view(world * n_local, mine)) if mine > 0 else \
            torch.zeros((world * n_local, world * n_local), dtype=torch.int64, device=w.device)
    else:
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


class PeerPlanes:
    """Packed planes of this rank's tracks in NVLink-mapped symmetric memory (torch.distributed._symmetric_memory), so that every
    GPU can read every rank's tracks directly and the NCCL exchange disappears.  The word axis is partitioned over the ranks
    (word_slices): each rank needs 1/world of every track.  Three ways to consume the peers' memory:
      mode="tma"    (default where the shape allows: gcd(64, n_local) >= 8, world <= 8) ONE kernel is both the exchange and the math:
                    the warp-specialised K2 ring whose producer lane issues its `cp.async.bulk.tensor.2d` tile loads against one tensor
                    map per rank buffer (`sola_pair_iou_st_peer`), so rows cross NVLink inside the kernel that reduces them, tile by
                    tile, with no staging copy and no second launch.  Measured at 2 GPUs, 64 tracks x 200 x 540x960 planes:
                    0.67 ms against 1.23 ms for "pull" and 1.28 ms for "direct" (profiles/r2_cfg5_2gpu_*.log);
      mode="pull"   the rank's slice is walked in chunks; `sola_pull_rows` copies chunk c+1 of all N tracks out of the
                    peers' memory (each remote word crosses NVLink exactly once) on a side stream while the TMA-staged K2 kernel
                    reduces chunk c on the main stream (`sola_pair_iou_st_accumulate`) — transfer and math overlap chunk by chunk;
      mode="direct" one K2 launch whose stage loads (cp.async) read the peers' rows in place (`sola_pair_iou_st_rows`); every row
                    tile is re-read once per pair tile it belongs to, so NVLink carries ~N/128 times the bytes of "pull".

        peers = PeerPlanes(n_local, T, h, w, device)             # collective: allocates + rendezvous once
        S.binarize_pack_resize(logits, resized_out=peers.local)  # producers write straight into the shared buffer
        inter = peers.pairwise_inter_matrix()                    # barrier, exchange fused with K2 over this rank's slice, all-reduce
    """

    def __init__(self, n_local: int, T: int, H: int, W: int, device, group=None, n_chunks: int = 4):
        import torch.distributed._symmetric_memory as symm_mem
        from . import packed as P
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        Wp = (W + 31) // 32
        self.buf = symm_mem.empty((n_local, T, H, Wp), dtype=torch.int32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.local = P.PackedMasks(self.buf, H, W)
        self.words = T * H * Wp
        self.n_tracks = n_local * self.world
        assert self.words % 4 == 0, "each track's planes must be a multiple of 16 bytes"
        ptrs = [int(self.handle.buffer_ptrs[r]) + i * self.words * 4 for r in range(self.world) for i in range(n_local)]
        self.row_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        lo, hi = word_slices(self.words, self.world)[self.rank]
        self.chunks = [(lo + a, lo + b) for a, b in word_slices(hi - lo, max(1, n_chunks)) if b > a]
        cw = max([b - a for a, b in self.chunks], default=0)
        self.scratch = [torch.empty((self.n_tracks, cw), dtype=torch.int32, device=device) for _ in range(2)] if cw else []
        self.side = torch.cuda.Stream(device=device)

    def tma_supported(self) -> bool:
        n_local = self.n_tracks // self.world
        box = 64
        while box > 1 and n_local % box:
            box >>= 1
        return self.world <= 8 and box >= 8

    def pairwise_inter_matrix(self, mode: str = "auto") -> torch.Tensor:
        from . import packed as P
        if mode == "auto":
            mode = "tma" if self.tma_supported() else "pull"
        self.handle.barrier()                 # every rank's planes are written (stream-ordered device barrier over the signal pads)
        if mode == "direct":
            inter = P.pairwise_inter_matrix_rows(self.row_ptrs, self.words, self.rank, self.world)
        elif mode == "tma":               # the K2 producer's TMA loads read the peers' planes in place
            inter = P.pairwise_inter_matrix_peer([int(self.handle.buffer_ptrs[r]) for r in range(self.world)], self.n_tracks // self.world,
                                                 self.words, self.rank, self.world, self.row_ptrs.device)
        else:
            N = self.n_tracks
            inter = torch.zeros((N, N), dtype=torch.int64, device=self.row_ptrs.device)
            main = torch.cuda.current_stream(self.row_ptrs.device)
            self.side.wait_stream(main)
            freed = [None, None]              # event: the K2 launch that last read scratch[k] has finished
            for c, (a, b) in enumerate(self.chunks):
                k = c & 1
                view = self.scratch[k].view(-1)[: N * (b - a)].view(N, b - a)
                with torch.cuda.stream(self.side):
                    if freed[k] is not None:
                        self.side.wait_event(freed[k])
                    P.pull_rows(self.row_ptrs, a, b - a, view)
                    pulled = torch.cuda.Event()
                    pulled.record(self.side)
                main.wait_event(pulled)
                P.pairwise_inter_accumulate(view, inter)
                freed[k] = torch.cuda.Event()
                freed[k].record(main)
        dist.all_reduce(inter, op=dist.ReduceOp.SUM, group=self.group)
        self.handle.barrier()                 # nobody overwrites its planes while a peer may still be reading them
        return inter
