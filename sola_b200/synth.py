"""Seeded synthetic inputs of the BASELINE.json shapes.  No dataset or checkpoint is available
offline, so benches and tests run on these; every generator works on CPU (tests, golden vectors) and on CUDA
(full-size bench inputs are generated in HBM chunk by chunk)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def smooth_logits(n: int, H: int, W: int, seed: int, device="cpu", *, cell: int = 16, bias: float = 0.3,
                  gain: float = 6.0, noise: float = 0.5, dtype=torch.float32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(n, H, W) logits: a coarse Gaussian field upsampled bicubically so the `> 0` masks look like blobs and the
    +-1 bands (stability score) are non-trivial, plus white noise."""
    g = _gen(seed, device)
    z = torch.randn((n, 1, H // cell + 2, W // cell + 2), generator=g, device=device)
    up = torch.nn.functional.interpolate(z, size=(H, W), mode="bicubic", align_corners=False)[:, 0]
    up = gain * (up - bias) + noise * torch.randn((n, H, W), generator=g, device=device)
    if out is not None:
        out.copy_(up)
        return out
    return up.to(dtype)


def adversarial_logit_planes(H: int, W: int, device="cpu") -> torch.Tensor:
    """Edge-case planes for K1: exact thresholds, signed zeros, NaN / inf, all-negative (-> nan stability score)."""
    vals = torch.tensor([0.0, -0.0, 1.0, -1.0, float("nan"), float("inf"), -float("inf"),
                         1.0000001, 0.99999994, -0.99999994, -1.0000001, 1e-45, -1e-45, 3.0, -3.0, 0.5], device=device)
    idx = torch.arange(H * W, device=device) % vals.numel()
    p0 = vals[idx].reshape(H, W)
    p1 = torch.full((H, W), -5.0, device=device)                 # nothing above any threshold
    p2 = torch.full((H, W), 5.0, device=device)                  # everything above every threshold
    p3 = vals[(idx * 7 + 3) % vals.numel()].reshape(H, W)
    return torch.stack([p0, p1, p2, p3])


def blob_masklet(T: int, H: int, W: int, seed: int, device="cpu", *, drift: float = 2.0, fill: float = 0.3) -> torch.Tensor:
    """(T, H, W) uint8 masklet: one smooth field thresholded, translated a little every frame."""
    g = _gen(seed, device)
    cell = 24
    z = torch.randn((1, 1, H // cell + 4, W // cell + 4), generator=g, device=device)
    big = torch.nn.functional.interpolate(z, size=(H + 64, W + 64), mode="bicubic", align_corners=False)[0, 0]
    thr = torch.quantile(big.flatten()[:: max(1, big.numel() // 65536)].float(), 1.0 - fill)
    walk = torch.cumsum(torch.randn((T, 2), generator=g, device=device) * drift, dim=0)
    walk = (walk - walk.mean(0)).clamp(-31, 31).round().long() + 32
    frames = [big[walk[t, 0]: walk[t, 0] + H, walk[t, 1]: walk[t, 1] + W] > thr for t in range(T)]
    return torch.stack(frames).to(torch.uint8)


def jf_pair(T: int, H: int, W: int, seed: int, device="cpu", *, flip: float = 0.02, empty_frames: int = 2):
    """(pred, gt) uint8 masklets for J&F: pred = gt XOR sparse noise; a few frames empty in both (union == 0 rule)."""
    gt = blob_masklet(T, H, W, seed, device)
    g = _gen(seed + 7919, device)
    noise = torch.rand((T, H, W), generator=g, device=device) < flip
    pred = gt ^ noise.to(torch.uint8)
    for t in range(min(empty_frames, T)):
        k = (t * 7 + 3) % T
        gt[k] = 0
        pred[k] = 0
    return pred, gt


def dedup_candidates(n: int, T: int, H: int, W: int, seed: int, device="cpu", *, bin_size: int = 4,
                     n_clusters: Optional[int] = None, jitter: int = 2) -> Tuple[torch.Tensor, List[dict]]:
    """Candidate table for the greedy filters: n candidates in ~n/3 clusters of near-duplicates.
    Returns (logits (n, T, H, W) fp32 — what SAM2 would emit when tracking each candidate —, prompts list).
    Each cluster is an object-like mask (a few large smooth blobs, ~5-20 % of the frame); members of a cluster are the
    same object shifted by up to `jitter` px with their own level / edge sharpness, so IoUs straddle 0.7 and the
    stability scores spread over ~0.8-0.99.  prompt k: frame_idx on the `bin_size` grid, segmentation = its own
    masklet at that frame (uint8), area-sorted descending with prompt_id = rank (generate_prompts_grid.py:131-133)."""
    g = _gen(seed, device)
    K = n_clusters or max(1, n // 3)
    cell = max(16, min(H, W) // 5)
    base = torch.randn((K, 1, H // cell + 4, W // cell + 4), generator=g, device=device)
    base = torch.nn.functional.interpolate(base, size=(H + 16, W + 16), mode="bicubic", align_corners=False)[:, 0]
    cluster = torch.randint(0, K, (n,), generator=g, device=device)
    shift = torch.randint(-jitter, jitter + 1, (n, 2), generator=g, device=device) + 8
    level = 0.9 + 0.5 * torch.rand((n,), generator=g, device=device)
    sharp = 15.0 + 65.0 * torch.rand((n,), generator=g, device=device)       # logit units per field unit: soft ... crisp edges
    logits = torch.empty((n, T, H, W), dtype=torch.float32, device=device)
    drift = torch.linspace(0, 0.15, T, device=device).view(T, 1, 1)
    for i in range(n):
        f = base[cluster[i], shift[i, 0]: shift[i, 0] + H, shift[i, 1]: shift[i, 1] + W]
        logits[i] = sharp[i] * (f[None] - level[i] - drift) + 0.5 * torch.randn((T, H, W), generator=g, device=device)
    n_bins = max(1, (T + bin_size - 1) // bin_size)
    frame_idx = (torch.randint(0, n_bins, (n,), generator=g, device=device) * bin_size).clamp(max=T - 1).cpu().numpy()
    segs = [(logits[i, int(frame_idx[i])] > 0).to(torch.uint8).cpu().numpy() for i in range(n)]
    areas = np.array([int(s.sum()) for s in segs])
    order = np.argsort(-areas, kind="stable")
    prompts = []
    for rank, i in enumerate(order.tolist()):
        prompts.append({"prompt_id": rank, "source": i, "frame_idx": int(frame_idx[i]), "segmentation": segs[i],
                        "area": int(areas[i])})
    return logits[torch.as_tensor(order, device=device)], prompts


MEVIS_SHAPES = [(360, 640), (480, 854), (720, 1280), (1080, 1920)]        # MeViS valid_u mixes 360p ... 1080p (recalled)


def object_pair(T: int, H: int, W: int, seed: int, device="cpu", *, shift: int = 3, level: float = 0.05, speckle: float = 0.0,
                empty_frames: int = 1):
    """(pred, gt) uint8 masklets whose errors look like a tracker's, not like noise: the prediction is the same smooth object field
    cut at a slightly different level and displaced by up to `shift` px (IoU ~0.8-0.95, contours a few pixels apart), optionally
    with a little speckle; a few frames are empty in both (union == 0 rule)."""
    g = _gen(seed, device)
    cell = max(24, min(H, W) // 6)
    z = torch.randn((1, 1, H // cell + 4, W // cell + 4), generator=g, device=device)
    big = torch.nn.functional.interpolate(z, size=(H + 64, W + 64), mode="bicubic", align_corners=False)[0, 0]
    thr = torch.quantile(big.flatten()[:: max(1, big.numel() // 65536)].float(), 0.75)
    walk = torch.cumsum(torch.randn((T, 2), generator=g, device=device) * 2.0, dim=0)
    walk = (walk - walk.mean(0)).clamp(-24, 24).round().long() + 32
    d = torch.randint(-shift, shift + 1, (T, 2), generator=g, device=device)
    gt = torch.stack([big[walk[t, 0]: walk[t, 0] + H, walk[t, 1]: walk[t, 1] + W] > thr for t in range(T)])
    pred = torch.stack([big[walk[t, 0] + d[t, 0]: walk[t, 0] + d[t, 0] + H, walk[t, 1] + d[t, 1]: walk[t, 1] + d[t, 1] + W] > thr + level
                        for t in range(T)])
    if speckle > 0:
        pred = pred ^ (torch.rand((T, H, W), generator=g, device=device) < speckle)
    gt, pred = gt.to(torch.uint8), pred.to(torch.uint8)
    for t in range(min(empty_frames, max(T - 1, 0))):
        k = (t * 7 + 3) % T
        gt[k] = 0
        pred[k] = 0
    return pred, gt


def mevis_like_sweep(n_videos: int, exprs_per_video: int, seed: int, device="cpu", *, t_range=(30, 200), shapes=None, pack=None):
    """BASELINE config 4 shape: a J&F evaluation sweep over `n_videos` multi-object videos of mixed resolution and length, each with
    `exprs_per_video` (video, expression) units.  Returns [(video_id, expression_id, pred, gt)] with uint8 (T, H, W) masklets, or —
    when `pack` (a callable such as sola_b200.pack_masks) is given — bit-packed planes, built unit by unit so the full-size sweep never
    holds more than one unit unpacked."""
    rng = np.random.default_rng(seed)
    shapes = shapes or MEVIS_SHAPES
    units = []
    for v in range(n_videos):
        H, W = shapes[int(rng.integers(0, len(shapes)))]
        T = int(rng.integers(t_range[0], t_range[1] + 1))
        for e in range(exprs_per_video):
            pred, gt = object_pair(T, H, W, seed * 7919 + v * 131 + e, device, shift=int(rng.integers(0, 5)),
                                   level=float(rng.uniform(-0.08, 0.08)), empty_frames=int(rng.integers(0, 3)))
            if rng.random() < 0.05:
                pred = torch.zeros_like(pred)                      # nothing selected: zero planes (dataloader.py:346-349) -> tp == 0 rule
            if pack is not None:
                pred, gt = pack(pred), pack(gt)
            units.append((f"v{v:03d}", f"{e}", pred, gt))
    return units
