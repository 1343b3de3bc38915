"""Drop-in mirror of the mask metrics of `track_generation/utils.py`."""
from __future__ import annotations

import numpy as np
import torch

from . import packed as P


@torch.no_grad()
def compute_mask_iou_torch(maskA, maskB) -> float:
    """utils.py:65-75 — inter / (|A| + |B| - inter) with NO empty guard: two empty masks raise
    ZeroDivisionError exactly like the reference's Python-float division."""
    inter, a, b = P.frame_counts(maskA, maskB)[:, 0].tolist()
    return float(inter) / float(a + b - inter)


def mask_metrics_from_counts(inter: np.ndarray, n_pred: np.ndarray, n_gt: np.ndarray):
    """Per-frame precision / recall / IoU from integer counts with the four empty-case rules of utils.py:157-168.
    Values are computed in float64 and stored as fp32, as the reference does by assigning into `torch.zeros(T)`."""
    inter = np.asarray(inter, dtype=np.int64)
    n_pred = np.asarray(n_pred, dtype=np.int64)
    n_gt = np.asarray(n_gt, dtype=np.int64)
    union = n_pred + n_gt - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(union == 0, 1.0, inter / union)
        prec = np.where(n_pred == 0, 1.0, np.where(n_gt == 0, 0.0, inter / n_pred))
        rec = np.where(n_pred == 0, np.where(n_gt == 0, 1.0, 0.0), np.where(n_gt == 0, 1.0, inter / n_gt))
    as_f32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).float()
    return as_f32(prec), as_f32(rec), as_f32(iou)


@torch.no_grad()
def compute_mask_metrics(pred_masks, gt_masks, reduction: str = "mean"):
    """utils.py:132-174 — (precision, recall, iou) over T frames: fp32 CPU tensors, 0-d for 'mean', (T,) for 'none'.
    One kernel + one D2H for all frames (the reference: 4 .item() syncs per frame)."""
    if reduction not in ("mean", "none"):
        raise ValueError(f"Invalid reduction method: {reduction}")
    c = P.frame_counts(pred_masks, gt_masks).cpu().numpy()
    prec, rec, iou = mask_metrics_from_counts(c[0], c[1], c[2])
    if reduction == "mean":
        return prec.mean(), rec.mean(), iou.mean()
    return prec, rec, iou


@torch.no_grad()
def compute_mask_metrics_batch(pred: P.PackedMasks, gt: P.PackedMasks):
    """Label generation for every (track, GT object) pair in one launch (generate_tokens_grid.py:253-264):
    pred (N, T, ...), gt (G, T, ...) packed -> three fp32 CPU tensors (N, G) of frame-mean precision / recall / IoU."""
    inter, area_p, area_g = P.frame_counts_packed(pred, gt)
    inter, area_p, area_g = inter.cpu().numpy(), area_p.cpu().numpy(), area_g.cpu().numpy()
    N, G, T = inter.shape
    out = [torch.zeros(N, G), torch.zeros(N, G), torch.zeros(N, G)]
    for i in range(N):
        for g in range(G):
            p, r, u = mask_metrics_from_counts(inter[i, g], area_p[i], area_g[g])
            out[0][i, g], out[1][i, g], out[2][i, g] = p.mean(), r.mean(), u.mean()
    return tuple(out)


def _bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even of non-negative integers to bfloat16 (8 significant bits), returned as float32 — what a
    bf16-autocast matmul stores for an exactly-integer fp32 accumulator."""
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


@torch.no_grad()
def compute_P(part_masks, full_mask, autocast_bf16: bool = False) -> torch.Tensor:
    """utils.py:178-192 — part-ness P[k] = |part_k ∩ full| / |part_k| as fp32 (N,) on the device (nan for an empty part).
    The reference evaluates this as a GEMV; here it is AND-popcount on packed planes (exact integers).
    `autocast_bf16=True` reproduces what the reference gets when it runs under `torch.autocast('cuda', bfloat16)`
    (generate_prompts_grid.py:59): the matmul output — the intersection — is rounded to bf16 before the division."""
    parts = P.pack_masks(part_masks)
    full = P.pack_masks(full_mask)
    N = int(parts.words.shape[0])
    inter, area_p, _ = P.frame_counts_packed(parts.reshape_lead(N, 1), full.reshape_lead(1, 1))
    inter = inter.view(N).cpu().numpy()
    area = area_p.view(N).cpu().numpy().astype(np.float32)
    num = _bf16_round(inter) if autocast_bf16 else inter.astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / area
    return torch.from_numpy(out.astype(np.float32)).to(parts.device)


@torch.no_grad()
def suppress_part_masks(masks, part_thresh: float = 0.7, autocast_bf16: bool = True) -> np.ndarray:
    """generate_prompts_grid.py:105-116 — masks (N, H, W) of ONE frame, already sorted by area descending.
    Walks the masks in order; every not-yet-part mask k marks as part each mask whose P against k exceeds `part_thresh`
    (strict, fp32), then un-marks itself.  All N x N intersections come from ONE K2 launch instead of N GEMVs.
    Returns the boolean is_part array (the caller keeps `~is_part`)."""
    packed = P.pack_masks(masks)
    N = int(packed.words.shape[0])
    inter = P.pairwise_inter_matrix(packed.reshape_lead(N, 1)).cpu().numpy()           # inter[k, full], diag = areas
    area = np.diag(inter).astype(np.float32)
    num = _bf16_round(inter) if autocast_bf16 else inter.astype(np.float32)
    thr = np.float32(part_thresh)
    is_part = np.zeros(N, dtype=bool)
    for k in range(N - 1):
        if is_part[k]:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            Pk = num[:, k] / area                                                       # P of every mask against full = mask k
        is_part[Pk > thr] = True
        is_part[k] = False
    return is_part
