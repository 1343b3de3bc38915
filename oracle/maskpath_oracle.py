"""TEST INFRASTRUCTURE ONLY — CPU restatement of SOLA's masklet-scoring path (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this file; the product (`sola_b200/`) never does and fails loudly without its CUDA library.

Every function restates ONE reference function (cited `file:line`, relative to the reference tree) with
the reference's own arithmetic — same dtype of each sum, same empty-case rules, same comparison
strictness — so that it can double as the timed CPU baseline.  Where the reference's fp32 arithmetic is
inexact (whole-volume sums above 2**24) an `*_exact` twin gives the integer truth the GPU path produces.

Pinning status: the reference ships no tests / golden vectors for this path.  The
pins are `tests/golden/*.npz`, produced by `oracle/gen_golden.py` by importing the unmodified reference
functions (`oracle/ref_shim.py`) in the builder container; `tests/test_oracle_golden.py` checks this file
against them.  Two pieces have no importable reference and stay **parity unpinned**: the script-level
greedy loops (restated in `oracle/greedy_oracle.py` from the script text) and the boundary-F extension
(`oracle/boundary_oracle.py`, external DAVIS definition).
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# B1  binarise                                     generate_tokens_grid.py:215,219,222 (gdino :232,237,240)
# --------------------------------------------------------------------------------------------------

def binarize(logits: torch.Tensor) -> torch.Tensor:
    """Strict `> 0.0` on fp32 logits, result as fp32 {0,1}.  0.0, -0.0 and NaN map to 0."""
    return (logits > 0.0).float()


def pack_bits(mask) -> np.ndarray:
    """Row-padded little-endian bit packing used by the GPU layout: (..., H, W) {0,1} -> (..., H, Wp) uint32,
    bit b of word w of a row = pixel 32*w + b, pad bits zero."""
    m = np.asarray(mask)
    m = (m != 0)
    W = m.shape[-1]
    Wp = (W + 31) // 32
    pad = Wp * 32 - W
    if pad:
        m = np.concatenate([m, np.zeros(m.shape[:-1] + (pad,), dtype=bool)], axis=-1)
    by = np.packbits(m, axis=-1, bitorder="little")            # (..., Wp*4) uint8
    return np.ascontiguousarray(by).view("<u4").reshape(m.shape[:-1] + (Wp,))


def unpack_bits(packed: np.ndarray, W: int) -> np.ndarray:
    p = np.ascontiguousarray(packed.astype("<u4"))
    by = p.view(np.uint8).reshape(p.shape[:-1] + (p.shape[-1] * 4,))
    return np.unpackbits(by, axis=-1, bitorder="little")[..., :W]


# --------------------------------------------------------------------------------------------------
# S1  stability score                              track_generation/prompt_generator.py:169-186
# --------------------------------------------------------------------------------------------------

def stability_counts(logit: np.ndarray, mask_threshold: float = 0.0, threshold_offset: float = 1.0):
    """(count(logit > thr+off), count(logit > thr-off)) with the reference's accumulator types:
    per-row totals in int16 (wraps past 32767 hits in one row), then int32 over rows."""
    logit = np.asarray(logit)
    hi = (logit > (mask_threshold + threshold_offset)).sum(-1, dtype=np.int16).sum(-1, dtype=np.int32)
    lo = (logit > (mask_threshold - threshold_offset)).sum(-1, dtype=np.int16).sum(-1, dtype=np.int32)
    return hi, lo


def get_stability_score(logit: np.ndarray, mask_threshold: float = 0.0, threshold_offset: float = 1.0):
    """hi / lo as float64; 0/0 -> nan (numpy RuntimeWarning, as in the reference)."""
    hi, lo = stability_counts(logit, mask_threshold, threshold_offset)
    return hi / lo


def stability_filter_keep(frame_idx: int, stability_score: float, bin_size: int, thresh: float) -> bool:
    """S2 — generate_tokens_gdino.py:162.  A prompt is *dropped* iff off-bin or score < thresh;
    `== thresh` and NaN both pass (NaN < x is False)."""
    return not (frame_idx % bin_size != 0 or stability_score < thresh)


# --------------------------------------------------------------------------------------------------
# I1 / I2 / I3  mask IoU                           track_generation/seg_utils.py:110-142, utils.py:65-75
# --------------------------------------------------------------------------------------------------

def _inter_and_total_f32(a: torch.Tensor, b: torch.Tensor):
    """The two fp32 reductions every IoU in the reference is built from, read back as Python floats."""
    inter = torch.sum(a * b).item()
    total = torch.sum(a + b).item()
    return inter, total


def compute_mask_iou(maskA: torch.Tensor, maskB: torch.Tensor) -> float:
    """seg_utils.py:129-142 — single (H,W) frame; union==0 -> 1.0."""
    inter, total = _inter_and_total_f32(maskA, maskB)
    union = total - inter
    return 1.0 if union == 0.0 else inter / union


def compute_masklet_iou(maskletA: torch.Tensor, maskletB: torch.Tensor, device="cpu") -> float:
    """seg_utils.py:110-125 — same formula over the whole (T,H,W) volume (fp32 sums: inexact > 2**24)."""
    inter, total = _inter_and_total_f32(maskletA.to(device), maskletB.to(device))
    union = total - inter
    return 1.0 if union == 0 else inter / union


def compute_mask_iou_torch(maskA: torch.Tensor, maskB: torch.Tensor) -> float:
    """utils.py:65-75 — no empty guard: both-empty raises ZeroDivisionError."""
    inter = (maskA * maskB).sum().item()
    return inter / (maskA.sum().item() + maskB.sum().item() - inter)


def iou_counts_exact(a, b):
    """Integer truth: (|A∩B|, |A|, |B|) as Python ints."""
    a = np.asarray(a) != 0
    b = np.asarray(b) != 0
    return int(np.count_nonzero(a & b)), int(np.count_nonzero(a)), int(np.count_nonzero(b))


def iou_from_counts(inter: int, area_a: int, area_b: int, empty_value: float = 1.0) -> float:
    union = area_a + area_b - inter
    return empty_value if union == 0 else inter / union


# --------------------------------------------------------------------------------------------------
# R1 / R2  resizes                                 seg_utils.py:145-160 ; generate_tokens_grid.py:271-272
# --------------------------------------------------------------------------------------------------

def default_target_shape(H: int, W: int):
    return (540, 960) if H < W else (960, 540)


def reshape_masklet(masklet: torch.Tensor, target_shape=None) -> torch.Tensor:
    """Bilinear (align_corners=False, no antialias) to 540x960 / 960x540, then `> 0.5`, fp32 {0,1}.
    T rides the channel dimension of a single NCHW image."""
    T, H, W = masklet.shape
    nh, nw = default_target_shape(H, W) if target_shape is None else target_shape
    up = torch.nn.functional.interpolate(masklet[None], size=(nh, nw), mode="bilinear")
    return (up > 0.5)[0].float()


def bilinear_source_index(out_size: int, in_size: int):
    """ATen's align_corners=False source coordinates in fp32 (area_pixel_compute_source_index):
    returns (i0, i1, lambda0, lambda1) arrays of length out_size."""
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = scale * (dst + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = src.astype(np.int32)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def resize_prompt_nearest(seg: np.ndarray, h: int, w: int) -> torch.Tensor:
    """R2 — legacy 'nearest' (src = floor(dst * in/out)) of a uint8 prompt mask to (h, w), fp32."""
    t = torch.from_numpy(np.ascontiguousarray(seg)).float()
    return torch.nn.functional.interpolate(t[None, None], size=(h, w), mode="nearest")[0, 0]


def nearest_source_index(out_size: int, in_size: int) -> np.ndarray:
    """ATen nearest_neighbor_compute_source_index: min(floor(dst * (in/out as fp32)), in-1)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


# --------------------------------------------------------------------------------------------------
# M1  per-frame precision / recall / IoU           track_generation/utils.py:132-174
# --------------------------------------------------------------------------------------------------

def compute_mask_metrics(pred_masks: torch.Tensor, gt_masks: torch.Tensor, reduction: str = "mean"):
    T = pred_masks.shape[0]
    prec, rec, iou = torch.zeros(T), torch.zeros(T), torch.zeros(T)
    for t in range(T):
        p, g = pred_masks[t], gt_masks[t]
        inter, total = (p * g).sum().item(), (p + g).sum().item()
        n_pred, n_gt = p.sum().item(), g.sum().item()
        union = total - inter
        iou[t] = 1.0 if union == 0 else inter / union
        if n_pred == 0:
            prec[t], rec[t] = 1.0, (1.0 if n_gt == 0 else 0.0)
        elif n_gt == 0:
            prec[t], rec[t] = 0.0, 1.0
        else:
            prec[t], rec[t] = inter / n_pred, inter / n_gt
    if reduction == "mean":
        return prec.mean(), rec.mean(), iou.mean()
    if reduction == "none":
        return prec, rec, iou
    raise ValueError(f"Invalid reduction method: {reduction}")


def mask_metrics_from_counts(inter, n_pred, n_gt):
    """Same rules from per-frame integer counts (arrays of length T) -> three fp32 tensors (T,)."""
    inter, n_pred, n_gt = (np.asarray(x, dtype=np.int64) for x in (inter, n_pred, n_gt))
    T = inter.shape[0]
    prec, rec, iou = torch.zeros(T), torch.zeros(T), torch.zeros(T)
    for t in range(T):
        i, p, g = int(inter[t]), int(n_pred[t]), int(n_gt[t])
        u = p + g - i
        iou[t] = 1.0 if u == 0 else i / u
        if p == 0:
            prec[t], rec[t] = 1.0, (1.0 if g == 0 else 0.0)
        elif g == 0:
            prec[t], rec[t] = 0.0, 1.0
        else:
            prec[t], rec[t] = i / p, i / g
    return prec, rec, iou


# --------------------------------------------------------------------------------------------------
# P1  part-ness                                    track_generation/utils.py:178-192
# --------------------------------------------------------------------------------------------------

def compute_P(part_masks: torch.Tensor, full_mask: torch.Tensor) -> torch.Tensor:
    N = part_masks.shape[0]
    flat = part_masks.reshape(N, -1)
    inter = flat @ full_mask.reshape(-1, 1)
    return (inter / flat.sum(dim=1, keepdim=True)).squeeze(1)


# --------------------------------------------------------------------------------------------------
# J1 / F1  region J and volumetric F               evaluator.py:227-247
# --------------------------------------------------------------------------------------------------

def compute_J(pred_masklet: torch.Tensor, gt_masklet: torch.Tensor):
    """Per-frame IoU (1.0 when the union is empty), float64 mean over frames."""
    vals = []
    for p, g in zip(pred_masklet, gt_masklet):
        inter = (p * g).sum().item()
        union = (p + g).sum().item() - inter
        vals.append(1.0 if union == 0 else inter / union)
    return np.mean(vals)


def compute_F(pred_masklet: torch.Tensor, gt_masklet: torch.Tensor) -> float:
    """Volumetric pixel F1: tp/fp/fn summed over all T*H*W in fp32; 0.0 when tp == 0."""
    tp = (pred_masklet * gt_masklet).sum().item()
    fp = ((1 - gt_masklet) * pred_masklet).sum().item()
    fn = (gt_masklet * (1 - pred_masklet)).sum().item()
    if tp == 0:
        return 0.0
    precision, recall = tp / (tp + fp), tp / (tp + fn)
    return 2 * precision * recall / (precision + recall)


def jf_counts_exact(pred, gt):
    """Integer truth per frame: inter[T], n_pred[T], n_gt[T] (int64)."""
    p = np.asarray(pred) != 0
    g = np.asarray(gt) != 0
    ax = (1, 2)
    return ((p & g).sum(ax, dtype=np.int64), p.sum(ax, dtype=np.int64), g.sum(ax, dtype=np.int64))


def J_from_counts(inter, n_pred, n_gt) -> float:
    inter, n_pred, n_gt = (np.asarray(x, dtype=np.int64) for x in (inter, n_pred, n_gt))
    vals = []
    for i, p, g in zip(inter.tolist(), n_pred.tolist(), n_gt.tolist()):
        u = p + g - i
        vals.append(1.0 if u == 0 else i / u)
    return np.mean(vals)


def F_from_counts(inter, n_pred, n_gt) -> float:
    tp = int(np.sum(inter))
    fp = int(np.sum(n_pred)) - tp
    fn = int(np.sum(n_gt)) - tp
    if tp == 0:
        return 0.0
    precision, recall = tp / (tp + fp), tp / (tp + fn)
    return 2 * precision * recall / (precision + recall)


# --------------------------------------------------------------------------------------------------
# O1 / O2  OR-merge of selected tracks / GT objects      dataloader.py:278-351
# --------------------------------------------------------------------------------------------------

def merge_selected_tracks(masklets, preds):
    """dataloader.py:319-350 with the file I/O removed.  `masklets` is the list of decoded (T,H,W) uint8
    arrays in directory order, `preds[i] > 0` selects track i.  Quirks kept: the first listed track fixes
    the output shape even when unselected (zeros of its shape); no tracks at all -> None."""
    merged = None
    for m, p in zip(masklets, preds):
        if p < 1 and merged is not None:
            continue
        if p > 0:
            merged = m if merged is None else np.logical_or(merged, m)
        elif merged is None:
            merged = np.zeros(m.shape, dtype=np.uint8)
    return merged


def merge_gt_objects(masklets):
    """dataloader.py:285-299: OR over the expression's GT objects (first one passes through untouched)."""
    merged = None
    for m in masklets:
        merged = m if merged is None else np.logical_or(merged, m)
    return merged


# --------------------------------------------------------------------------------------------------
# E1  J&F sweep                                    evaluator.py:174-225 (I/O removed)
# --------------------------------------------------------------------------------------------------

def jf_sweep(units):
    """`units`: iterable of (video_id, expression_id, pred_masklet | None, gt_masklet) with uint8 arrays.
    Returns (per-unit dict, mean_J, mean_F, mean_JF) exactly as compute_JF_metrics accumulates them."""
    out, Js, Fs, JFs = {}, [], [], []
    for vid, eid, pred, gt in units:
        if pred is None:
            J = F = JF = 0.0
        else:
            g = torch.from_numpy(np.ascontiguousarray(gt)).float()
            p = torch.from_numpy(np.ascontiguousarray(pred)).float()
            J, F = float(compute_J(p, g)), float(compute_F(p, g))
            JF = (J + F) / 2
        out.setdefault(vid, {})[eid] = {"J": J, "F": F, "JF": JF}
        Js.append(J), Fs.append(F), JFs.append(JF)
    return out, np.mean(Js), np.mean(Fs), np.mean(JFs)


# --------------------------------------------------------------------------------------------------
# X1  recall helpers (no pixels; listed because the north-star names tools/metric.py)   tools/metric.py:2-59
# --------------------------------------------------------------------------------------------------

def recall_per_track(gt_anno_ids, preds, labels, corresponding_gt_anno_ids):
    out = []
    for gid in gt_anno_ids:
        hit = miss = 0
        for pred, label, cid in zip(preds, labels, corresponding_gt_anno_ids):
            if cid == gid and label == 1:
                if pred > 0:
                    hit += 1
                else:
                    miss += 1
        if hit + miss:
            out.append(hit / (hit + miss))
    return out


def recall_per_exp(gt_anno_ids, preds, labels, corresponding_gt_anno_ids):
    found = 0
    for gid in gt_anno_ids:
        found += any(cid == gid and label == 1 and pred > 0
                     for pred, label, cid in zip(preds, labels, corresponding_gt_anno_ids))
    return found / len(gt_anno_ids)
