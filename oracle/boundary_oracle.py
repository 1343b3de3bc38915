"""TEST INFRASTRUCTURE ONLY — boundary F-measure oracle (numpy).

**Parity unpinned / external spec.**  The reference tree has no boundary measure: its `F` is the
volumetric pixel F1 of `evaluator.py:239-247`.  BASELINE.json's north-star nevertheless
asks for "seg2bmap boundary plus disk-dilation match", so this file restates the public DAVIS-2017 /
MeViS evaluation definition (`davis2017/metrics.py: f_measure, seg2bmap`, recalled — not present in
/root/reference): boundary map by XOR with the east / south / south-east neighbours, dilation by
`skimage.morphology.disk(bound_pix)` with `cv2.dilate` border rules (outside the image contributes
nothing), matches counted per frame.  `tests/test_oracle_golden.py` cross-checks the dilation against
`cv2.dilate` when OpenCV is importable.
"""
from __future__ import annotations

import math

import numpy as np


def bound_pix_for(H: int, W: int, bound_th: float = 0.008) -> int:
    return int(bound_th) if bound_th >= 1 else int(math.ceil(bound_th * math.sqrt(H * H + W * W)))


def seg2bmap(seg: np.ndarray) -> np.ndarray:
    seg = np.asarray(seg) != 0
    e = np.zeros_like(seg)
    s = np.zeros_like(seg)
    se = np.zeros_like(seg)
    e[:, :-1] = seg[:, 1:]
    s[:-1, :] = seg[1:, :]
    se[:-1, :-1] = seg[1:, 1:]
    b = (seg ^ e) | (seg ^ s) | (seg ^ se)
    b[-1, :] = seg[-1, :] ^ e[-1, :]
    b[:, -1] = seg[:, -1] ^ s[:, -1]
    b[-1, -1] = False
    return b


def disk_half_widths(r: int):
    """Half-width of the disk footprint on each row dy in [-r, r]: max dx with dx^2 + dy^2 <= r^2."""
    return [int(math.isqrt(r * r - dy * dy)) for dy in range(-r, r + 1)]


def dilate_disk(b: np.ndarray, r: int) -> np.ndarray:
    b = np.asarray(b, dtype=bool)
    H, W = b.shape
    out = np.zeros_like(b)
    if r == 0:
        return b.copy()
    # horizontal dilation by every distinct half-width, via a running count
    csum = np.concatenate([np.zeros((H, 1), np.int32), np.cumsum(b, axis=1, dtype=np.int32)], axis=1)
    cache = {}
    for dy, hw in zip(range(-r, r + 1), disk_half_widths(r)):
        if hw not in cache:
            x = np.arange(W)
            lo = np.maximum(x - hw, 0)
            hi = np.minimum(x + hw, W - 1) + 1
            cache[hw] = (csum[:, hi] - csum[:, lo]) > 0
        hd = cache[hw]
        # out[y] |= hd[y + dy]  (structuring element is symmetric, sign of dy is irrelevant)
        if dy >= 0:
            out[: H - dy if dy else H] |= hd[dy:]
        else:
            out[-dy:] |= hd[: H + dy]
    return out


def boundary_counts(pred: np.ndarray, gt: np.ndarray, bound_th: float = 0.008):
    """Integer counts for one frame: (n_fg, n_gt, fg_match, gt_match)."""
    H, W = pred.shape
    r = bound_pix_for(H, W, bound_th)
    fg_b, gt_b = seg2bmap(pred), seg2bmap(gt)
    fg_dil, gt_dil = dilate_disk(fg_b, r), dilate_disk(gt_b, r)
    return (int(fg_b.sum()), int(gt_b.sum()), int((fg_b & gt_dil).sum()), int((gt_b & fg_dil).sum()))


def f_from_boundary_counts(n_fg: int, n_gt: int, fg_match: int, gt_match: int) -> float:
    if n_fg == 0 and n_gt > 0:
        precision, recall = 1.0, 0.0
    elif n_fg > 0 and n_gt == 0:
        precision, recall = 0.0, 1.0
    elif n_fg == 0 and n_gt == 0:
        precision, recall = 1.0, 1.0
    else:
        precision, recall = fg_match / float(n_fg), gt_match / float(n_gt)
    return 0.0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)


def boundary_f_frame(pred: np.ndarray, gt: np.ndarray, bound_th: float = 0.008) -> float:
    return f_from_boundary_counts(*boundary_counts(pred, gt, bound_th))


def boundary_f_masklet(pred: np.ndarray, gt: np.ndarray, bound_th: float = 0.008) -> float:
    """Mean over frames (DAVIS / MeViS convention)."""
    return float(np.mean([boundary_f_frame(p, g, bound_th) for p, g in zip(pred, gt)]))
