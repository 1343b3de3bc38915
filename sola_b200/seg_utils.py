"""Drop-in mirror of the mask helpers of `track_generation/seg_utils.py` (same names, arguments, return types,
empty-case rules), computed by the sm_100a kernels.  Inputs may be CUDA tensors (no copy), CPU tensors or numpy
arrays (one H2D copy).  Masks must be {0,1}-valued as everywhere in the reference."""
from __future__ import annotations

import torch

from . import packed as P


@torch.no_grad()
def compute_mask_iou(maskA, maskB) -> float:
    """seg_utils.py:129-142 — IoU of two (H, W) masks; empty union -> 1.0.  One kernel, one D2H of 3 ints
    (the reference: 4 ATen launches + 2 .item() syncs)."""
    inter, a, b = P.frame_counts(maskA, maskB)[:, 0].tolist()
    union = a + b - inter
    if union == 0:
        return 1.0
    return inter / union


@torch.no_grad()
def compute_masklet_iou(maskletA, maskletB, device=None) -> float:
    """seg_utils.py:110-125 — IoU over the whole (N, H, W) volume; empty union -> 1.0.
    Counts are exact integers (the reference's fp32 sums drift above 2**24 set pixels; |delta| < 1e-6)."""
    a = P.to_device(maskletA, device=device)
    c = P.frame_counts(a, P.to_device(maskletB, device=a.device)).cpu().numpy()          # per-frame int32 counts, one read-back
    inter, na, nb = (int(v) for v in c.sum(axis=1, dtype="int64"))
    union = na + nb - inter
    if union == 0:
        return 1.0
    return inter / union


def reshape_masklet(masklet, target_shape=None) -> torch.Tensor:
    """seg_utils.py:145-160 — bilinear resize (align_corners=False) to 540x960 / 960x540 (or `target_shape`),
    `> 0.5`, returned as fp32 {0,1} (N, H', W') on the device.  Reproduces ATen-CUDA's interpolation arithmetic."""
    _, f32 = P.resize_bilinear_bin_f32(masklet, target_shape, want_packed=False, want_f32=True)
    return f32


def reshape_masklet_packed(masklet: P.PackedMasks, target_shape=None) -> P.PackedMasks:
    """Same operation on bit-packed planes (input must be a {0,1} masklet, which it is by construction)."""
    return P.resize_bilinear_bin(masklet, target_shape)
