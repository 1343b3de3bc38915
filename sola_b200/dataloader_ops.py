"""Device versions of the two OR-merges of `dataloader.py` that feed J&F (AlignDataset.get_sam2_masklet :305-351,
get_gt_masklet :278-303), with the file I/O and RLE decoding left to the caller."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import packed as P


def merge_selected_tracks(masklets: Optional[P.PackedMasks], preds: Sequence[float]) -> Optional[P.PackedMasks]:
    """masklets (K, T, H, Wp) in directory order, preds[k] > 0 selects.  No tracks at all -> None (the reference
    returns None when the directories are empty); nothing selected -> all-zero planes of the tracks' shape."""
    if masklets is None or masklets.words.shape[0] == 0:
        return None
    preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
    return P.or_merge(masklets, select=(preds[: masklets.words.shape[0]] > 0))


def merge_gt_objects(masklets: P.PackedMasks) -> P.PackedMasks:
    """OR over the expression's GT objects."""
    return P.or_merge(masklets, select=None)


def get_sam2_masklet_packed(rle_masklets: Sequence, preds: Sequence[float], device=None) -> Optional[P.PackedMasks]:
    """Device version of AlignDataset.get_sam2_masklet (dataloader.py:305-351) once the per-track JSON files are read:
    `rle_masklets[k]` is track k's `sam2_masklet_info['rle']` list in directory order.  RLE decode, selection and OR-merge all
    happen on packed bits; the result feeds evaluator / JFSweep directly (no uint8 (T,H,W) arrays, no fp32 H2D copy)."""
    from . import rle
    if len(rle_masklets) == 0:
        return None
    preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
    selected = [k for k in range(len(rle_masklets)) if preds[k] > 0]
    if not selected:
        first = rle.decode_rle_masklet_packed(rle_masklets[0], device)          # shape donor: zeros of the first track's shape (:346-349)
        return P.PackedMasks(torch.zeros_like(first.words), first.H, first.W)
    planes = [rle.decode_rle_masklet_packed(rle_masklets[k], device) for k in selected]
    stacked = P.PackedMasks(torch.stack([p.words for p in planes]), planes[0].H, planes[0].W)
    return P.or_merge(stacked)


def png_planes(masklet: P.PackedMasks) -> np.ndarray:
    """(T, H, W) uint8 planes with foreground 255, as inference.py:89-91 writes them (`(mask * 255).astype(np.uint8)`)."""
    return P.unpack_masks(masklet, torch.uint8, one_value=255).cpu().numpy()


def save_packed_masklet(path: str, masklet: P.PackedMasks) -> None:
    """Packed sidecar format: the uint32 planes + shape, 32x smaller than uint8 masks and loadable
    straight into the kernels without an RLE decode."""
    np.savez_compressed(path, words=masklet.numpy_u32(), H=masklet.H, W=masklet.W)


def load_packed_masklet(path: str, device=None) -> P.PackedMasks:
    z = np.load(path)
    return P.PackedMasks.from_numpy_u32(z["words"], int(z["H"]), int(z["W"]), device)
