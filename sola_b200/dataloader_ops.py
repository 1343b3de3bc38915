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
