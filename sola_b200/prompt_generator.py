"""Drop-in mirror of `PromptGenerator.get_stability_score` (track_generation/prompt_generator.py:169-186)."""
from __future__ import annotations

import numpy as np

from . import packed as P


def get_stability_score(logit, mask_threshold: float = 0.0, threshold_offset: float = 1.0):
    """count(logit > thr+off) / count(logit > thr-off) as float64 (`nan` with numpy's RuntimeWarning when the
    denominator is 0).  `logit`: numpy / tensor (..., H, W) fp32; a 2-D input returns a numpy float64 scalar.
    Both counts come from one pass of the K1 kernel (the reference makes two numpy passes).
    Not reproduced: the reference's int16 row accumulator wraps for rows with more than 32767 hits (W > 32767)."""
    _, counts = P.binarize_pack_stability(logit, mask_threshold, threshold_offset, want_packed=False)
    c = counts.cpu().numpy()
    hi, lo = c[0], c[2]
    if hi.ndim == 0:
        hi, lo = np.int32(hi), np.int32(lo)
    return hi / lo


class PromptGenerator:
    """Carrier for the method form `self.get_stability_score(logit, ...)`; GroundingDINO / SAM2 inference is out of
    scope and stays in the reference."""

    def get_stability_score(self, logit, mask_threshold: float = 0.0, threshold_offset: float = 1.0):
        return get_stability_score(logit, mask_threshold, threshold_offset)


def stability_filter_keep(frame_idx: int, stability_score: float, bin_size: int, stability_score_thresh: float) -> bool:
    """generate_tokens_gdino.py:162 — keep unless off-bin or score < thresh (== thresh and NaN are kept)."""
    return not (frame_idx % bin_size != 0 or stability_score < stability_score_thresh)
