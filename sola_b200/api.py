"""The batched entry points named in SURVEY.md §8(b), as thin compositions of packed.py / dedup.py / evaluator.py.
(The per-call drop-ins with the reference signatures live in seg_utils.py, utils.py, prompt_generator.py, evaluator.py.)"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import dedup, evaluator, packed as P

binarize_pack_stability = P.binarize_pack_stability          # logits[N,T,H,W], thr, off -> (PackedMasks, counts[3,N,T])
binarize_pack_resize = P.binarize_pack_resize                # same + bilinear-resized planes, one pass (fused K1+R1)


def pairwise_iou_matrix(packed_tracks: P.PackedMasks):
    """(N, T, H, Wp) -> (float64 IoU (N, N), int64 intersections (N, N), int64 areas (N,)); empty ∪ empty -> 1.0."""
    inter = P.pairwise_inter_matrix(packed_tracks).cpu().numpy()
    return P.iou_matrix_from_inter(inter), inter, np.diag(inter).copy()


def gathered_iou(packed_tracks: P.PackedMasks, frame_idx, packed_prompts: P.PackedMasks) -> np.ndarray:
    """M[i, j] = IoU(track_i[frame_idx[j]], prompt_j) as float64 (N, P) — the matrix the reference greedy walks."""
    c = P.gathered_inter(packed_tracks, packed_prompts, frame_idx).cpu().numpy()
    return dedup.iou_from_counts(c[0], c[1], c[2])


def greedy_filter(M: np.ndarray, prompts: Sequence[dict], n_frames: int, **rules) -> dict:
    """Replay generate_tokens_grid.py / generate_tokens_gdino.py's loop (mode='grid' | 'gdino' in `rules`) on a precomputed
    gathered IoU matrix; returns tracked / filtered / not_used ids, batches, filtered_by, filtered_iou."""
    state = dedup.GreedyState(prompts, n_frames, **rules)
    while (batch := state.next_batch()) is not None:
        state.apply_iou_rows(batch, np.asarray(M)[batch])
    return state.result()


def jf_batch(packed_pred_words: torch.Tensor, packed_gt_words: torch.Tensor, frame_offsets: torch.Tensor,
             unit_frame_ranges: Optional[Sequence] = None):
    """Ragged J&F: flat int32 word buffers of many units + int64 device word offsets per frame -> per-frame int32 counts
    (3, n_frames); with `unit_frame_ranges` [(f0, f1), ...] also the per-unit (J, F) list (evaluator.py:227-247 formulas)."""
    counts = P.frame_counts_ragged(packed_pred_words, packed_gt_words, frame_offsets)
    if unit_frame_ranges is None:
        return counts
    c = counts.cpu().numpy()
    jf = [(float(evaluator.J_from_counts(c[0, a:b], c[1, a:b], c[2, a:b])), float(evaluator.F_from_counts(c[0, a:b], c[1, a:b], c[2, a:b])))
          for a, b in unit_frame_ranges]
    return counts, jf
