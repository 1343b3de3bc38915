"""TEST INFRASTRUCTURE ONLY — CPU restatement of the two script-level greedy track filters.

**Parity unpinned**: `generate_tokens_grid.py` / `generate_tokens_gdino.py` execute at module level
(argparse, SAM2 checkpoint, dataset paths) and cannot be imported, so these loops are restated from the
script text; SAM2 propagation is replaced by `track_fn`, a lookup into a synthetic masklet table.  The IoU
inside the loop is computed by `maskpath_oracle.compute_mask_iou` + `reshape_masklet` + nearest resize,
i.e. by the pinned part of the oracle (or by the imported reference functions when `iou_impl` is given).

Status codes (generate_tokens_grid.py:136): 0 not tracked, 1 tracked, 2 filtered, 3 not used.
"""
from __future__ import annotations

import numpy as np
import torch

from . import maskpath_oracle as O


class _Impl:
    """The three reference callables the loops use; swap in `ref_shim` to run on the real reference."""
    compute_mask_iou = staticmethod(O.compute_mask_iou)
    reshape_masklet = staticmethod(O.reshape_masklet)


def _suppress(member_id, resized_track, prompts, miou_thresh, impl, log):
    """generate_tokens_grid.py:266-278 == generate_tokens_gdino.py:288-300."""
    n = 0
    for cand in prompts:
        if cand["status"] > 0:
            continue
        pred_mask = resized_track[cand["frame_idx"]]
        h, w = pred_mask.shape
        pm = torch.from_numpy(cand["segmentation"]).float()
        pm = torch.nn.functional.interpolate(pm[None, None], size=(h, w), mode="nearest")[0, 0]
        iou = impl.compute_mask_iou(pred_mask, pm)
        if log is not None:
            log.append((member_id, cand["prompt_id"], iou))
        if iou > miou_thresh:
            cand["status"] = 2
            cand["filtered_by"] = member_id
            cand["filtered_iou"] = iou
            n += 1
    return n


def grid_greedy(prompts, n_frames, track_fn, *, bin_size=4, n_max_tracks=64, batch_size=4, miou_thresh=0.7,
                impl=_Impl, log=None):
    """generate_tokens_grid.py:133-139 (status init), :148-195 (batching), :266-278 (suppression), :287-292.

    prompts: list of dicts {prompt_id, frame_idx, segmentation (H,W) uint8}, in file order (area-descending).
    track_fn(frame_idx, [prompt dicts]) -> {prompt_id: (T,H,W) fp32 {0,1} masklet}.
    Returns dict(status=..., tracked, filtered, not_used, not_tracked, batches, filtered_by, filtered_iou)."""
    for p in prompts:
        p["status"] = 3 if p["frame_idx"] % bin_size != 0 else 0
        p.pop("filtered_by", None), p.pop("filtered_iou", None)
    n_tracked = n_filtered = 0
    batches = []
    while True:
        if n_tracked >= n_max_tracks:                      # :149 (the idx half of the test is dead code)
            break
        frame, batch = None, []
        for p in prompts:                                  # :165-186 — scans the whole list
            if p["status"] > 0:
                continue
            if frame is None:
                frame = p["frame_idx"]
            elif p["frame_idx"] != frame:
                continue
            batch.append(p)
            p["status"] = 1
            cap = 2 if n_frames > 200 else batch_size
            if len(batch) >= cap or n_tracked + len(batch) >= n_max_tracks:
                break
        if frame is None:
            break
        n_tracked += len(batch)                            # :194
        batches.append([p["prompt_id"] for p in batch])
        masklets = track_fn(frame, batch)
        resized = {pid: impl.reshape_masklet(m) for pid, m in masklets.items()}     # :248-250
        for p in batch:                                    # :252, batch order
            n_filtered += _suppress(p["prompt_id"], resized[p["prompt_id"]], prompts, miou_thresh, impl, log)
    return _result(prompts, batches, n_tracked, n_filtered, n_max_tracks, check_complete=True)


def gdino_greedy(prompts, expression_id, n_frames, track_fn, *, bin_size=4, stability_score_thresh=0.85,
                 n_max_tracks=16, batch_size=4, miou_thresh=0.7, impl=_Impl, log=None):
    """generate_tokens_gdino.py:155-166 (candidate filter), :169-206 (batching), :288-300, :311-313.

    prompts: all prompts of the video; only those with p['expression_id'] == expression_id take part."""
    cands = []
    n_not_used = 0
    for p in prompts:
        if p["expression_id"] != expression_id:
            continue
        p["status"] = 0
        p.pop("filtered_by", None), p.pop("filtered_iou", None)
        if p["frame_idx"] % bin_size != 0 or p["stability_score"] < stability_score_thresh:   # :162
            p["status"] = 3
            n_not_used += 1
        else:
            cands.append(p)
    n_tracked = n_filtered = 0
    batches = []
    while True:
        if n_tracked >= n_max_tracks:                      # :170
            break
        frame, batch = None, []
        for p in cands:
            if p["status"] > 0:
                continue
            if frame is None:
                frame = p["frame_idx"]
            elif p["frame_idx"] != frame:
                break                                      # :194-196 — first other-frame candidate ends the batch
            batch.append(p)
            p["status"] = 1
            n_tracked += 1                                 # :187,193 — counted at append time
            if n_frames > 200 and len(batch) >= 2:
                break
            if len(batch) >= batch_size:
                break
            if len(batch) + n_tracked >= n_max_tracks:     # :201 — n_tracked already includes the batch
                break
        if frame is None:
            break
        batches.append([p["prompt_id"] for p in batch])
        masklets = track_fn(frame, batch)
        resized = {pid: impl.reshape_masklet(m) for pid, m in masklets.items()}
        for p in batch:
            n_filtered += _suppress(p["prompt_id"], resized[p["prompt_id"]], cands, miou_thresh, impl, log)
    res = _result(cands, batches, n_tracked, n_filtered, n_max_tracks, check_complete=False)
    res["n_not_used"] = n_not_used
    res["not_used"] = []          # :311 lists status==3 among *candidates*, which is always empty
    return res


def _result(prompts, batches, n_tracked, n_filtered, n_max_tracks, check_complete):
    by = lambda s: [p["prompt_id"] for p in prompts if p["status"] == s]
    tracked, filtered, not_used, not_tracked = by(1), by(2), by(3), by(0)
    if check_complete and len(tracked) < n_max_tracks:     # generate_tokens_grid.py:291-292
        assert not not_tracked, f"untracked prompts left: {not_tracked}"
    return {
        "status": {p["prompt_id"]: p["status"] for p in prompts},
        "tracked": tracked, "filtered": filtered, "not_used": not_used, "not_tracked": not_tracked,
        "batches": batches, "n_tracked": n_tracked, "n_filtered": n_filtered,
        "filtered_by": {p["prompt_id"]: p["filtered_by"] for p in prompts if p["status"] == 2},
        "filtered_iou": {p["prompt_id"]: p["filtered_iou"] for p in prompts if p["status"] == 2},
    }


def dedup_matrix_greedy(iou: np.ndarray, miou_thresh: float = 0.7):
    """Spatio-temporal variant used by BASELINE configs 2/5 (no reference caller; semantics of
    seg_utils.compute_masklet_iou:110-125 + the same strict `>` suppression in id order):
    visit tracks in index order; a track still alive suppresses every later alive track j with
    iou[i, j] > thresh.  Returns (kept ids, {suppressed id: suppressor id})."""
    n = iou.shape[0]
    alive = np.ones(n, dtype=bool)
    by = {}
    for i in range(n):
        if not alive[i]:
            continue
        for j in range(i + 1, n):
            if alive[j] and iou[i, j] > miou_thresh:
                alive[j] = False
                by[j] = i
    return [i for i in range(n) if alive[i]], by


def part_suppression(masks: torch.Tensor, part_thresh: float = 0.7, compute_P=O.compute_P, autocast_bf16: bool = False):
    """generate_prompts_grid.py:105-116 (script loop, parity unpinned): masks (N,H,W) fp32 sorted by area descending.
    `autocast_bf16` emulates the bf16 rounding of the GEMV output that the reference's CUDA autocast context applies."""
    n = masks.shape[0]
    is_part = torch.tensor([False] * n)
    for k in range(n - 1):
        if is_part[k]:
            continue
        if autocast_bf16:
            flat = masks.reshape(n, -1)
            inter = (flat @ masks[k].reshape(-1, 1)).to(torch.bfloat16).float()
            Pk = (inter / flat.sum(dim=1, keepdim=True)).squeeze(1)
        else:
            Pk = compute_P(masks, masks[k])
        is_part[Pk > part_thresh] = True
        is_part[k] = False
    return is_part
