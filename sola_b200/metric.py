"""Mirror of `tools/metric.py` (recall over per-track scalar predictions).  No pixel work and no caller in the
reference; kept host-side so that `from tools import metric` call sites have a drop-in."""
from __future__ import annotations


def recall_per_track(gt_anno_ids, preds, labels, corresponding_gt_anno_ids):
    """tools/metric.py:2-32 — per GT track, recall of its positively-labelled candidates (tracks without any are skipped)."""
    rows = list(zip(preds, labels, corresponding_gt_anno_ids))
    out = []
    for gid in gt_anno_ids:
        hits = [bool(p > 0) for p, l, c in rows if c == gid and l == 1]
        if hits:
            out.append(sum(hits) / len(hits))
    return out


def recall_per_exp(gt_anno_ids, preds, labels, corresponding_gt_anno_ids):
    """tools/metric.py:34-59 — fraction of GT tracks with at least one selected positive candidate."""
    rows = list(zip(preds, labels, corresponding_gt_anno_ids))
    detected = sum(any(c == gid and l == 1 and p > 0 for p, l, c in rows) for gid in gt_anno_ids)
    return detected / len(gt_anno_ids)
