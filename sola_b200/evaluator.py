"""Drop-in mirror of the J / F part of `evaluator.py` (Evaluator.compute_J, compute_F, compute_JF_metrics).

`F` here is what the reference computes: a volumetric pixel F1 over the whole (T, H, W) masklet
(evaluator.py:239-247).  The boundary F-measure that BASELINE.json's north-star also names has no reference
implementation; it is provided separately as `compute_F_boundary` (DAVIS definition, parity unpinned)."""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np
import torch

from . import packed as P


# ---- host formulas over exact integer counts -------------------------------------------------------------------

def J_from_counts(inter, n_pred, n_gt):
    """evaluator.py:227-237 — mean over frames of inter/union (1.0 for an empty union), numpy float64.  Vectorised: the counts are exact
    integers < 2**53, so the float64 quotients equal the reference's Python-float divisions, and np.mean sums them in the same order."""
    i, p, g = (np.asarray(x, dtype=np.int64) for x in (inter, n_pred, n_gt))
    union = p + g - i
    with np.errstate(divide="ignore", invalid="ignore"):
        Js = np.where(union == 0, 1.0, i / np.where(union == 0, 1, union))
    return np.mean(Js)


def F_from_counts(inter, n_pred, n_gt) -> float:
    """evaluator.py:239-247 — tp/fp/fn summed over the volume; 0.0 when tp == 0."""
    tp = int(np.sum(inter, dtype=np.int64))
    fp = int(np.sum(n_pred, dtype=np.int64)) - tp
    fn = int(np.sum(n_gt, dtype=np.int64)) - tp
    if tp == 0:
        return 0.0
    precision = tp / (tp + fp)
    recall = tp / (tp + fn)
    return 2 * precision * recall / (precision + recall)


def F_boundary_from_counts(n_fg, n_gt, fg_match, gt_match) -> float:
    """DAVIS boundary F per frame, averaged over frames (oracle/boundary_oracle.py is the spec): precision = fg_match / n_fg,
    recall = gt_match / n_gt with the three empty-boundary rules, F = 2PR / (P + R) (0 when P + R == 0)."""
    a, b, fm, gm = (np.asarray(x, dtype=np.int64) for x in (n_fg, n_gt, fg_match, gt_match))
    with np.errstate(divide="ignore", invalid="ignore"):
        p = np.where(a == 0, 1.0, np.where(b == 0, 0.0, fm / np.where(a == 0, 1, a)))
        r = np.where(a == 0, np.where(b == 0, 1.0, 0.0), np.where(b == 0, 1.0, gm / np.where(b == 0, 1, b)))
        f = np.where(p + r == 0, 0.0, 2 * p * r / np.where(p + r == 0, 1.0, p + r))
    return float(np.mean(f))


def sweep_metrics_from_counts(counts, offsets, frames, with_boundary: bool = False):
    """The three formulas above for EVERY unit of a sweep at once.  counts: (7, total_frames) host array as the fused kernel writes
    it; unit u occupies columns offsets[u] : offsets[u] + frames[u].  Returns (J, F, F_boundary | None, int_totals) — float64 arrays of
    one value per unit, bit-equal to calling J_from_counts / F_from_counts / F_boundary_from_counts unit by unit: the per-frame ratios
    are elementwise float64 operations (evaluated here in one pass over all frames), the volume sums are exact integers
    (np.add.reduceat), and the per-unit frame means still go through np.mean on the unit's own contiguous slice, so the summation
    order is the reference's.  A per-unit Python loop over small arrays costs 25-45 us per unit — more than the kernel's share."""
    c = np.asarray(counts)
    offsets = np.asarray(offsets, dtype=np.int64)
    frames = np.asarray(frames, dtype=np.int64)
    n = len(frames)
    J, F = np.zeros(n, np.float64), np.zeros(n, np.float64)
    Fb = np.zeros(n, np.float64) if with_boundary else None
    live = np.nonzero(frames > 0)[0]
    i, p, g = (c[k].astype(np.int64) for k in range(3))
    union = p + g - i
    with np.errstate(divide="ignore", invalid="ignore"):
        Js = np.where(union == 0, 1.0, i / np.where(union == 0, 1, union))
        if with_boundary:
            a, b, fm, gm = (c[k].astype(np.int64) for k in range(3, 7))
            pr = np.where(a == 0, 1.0, np.where(b == 0, 0.0, fm / np.where(a == 0, 1, a)))
            rc = np.where(a == 0, np.where(b == 0, 1.0, 0.0), np.where(b == 0, 1.0, gm / np.where(b == 0, 1, b)))
            fb = np.where(pr + rc == 0, 0.0, 2 * pr * rc / np.where(pr + rc == 0, 1.0, pr + rc))
    sums = np.zeros((3, n), np.int64)
    if len(live):
        # reduceat needs strictly usable start indices: reduce over the live units only (empty units keep sum 0)
        order = live[np.argsort(offsets[live], kind="stable")]
        starts = offsets[order]
        assert np.all(starts[1:] >= starts[:-1] + frames[order][:-1]) and starts[-1] + frames[order][-1] <= c.shape[1], "units overlap / exceed the table"
        contiguous = np.all(starts[1:] == starts[:-1] + frames[order][:-1])
        for k, v in enumerate((i, p, g)):
            if contiguous:
                sums[k, order] = np.add.reduceat(v[: starts[-1] + frames[order][-1]], starts)
            else:
                sums[k, order] = [int(v[s: s + t].sum()) for s, t in zip(starts, frames[order])]
    tp, fp, fn = sums[0], sums[1] - sums[0], sums[2] - sums[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        tpf = tp.astype(np.float64)                       # exact: volume sums < 2**53
        prec, rec = tpf / (tp + fp), tpf / (tp + fn)
        F = np.where(tp == 0, 0.0, 2 * prec * rec / (prec + rec))
    for u in range(n):
        o, t = int(offsets[u]), int(frames[u])
        J[u] = np.mean(Js[o: o + t]) if t else np.float64("nan")
        if with_boundary:
            Fb[u] = np.mean(fb[o: o + t]) if t else np.float64("nan")
    return J, F, Fb, sums.sum(axis=1)


# ---- per-call drop-ins -----------------------------------------------------------------------------------------

_last_counts = None       # (key, counts): the reference calls compute_J then compute_F on the SAME two tensors (evaluator.py:201-202)


def _tensor_key(t):
    return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.dtype, str(t.device)) if isinstance(t, torch.Tensor) else None


def jf_counts(pred_masklet, gt_masklet) -> np.ndarray:
    """(3, T) int32 host array [inter, n_pred, n_gt] for {0,1} masklets (fp32 / uint8; device, CPU or numpy).
    The last result is kept, keyed by both tensors' storage pointer AND version counter (any in-place write invalidates it), so the
    reference's `compute_J(p, g)` followed by `compute_F(p, g)` reads the masklets once and synchronises once instead of twice."""
    global _last_counts
    ka, kb = _tensor_key(pred_masklet), _tensor_key(gt_masklet)
    key = (ka, kb) if ka is not None and kb is not None else None
    if key is not None and _last_counts is not None and _last_counts[0] == key:
        return _last_counts[1]
    c = P.frame_counts(pred_masklet, gt_masklet).cpu().numpy()
    # hold the tensors too: a freed and re-allocated buffer could otherwise reproduce (pointer, version 0) with other contents
    _last_counts = (key, c, pred_masklet, gt_masklet) if key is not None else None
    return c


def compute_J(pred_masklet, gt_masklet):
    c = jf_counts(pred_masklet, gt_masklet)
    return J_from_counts(c[0], c[1], c[2])


def compute_F(pred_masklet, gt_masklet) -> float:
    c = jf_counts(pred_masklet, gt_masklet)
    return F_from_counts(c[0], c[1], c[2])


def compute_JF(pred_masklet, gt_masklet):
    """J and F from ONE pass over the two masklets (the reference reads them 2*T + 8 times)."""
    c = jf_counts(pred_masklet, gt_masklet)
    return J_from_counts(c[0], c[1], c[2]), F_from_counts(c[0], c[1], c[2])


def jf_from_accumulators(inter, uni, totals):
    """(J, F) from the device accumulators of packed.jf_accumulators / the C ABI's sola_jf_*: J = np.mean of the per-frame ratios with
    the union == 0 -> 1.0 rule (evaluator.py:231-236); F = 2PR/(P+R) from tp / fp / fn with the tp == 0 -> 0.0 rule (evaluator.py:243-247)."""
    i = inter.cpu().numpy().astype(np.float64) if isinstance(inter, torch.Tensor) else np.asarray(inter, np.float64)
    u = uni.cpu().numpy().astype(np.float64) if isinstance(uni, torch.Tensor) else np.asarray(uni, np.float64)
    tp, fp, fn = (int(x) for x in (totals.cpu().tolist() if isinstance(totals, torch.Tensor) else totals))
    with np.errstate(divide="ignore", invalid="ignore"):
        J = np.mean(np.where(u == 0, 1.0, i / np.where(u == 0, 1.0, u))) if len(u) else np.float64("nan")
    if tp == 0:
        return J, 0.0
    prec, rec = tp / (tp + fp), tp / (tp + fn)
    return J, 2 * prec * rec / (prec + rec)


def compute_F_boundary(pred_masklet, gt_masklet, bound_th: float = 0.008) -> float:
    pp = pred_masklet if isinstance(pred_masklet, P.PackedMasks) else P.pack_masks(pred_masklet)
    gp = gt_masklet if isinstance(gt_masklet, P.PackedMasks) else P.pack_masks(gt_masklet)
    c = P.boundary_counts(pp, gp, bound_th).cpu().numpy()
    return F_boundary_from_counts(c[0], c[1], c[2], c[3])


def compute_JF_all(pred_masklet, gt_masklet, bound_th: float = 0.008):
    """(J, F_ref_dice, F_boundary) of one unit from ONE launch of the fused kernel (csrc/jf_fused.cu) and one read-back:
    J per evaluator.py:227-237, F per evaluator.py:239-247, F_boundary per the DAVIS definition (oracle/boundary_oracle.py)."""
    pp = pred_masklet if isinstance(pred_masklet, P.PackedMasks) else P.pack_masks(pred_masklet)
    gp = gt_masklet if isinstance(gt_masklet, P.PackedMasks) else P.pack_masks(gt_masklet)
    c = P.jf_boundary_counts(pp, gp, bound_th).cpu().numpy()
    return (float(J_from_counts(c[0], c[1], c[2])), float(F_from_counts(c[0], c[1], c[2])),
            F_boundary_from_counts(c[3], c[4], c[5], c[6]))


def _as_packed(m, device) -> P.PackedMasks:
    """A masklet in any accepted form -> (T, H, Wp) PackedMasks on `device` (uint8 / bool / fp32 arrays are uploaded as they are
    — 1 B/px for the dataloader's uint8 masks instead of the reference's fp32 4 B/px, evaluator.py:199-200 — and bit-packed there)."""
    if isinstance(m, P.PackedMasks):
        return m if m.words.dim() == 3 else m.reshape_lead(m.n_frames)
    x = P.to_device(m, device=device)
    if x.dim() == 2:
        x = x[None]
    return P.pack_masks(x)


class JFSweep:
    """Batched J&F over many (video, expression) units of different shapes.  add() uploads / bit-packs a unit and returns
    immediately; finish() runs ONE launch of the fused J&F kernel over every unit added (region counts and, with
    `with_boundary`, the boundary-match counts from the same staged tile), reads the (7, total_frames) count table back ONCE and
    evaluates the reference's formulas on the host in float64.  It also returns the integer accumulators
    [Σ inter, Σ|pred|, Σ|gt|] that make multi-GPU reductions bit-reproducible."""

    def __init__(self, device=None, with_boundary: bool = False, bound_th: float = 0.008):
        self.device = P._dev(device)
        self.with_boundary = with_boundary
        self.bound_th = bound_th
        self._keys, self._pairs = [], []

    def add(self, key, pred_masklet, gt_masklet) -> None:
        """pred_masklet None reproduces evaluator.py:194-197 (J = F = JF = 0)."""
        self._keys.append(key)
        if pred_masklet is None:
            self._pairs.append(None)
            return
        p, g = _as_packed(pred_masklet, self.device), _as_packed(gt_masklet, self.device)
        assert (p.H, p.W) == (g.H, g.W) and p.n_frames == g.n_frames, f"pred / gt masklets differ in shape for {key!r}"
        self._pairs.append((p, g))

    def finish(self):
        live = [pg for pg in self._pairs if pg is not None]
        plan = P.JFSweepPlan(live, with_boundary=self.with_boundary, bound_th=self.bound_th)
        flat = plan.run().cpu().numpy() if live else np.zeros((7, 0), np.int32)
        Js, Fs, Fbs, totals = sweep_metrics_from_counts(flat, plan.offsets, plan.frames, self.with_boundary)
        results, k = [], 0
        for key, pg in zip(self._keys, self._pairs):
            if pg is None:
                results.append((key, {"J": 0.0, "F": 0.0, "JF": 0.0}))
                continue
            J, F = float(Js[k]), float(Fs[k])
            rec = {"J": J, "F": F, "JF": (J + F) / 2}
            if self.with_boundary:
                rec["F_boundary"] = float(Fbs[k])
            k += 1
            results.append((key, rec))
        self._keys, self._pairs = [], []
        return results, totals


class Evaluator:
    """J&F half of the reference `Evaluator` (evaluator.py:174-247) with the same method names.  Model inference
    (`evaluate`, evaluator.py:54-172) is out of scope: construct this with the `pred_dict` it produced, or graft the
    three methods onto a reference Evaluator instance."""

    def __init__(self, loader_dict=None, pred_dict=None, eval_output_dir: Optional[str] = None, data_type: str = "valid",
                 eval_weight_epoch: int = 0, device=None, dataset=None, with_boundary: bool = False):
        """`dataset`: the object compute_JF_metrics asks for masklets — defaults to loader_dict["valid"].dataset as in the reference
        (evaluator.py:176).  Pass `dataloader_ops.AlignDatasetAdapter.from_dataset(reference_dataset)` to keep RLE decode, OR-merge and
        counting on the device (bit-packed end to end).  `with_boundary` adds the north-star's boundary F as `F_boundary` per expression
        and `mean_F_boundary` (extension; the reference's JSON keys are unchanged)."""
        self.loader_dict = loader_dict
        self.dataset = dataset
        self.with_boundary = with_boundary
        self.pred_dict = pred_dict or {}
        self.eval_output_dir = eval_output_dir
        self.data_type = data_type
        self.eval_weight_epoch = eval_weight_epoch
        self.device = P._dev(device) if torch.cuda.is_available() else device
        self.metrics = {}

    def compute_J(self, pred_masklet, gt_masklet):
        return compute_J(pred_masklet, gt_masklet)

    def compute_F(self, pred_masklet, gt_masklet):
        return compute_F(pred_masklet, gt_masklet)

    def compute_JF_metrics(self):
        """evaluator.py:174-225 — same traversal order, same JSON schema, same float64 means; masks are counted by
        the batched sweep, one device read-back per video instead of 2*T + 3 `.item()` per expression."""
        dataset = self.dataset if self.dataset is not None else self.loader_dict["valid"].dataset
        JF_dict, Js, Fs, JFs, Fbs = {}, [], [], [], []
        for video_id in self.pred_dict:
            JF_dict[video_id] = {}
            dataset.set_video(video_id)
            sweep = JFSweep(self.device, with_boundary=self.with_boundary)
            for expression_id, pred_info in self.pred_dict[video_id].items():
                gt_masklet = dataset.get_gt_masklet(video_id, expression_id)
                pred_masklet = dataset.get_sam2_masklet(
                    video_id=video_id, expression_id=expression_id, preds=pred_info["pred"],
                    root_types=pred_info["root_type"], prompt_types=pred_info["prompt_type"],
                    sam2_anno_ids=pred_info["sam2_anno_id"])
                sweep.add(expression_id, pred_masklet, gt_masklet)
            results, _ = sweep.finish()
            for expression_id, rec in results:
                JF_dict[video_id][expression_id] = {
                    "expression": self.pred_dict[video_id][expression_id]["expression"],
                    "J": rec["J"], "F": rec["F"], "JF": rec["JF"],
                }
                Js.append(rec["J"]), Fs.append(rec["F"]), JFs.append(rec["JF"])
                if self.with_boundary and "F_boundary" in rec:
                    JF_dict[video_id][expression_id]["F_boundary"] = rec["F_boundary"]
                    Fbs.append(rec["F_boundary"])
        if Fbs:
            self.metrics["mean_F_boundary"] = np.mean(Fbs)
        self.metrics["mean_J"] = np.mean(Js)
        self.metrics["mean_F"] = np.mean(Fs)
        self.metrics["mean_JF"] = np.mean(JFs)
        if self.eval_output_dir is not None:
            path = os.path.join(self.eval_output_dir, f"{self.data_type}_JF_metrics_{self.eval_weight_epoch}epoch.json")
            with open(path, "w") as f:
                json.dump(JF_dict, f, indent=4)
        return JF_dict
