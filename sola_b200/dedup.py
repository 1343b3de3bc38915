"""Greedy `miou_thresh` track de-duplication — the filtering logic of `generate_tokens_grid.py` and
`generate_tokens_gdino.py`, with SAM2 propagation left to the caller.

Two layers:
  * `GreedyState` — pure host logic (no GPU): candidate selection, batching rules and in-order suppression, a
    restatement of generate_tokens_grid.py:133-139,148-195,266-278,287-292 and
    generate_tokens_gdino.py:155-206,288-300,311-313.  It consumes rows of the gathered IoU matrix.
  * `TrackDedup` — the device session: prompt masks are nearest-resized and bit-packed ONCE (the reference redoes
    the H2D copy + resize for every (tracked, remaining) pair, generate_tokens_grid.py:271-272); every tracked batch
    is binarised/packed (K1), bilinear-resized (R1) and compared against all prompts in one launch (K2 gather);
    one small D2H per batch feeds `GreedyState`.
Also `dedup_matrix` — the spatio-temporal N x N variant (seg_utils.compute_masklet_iou semantics) that BASELINE
configs 2 and 5 name.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import packed as P

NOT_TRACKED, TRACKED, FILTERED, NOT_USED = 0, 1, 2, 3      # generate_tokens_grid.py:136


class GreedyState:
    """Host-side state machine.  `prompts`: list of dicts with at least prompt_id, frame_idx (and, for mode
    'gdino', expression_id and stability_score), in file order (area-descending, generate_prompts_grid.py:131-133)."""

    def __init__(self, prompts: Sequence[dict], n_frames: int, *, mode: str = "grid", bin_size: int = 4,
                 n_max_tracks: Optional[int] = None, batch_size: int = 4, miou_thresh: float = 0.7,
                 stability_score_thresh: float = 0.85, expression_id=None):
        assert mode in ("grid", "gdino")
        self.mode = mode
        self.n_frames = n_frames
        self.bin_size = bin_size
        self.n_max_tracks = n_max_tracks if n_max_tracks is not None else (64 if mode == "grid" else 16)
        self.batch_size = batch_size
        self.miou_thresh = miou_thresh
        self.n_tracked = self.n_filtered = self.n_not_used = 0
        self.batches: List[List] = []
        self.filtered_by: Dict = {}
        self.filtered_iou: Dict = {}
        # candidate list (indices into `prompts`) and per-prompt status
        self.prompts = list(prompts)
        self.status = np.full(len(self.prompts), NOT_USED, dtype=np.int8)
        self.members: List[int] = []          # prompts taking part in batching / suppression, in order
        for k, p in enumerate(self.prompts):
            if mode == "grid":
                # every prompt of the video takes part; off-bin ones are marked not-used (grid :134-139)
                self.members.append(k)
                if p["frame_idx"] % bin_size != 0:
                    self.n_not_used += 1
                else:
                    self.status[k] = NOT_TRACKED
            else:
                if p["expression_id"] != expression_id:
                    continue
                # gdino :162 — `score < thresh` is False for NaN and for score == thresh: both are kept
                if p["frame_idx"] % bin_size != 0 or p["stability_score"] < stability_score_thresh:
                    self.n_not_used += 1
                else:
                    self.status[k] = NOT_TRACKED
                    self.members.append(k)
        self.frame_idx = np.array([p["frame_idx"] for p in self.prompts], dtype=np.int64)
        if self.frame_idx.size and (self.frame_idx.min() < 0 or self.frame_idx.max() >= n_frames):
            # masklets[prompt_id][frame_idx] raises in the reference (generate_tokens_grid.py:273); never compare against a wrong frame
            raise IndexError(f"prompt frame_idx outside [0, {n_frames}): min {self.frame_idx.min()}, max {self.frame_idx.max()}")
        self._members_arr = np.asarray(self.members, dtype=np.int64)

    # -- batching ------------------------------------------------------------------------------------------------
    def next_batch(self) -> Optional[List[int]]:
        """Indices (into `prompts`) of the next same-frame batch to hand to SAM2, or None when finished.
        Members are marked tracked immediately, so batch-mates never suppress each other."""
        if self.n_tracked >= self.n_max_tracks:
            return None
        cap = 2 if self.n_frames > 200 else self.batch_size
        frame, batch = None, []
        for k in self.members:
            if self.status[k] != NOT_TRACKED:
                continue
            if frame is None:
                frame = self.frame_idx[k]
            elif self.frame_idx[k] != frame:
                if self.mode == "grid":
                    continue                               # grid :178-179 keeps scanning the whole list
                break                                      # gdino :194-196 stops at the first other-frame candidate
            batch.append(k)
            self.status[k] = TRACKED
            if self.mode == "grid":
                if len(batch) >= cap or self.n_tracked + len(batch) >= self.n_max_tracks:       # grid :181-186
                    break
            else:
                self.n_tracked += 1                        # gdino :187,193 — counted at append time
                if self.n_frames > 200 and len(batch) >= 2:
                    break
                if len(batch) >= self.batch_size:
                    break
                if len(batch) + self.n_tracked >= self.n_max_tracks:                            # gdino :201 (double count)
                    break
        if frame is None:
            return None
        if self.mode == "grid":
            self.n_tracked += len(batch)                   # grid :194
        self.batches.append([self.prompts[k]["prompt_id"] for k in batch])
        return batch

    # -- suppression -----------------------------------------------------------------------------------------------
    def apply_iou_rows(self, batch: Sequence[int], iou_rows: np.ndarray) -> int:
        """iou_rows[b, k] = IoU(resized track of batch[b] at frame_idx[k], nearest-resized prompt k) for every
        prompt k (float64).  Walks batch members in order and candidates in list order, strict `>` (grid :266-278)."""
        n = 0
        members = self._members_arr
        for b, member in enumerate(batch):
            row = np.asarray(iou_rows[b])
            # candidates still untracked, in list order; a member's hits do not depend on each other, only on the
            # state left by the previous members, so one vectorised pass per member reproduces the reference's inner loop
            hit = members[(self.status[members] == NOT_TRACKED) & (row[members] > self.miou_thresh)]
            if hit.size == 0:
                continue
            self.status[hit] = FILTERED
            member_id = self.prompts[member]["prompt_id"]
            for k in hit.tolist():
                pid = self.prompts[k]["prompt_id"]
                self.filtered_by[pid] = member_id
                self.filtered_iou[pid] = float(row[k])
            n += int(hit.size)
        self.n_filtered += n
        return n

    def result(self) -> dict:
        ids = lambda s: [self.prompts[k]["prompt_id"] for k in self.members if self.status[k] == s]
        res = {
            "status": {self.prompts[k]["prompt_id"]: int(self.status[k]) for k in self.members},
            "tracked": ids(TRACKED), "filtered": ids(FILTERED), "not_tracked": ids(NOT_TRACKED),
            # gdino :311 lists status==3 among the *candidates*, which is empty by construction
            "not_used": ids(NOT_USED) if self.mode == "grid" else [],
            "batches": self.batches, "n_tracked": self.n_tracked, "n_filtered": self.n_filtered,
            "n_not_used": self.n_not_used,
            "filtered_by": dict(self.filtered_by), "filtered_iou": dict(self.filtered_iou),
        }
        if self.mode == "grid" and len(res["tracked"]) < self.n_max_tracks:
            assert not res["not_tracked"], f"NOT TRACKED PROMPT MASKS ARE FOUND: {res['not_tracked']}"     # grid :291-292
        return res


def iou_from_counts(inter, area_a, area_b) -> np.ndarray:
    """float64 inter/union with the reference's empty rule (union == 0 -> 1.0, seg_utils.py:139-140)."""
    inter = np.asarray(inter, dtype=np.int64)
    union = np.asarray(area_a, dtype=np.int64) + np.asarray(area_b, dtype=np.int64) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / union
    return np.where(union == 0, 1.0, iou)


class TrackDedup:
    """Device session around `GreedyState` for one video (grid) or one expression (gdino).

    prompts[k]['segmentation'] is the decoded (H, W) uint8 prompt mask.  Typical loop:

        dd = TrackDedup(prompts, n_frames, mode='grid')
        while (batch := dd.next_batch()) is not None:
            logits = sam2_propagate([dd.prompts[k] for k in batch])      # (B, T, H, W) fp32, stays on the GPU
            packed, counts = dd.submit_logits(batch, logits)             # K1 + R1 + K2-gather + suppression
        result = dd.result()
    """

    def __init__(self, prompts: Sequence[dict], n_frames: int, *, device=None, target_shape=None,
                 prompt_masks: Optional[torch.Tensor] = None, **rules):
        """`prompt_masks`: optional (P, H, W) uint8 tensor (host or device) holding every prompt's mask in list order;
        when omitted the masks are taken from prompts[k]['segmentation']."""
        self.state = GreedyState(prompts, n_frames, **rules)
        self.prompts = self.state.prompts
        self.n_frames = n_frames
        if prompt_masks is None and self.prompts:
            prompt_masks = np.stack([np.asarray(p["segmentation"]) for p in self.prompts]).astype(np.uint8, copy=False)
        if prompt_masks is not None and len(prompt_masks):
            self.device = prompt_masks.device if isinstance(prompt_masks, torch.Tensor) and prompt_masks.is_cuda else P._dev(device)
            H, W = int(prompt_masks.shape[-2]), int(prompt_masks.shape[-1])
            self.native_shape = (H, W)
            self.target_shape = tuple(target_shape) if target_shape is not None else P.default_target_shape(H, W)
            # R2 hoisted: one H2D of uint8 (1 B/px) and one nearest-resize launch for all prompts of the unit
            self.prompt_planes = P.resize_nearest(P.to_device(prompt_masks, device=self.device), *self.target_shape)
            self.frame_idx_dev = P.to_device(self.state.frame_idx.astype(np.int32), device=self.device)
        else:
            self.device = P._dev(device)
            self.native_shape = self.target_shape = None
            self.prompt_planes = self.frame_idx_dev = None

    def next_batch(self):
        return self.state.next_batch()

    def submit_logits(self, batch, logits, mask_threshold: float = 0.0, threshold_offset: float = 1.0):
        """logits (B, T, H, W) fp32/bf16 of the batch members, in batch order.  Returns (native-resolution packed
        masklets, stability counts) so the caller can RLE-encode / store them."""
        packed, counts, resized = P.binarize_pack_resize(logits, mask_threshold, threshold_offset, self.target_shape)   # K1 + R1, one pass
        self.submit_resized(batch, resized)
        return packed, counts

    def submit_masks(self, batch, masklets):
        """masklets (B, T, H, W) {0,1} fp32 / uint8 (what `torch.cat(...)` holds in the reference, grid :223-225)."""
        packed = P.pack_masks(masklets)
        self.submit_packed(batch, packed)
        return packed

    def submit_packed(self, batch, packed: P.PackedMasks) -> np.ndarray:
        resized = P.resize_bilinear_bin(packed, self.target_shape)                   # R1 (seg_utils.reshape_masklet)
        return self.submit_resized(batch, resized)

    def submit_resized(self, batch, resized: P.PackedMasks) -> np.ndarray:
        c = P.gathered_inter(resized, self.prompt_planes, self.frame_idx_dev).cpu().numpy()     # one D2H per batch
        rows = iou_from_counts(c[0], c[1], c[2])
        self.state.apply_iou_rows(batch, rows)
        return rows

    def run_offline(self, resized: P.PackedMasks) -> dict:
        """All candidates' (resized, packed) masklets are already known — `resized` is (P, T, h, wp) in prompt order.
        The outcome of the reference loop is a pure function of M[i][j] = IoU(track_i[frame_j], prompt_j) and the
        batching rules, so M is computed in ONE launch and read back once; the host then replays the loop."""
        c = P.gathered_inter(resized, self.prompt_planes, self.frame_idx_dev).cpu().numpy()
        M = iou_from_counts(c[0], c[1], c[2])
        while (batch := self.state.next_batch()) is not None:
            self.state.apply_iou_rows(batch, M[batch])
        return self.state.result()

    def result(self) -> dict:
        return self.state.result()


def dedup_matrix(tracks: P.PackedMasks, miou_thresh: float = 0.7):
    """Spatio-temporal de-duplication of N candidate tracks (BASELINE configs 2 / 5): full N x N IoU from exact
    integer intersections, then index-order greedy suppression with the same strict `>`.
    Returns (kept ids, {suppressed: suppressor}, float64 IoU matrix, int64 intersection matrix)."""
    inter = P.pairwise_inter_matrix(tracks).cpu().numpy()
    iou = P.iou_matrix_from_inter(inter)
    n = iou.shape[0]
    alive = np.ones(n, dtype=bool)
    by = {}
    for i in range(n):
        if not alive[i]:
            continue
        hit = np.nonzero(alive[i + 1:] & (iou[i, i + 1:] > miou_thresh))[0] + i + 1
        alive[hit] = False
        for j in hit.tolist():
            by[j] = i
    return [i for i in range(n) if alive[i]], by, iou, inter


class VideoDedupJob:
    """Whole-video pass with the device work and the host work decoupled, for streaming many videos:

        enqueue(logits, prompt_masks)  K1 -> R1 -> R2 -> K2-gather -> K2 N x N, then async read-back of the three small
                                       result tensors into pinned host memory; returns immediately (no synchronisation)
        finish()                       waits for THIS job's event only, then replays the reference greedy loop, the
                                       spatio-temporal greedy and the float64 stability scores on the host

    Two jobs used alternately keep the GPU busy while the host post-processes the previous video (bench.py).
    With `tail_stream` set, everything after K1+R1 (R2, K2 gather, K2 N x N, label counts, read-backs: integer-pipe and latency-bound
    kernels on a few hundred MB) is issued on that stream behind an event, so that it runs UNDER the HBM-bound K1+R1 of the next
    video, which the caller enqueues on its own stream with per-job output buffers."""

    def __init__(self, prompt_meta: Sequence[dict], n_frames: int, *, device=None, mode: str = "grid",
                 st_on_resized: bool = True, fused: bool = True, tail_stream: Optional[torch.cuda.Stream] = None,
                 aux_stream: Optional[torch.cuda.Stream] = None, **rules):
        self.st_on_resized = st_on_resized
        self.fused = fused
        self.tail_stream = tail_stream
        self._k1_done = torch.cuda.Event() if tail_stream is not None else None
        # `aux_stream`: R2 runs beside K1+R1 (enqueue_prompts), and K2 gather + label counts run beside K2 N x N
        self.aux_stream = aux_stream
        self._fork, self._join = (torch.cuda.Event(), torch.cuda.Event()) if aux_stream is not None else (None, None)
        self._planes: Optional[P.PackedMasks] = None
        self.timing_events = None            # set to a list to collect (start, end) CUDA events around the K2 N x N launch
        self.prompt_meta = list(prompt_meta)
        self.n_frames = n_frames
        self.mode, self.rules = mode, rules
        self.device = P._dev(device)
        self.frame_idx = np.array([p["frame_idx"] for p in self.prompt_meta], dtype=np.int32)
        if self.frame_idx.size and (self.frame_idx.min() < 0 or self.frame_idx.max() >= n_frames):
            raise IndexError(f"prompt frame_idx outside [0, {n_frames})")
        self.frame_idx_dev = P.to_device(self.frame_idx, device=self.device)
        self.event = torch.cuda.Event()
        self._host = None
        self._host_labels = None
        self.packed = self.counts = self.resized = None
        self.gt_planes: Optional[P.PackedMasks] = None      # (G, T, h, wp) GT masklets at the resized shape -> label metrics in the step

    def _pinned(self, n_tracks: int, n_prompts: int, T: int):
        shapes = ((3, n_tracks, n_prompts), (n_tracks, n_tracks), (3, n_tracks, T))
        if self._host is None or tuple(tuple(t.shape) for t in self._host) != shapes:
            self._host = (torch.empty((3, n_tracks, n_prompts), dtype=torch.int32).pin_memory(),
                          torch.empty((n_tracks, n_tracks), dtype=torch.int64).pin_memory(),
                          torch.empty((3, n_tracks, T), dtype=torch.int32).pin_memory())
        return self._host

    def set_gt_masklets(self, gt_planes: Optional[P.PackedMasks]) -> None:
        """GT masklets of the video, (G, T, h, wp) packed at the resized shape (what `gt_masklets` holds at generate_tokens_grid.py:112).
        When set, every step also produces the training labels of generate_tokens_grid.py:253-264: frame-mean precision / recall / IoU of
        every track against every GT object (utils.compute_mask_metrics), from ONE batched count launch on the resized planes."""
        self.gt_planes = gt_planes

    def enqueue(self, logits: torch.Tensor, prompt_masks: torch.Tensor, *, mask_threshold: float = 0.0, threshold_offset: float = 1.0,
                packed_out: Optional[P.PackedMasks] = None, counts_out: Optional[torch.Tensor] = None, target_shape=None):
        """logits (N, T, H, W) fp32/bf16 and prompt_masks (N, H, W) uint8, both on the device, in prompt order."""
        if self.aux_stream is not None:
            self.enqueue_prompts(prompt_masks, target_shape)                                                          # R2 beside K1+R1
        if self.fused:
            # K1 + R1 in one pass: the resize work hides under the HBM time of reading the logits
            packed, counts, resized = P.binarize_pack_resize(logits, mask_threshold, threshold_offset, target_shape,
                                                             out=packed_out, counts_out=counts_out)
            self._enqueue_tail(packed, resized, counts, prompt_masks)
        else:
            packed, counts = P.binarize_pack_stability(logits, mask_threshold, threshold_offset, out=packed_out, counts_out=counts_out)   # K1
            self.enqueue_after_k1(packed, counts, prompt_masks, target_shape=target_shape)

    def enqueue_after_k1(self, packed: P.PackedMasks, counts: torch.Tensor, prompt_masks: torch.Tensor, *, target_shape=None):
        """Same, for callers that already ran K1 (e.g. chunk by chunk behind H2D copies)."""
        self._enqueue_tail(packed, P.resize_bilinear_bin(packed, target_shape), counts, prompt_masks)                # R1

    def enqueue_prompts(self, prompt_masks: torch.Tensor, target_shape=None) -> None:
        """R2 ahead of time: the nearest resize + pack of the prompt masks depends on nothing K1+R1 produces, so with an `aux_stream`
        a caller may issue it BEFORE the K1+R1 launch and it runs beside that kernel instead of after it."""
        oh, ow = P.default_target_shape(int(prompt_masks.shape[-2]), int(prompt_masks.shape[-1])) if target_shape is None else target_shape
        if self.aux_stream is None:
            self._planes = P.resize_nearest(prompt_masks, oh, ow)
            return
        self._fork.record(torch.cuda.current_stream(self.device))
        self.aux_stream.wait_event(self._fork)
        with torch.cuda.stream(self.aux_stream):
            prompt_masks.record_stream(self.aux_stream)
            self._planes = P.resize_nearest(prompt_masks, oh, ow)

    def _enqueue_tail(self, packed: P.PackedMasks, resized: P.PackedMasks, counts: torch.Tensor, prompt_masks: torch.Tensor):
        if self.tail_stream is None:
            return self._tail(packed, resized, counts, prompt_masks)
        self._k1_done.record(torch.cuda.current_stream(self.device))
        self.tail_stream.wait_event(self._k1_done)
        with torch.cuda.stream(self.tail_stream):
            for t in (packed.words if packed is not None else None, resized.words, counts, prompt_masks):
                if t is not None:
                    t.record_stream(self.tail_stream)
            self._tail(packed, resized, counts, prompt_masks)

    def _tail(self, packed: Optional[P.PackedMasks], resized: P.PackedMasks, counts: torch.Tensor, prompt_masks: torch.Tensor):
        N, T = int(resized.words.shape[0]), int(resized.words.shape[1])
        self.packed, self.resized = packed, resized
        main = torch.cuda.current_stream(self.device)
        aux = self.aux_stream
        if self._planes is None or (self._planes.H, self._planes.W) != (resized.H, resized.W):
            self._planes = None
            self.enqueue_prompts(prompt_masks, (resized.H, resized.W))                                                # R2
        planes, self._planes = self._planes, None
        hg, hi, hc = self._pinned(N, int(planes.words.shape[0]), T)

        def side():
            """K2 gather + label counts + their read-backs: independent of K2 N x N (all three only read the resized planes)."""
            g = P.gathered_inter(resized, planes, self.frame_idx_dev)                                                  # K2 gather
            hg.copy_(g, non_blocking=True)
            if self.gt_planes is not None:
                # label metrics (generate_tokens_grid.py:253-264): per-frame |track ∩ gt|, |track|, |gt| for all N x G x T frame pairs
                G = int(self.gt_planes.words.shape[0])
                li, la, lb = P.frame_counts_packed(resized, self.gt_planes)
                shapes = ((N, G, T), (N, T), (G, T))
                if self._host_labels is None or tuple(tuple(t.shape) for t in self._host_labels) != shapes:
                    self._host_labels = tuple(torch.empty(sh, dtype=torch.int32).pin_memory() for sh in shapes)
                for h, d in zip(self._host_labels, (li, la, lb)):
                    h.copy_(d, non_blocking=True)

        if aux is not None:
            # the integer-pipe-bound N x N kernel on this stream, the latency / POPC-bound gather and label counts beside it
            self._fork.record(main)
            aux.wait_event(self._fork)
            with torch.cuda.stream(aux):
                resized.words.record_stream(aux)
                side()
        # K2 N x N on the resized planes: after generate_tokens_grid.py:248-250 only the 540x960 masklets exist, so a
        # masklet-vs-masklet IoU (seg_utils.compute_masklet_iou) in that flow compares those
        if self.timing_events is not None:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
        inter = P.pairwise_inter_matrix(self.resized if self.st_on_resized else self.packed)                           # K2 N x N
        if self.timing_events is not None:
            eb.record()
            self.timing_events.append((ea, eb))
        if aux is not None:
            # the read-backs leave on the aux stream as well, so the launching stream goes straight from K2 to the next video's
            # K1+R1; this job's outputs (resized planes, counts) must not be overwritten before finish() — callers alternate two sets
            self._join.record(main)
            aux.wait_event(self._join)
            with torch.cuda.stream(aux):
                inter.record_stream(aux)
                counts.record_stream(aux)
                hi.copy_(inter, non_blocking=True)
                hc.copy_(counts.reshape(3, N, T), non_blocking=True)
                self.event.record(aux)
        else:
            hi.copy_(inter, non_blocking=True)
            hc.copy_(counts.reshape(3, N, T), non_blocking=True)
            side()
            self.event.record(main)
        self.counts = counts

    def finish(self, miou_thresh_st: Optional[float] = None) -> dict:
        self.event.synchronize()
        hg, hi, hc = (t.numpy() for t in self._host)
        state = GreedyState(self.prompt_meta, self.n_frames, mode=self.mode, **self.rules)
        M = iou_from_counts(hg[0], hg[1], hg[2])
        while (batch := state.next_batch()) is not None:
            state.apply_iou_rows(batch, M[batch])
        res = state.result()
        thr = state.miou_thresh if miou_thresh_st is None else miou_thresh_st
        iou = P.iou_matrix_from_inter(hi)
        n = iou.shape[0]
        alive = np.ones(n, dtype=bool)
        by = {}
        for i in range(n):
            if not alive[i]:
                continue
            hit = np.nonzero(alive[i + 1:] & (iou[i, i + 1:] > thr))[0] + i + 1
            alive[hit] = False
            for j in hit.tolist():
                by[j] = i
        with np.errstate(divide="ignore", invalid="ignore"):
            stability = hc[0] / hc[2]
        res.update({"kept_spatiotemporal": np.nonzero(alive)[0].tolist(), "suppressed_by_spatiotemporal": by,
                    "inter": hi.copy(), "stability": stability, "iou_gather": M})
        if self.gt_planes is not None and self._host_labels is not None:
            res["labels"] = label_metrics_from_counts(*(t.numpy() for t in self._host_labels))
        return res


def label_metrics_from_counts(inter: np.ndarray, area_p: np.ndarray, area_g: np.ndarray) -> dict:
    """inter (N, G, T), area_p (N, T), area_g (G, T) -> {'precision', 'recall', 'iou'}: fp32 (N, G) arrays holding what
    utils.compute_mask_metrics(...)[k].squeeze().item() returns for every (track, GT object) (utils.py:132-174): per-frame float64
    ratios with the four empty-case rules, stored as fp32, then an fp32 mean over the T frames."""
    inter = inter.astype(np.int64)
    n_p = area_p.astype(np.int64)[:, None, :]
    n_g = area_g.astype(np.int64)[None, :, :]
    union = n_p + n_g - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(union == 0, 1.0, inter / union)
        prec = np.where(n_p == 0, 1.0, np.where(n_g == 0, 0.0, inter / n_p))
        rec = np.where(n_p == 0, np.where(n_g == 0, 1.0, 0.0), np.where(n_g == 0, 1.0, inter / n_g))
    # torch's fp32 .mean() over a contiguous (T,) tensor; reproduce it with torch so the rounding is the reference's
    as_mean = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).float().mean(dim=-1).numpy()
    return {"precision": as_mean(prec), "recall": as_mean(rec), "iou": as_mean(iou)}
