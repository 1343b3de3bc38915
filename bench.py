#!/usr/bin/env python
"""bench.py — masklet-frames/s of the masklet-scoring hot path (dedup + J&F) on BASELINE.json config 2
("Grid-prompt track dedup: 64 synthetic SAM2 masklets x 80 frames at 720x1280 with stability score and miou_thresh 0.7").

One step = one video's pass over the path, inputs already in HBM:
    K1+R1  binarise + bit-pack + stability counts over 64 x 80 fp32 logit planes, fused with the bilinear resize to 540 x 960 + > 0.5
           (generate_tokens_grid.py:215-224, prompt_generator.py:169-186, seg_utils.reshape_masklet)            dominant, HBM bound
    R2     nearest resize + pack of the 64 prompt masks
    G1     reference-order greedy filter: K2-gather IoU per tracked batch + host suppression (generate_tokens_grid.py:266-278)
    K2     64 x 64 spatio-temporal intersection matrix of the resized masklets + index-order greedy (compute_masklet_iou semantics)
    M1     the J-type half the grid loop really runs: per-frame |track ∩ gt|, |track|, |gt| of every track against G = 3 GT masklets
           -> frame-mean precision / recall / IoU labels (generate_tokens_grid.py:253-264, utils.compute_mask_metrics)
    one D2H of the count tables -> float64 scores on the host (overlapped with the next step)
Launches that only share read-only inputs run side by side on a second stream (dedup.VideoDedupJob(aux_stream=...)): R2 beside
K1+R1, the K2 gather, the label counts and the read-backs beside K2 N x N; `--no-aux` issues everything on one stream, `--overlap`
additionally puts the whole tail of video k under K1+R1 of video k+1 (measured slower for the dominant kernel; off).
`value` = masklet-frames processed by all ranks / max-over-ranks device time (weak scaling: one video per GPU per step, no
data-path collective).  `e2e` = the same step with the logits, prompt masks and GT masks starting in pinned HOST memory (H2D inside
the timed region, results read back).
A SECOND timed region measures the evaluation half of the metric on a BASELINE config-4-shaped sweep (MeViS-like: mixed 360p-1080p
units of 30-120 frames, bit-packed, resident in HBM): ONE launch of the fused J & F kernel per sweep (region counts + boundary-match
counts), one read-back, the reference's J / F formulas on the host, and — for N > 1 — the final NCCL sum of the accumulators.  It
fills `jf_stage` and `roofline_jf`.  With N > 1 a config-5-shaped slice (32 tracks per GPU x 200 frames of 1080p-derived planes) is
also pushed through the two exchange paths (peer-memory pull pipelined with K2, NCCL all-to-all) and checked against the single-rank
matrix (`cfg5`).
`--impl reference` times the oracle port of the reference's own CPU path for the config-2 step on the host cores: a FULL 64-track
step, timed once (`--steps` is honoured as min(steps, 1); the step takes tens of seconds).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "masklet-frames/s (dedup+J&F)"
UNIT = "masklet-frames/s"
CFG = dict(n_tracks=64, n_frames=80, H=720, W=1280, miou_thresh=0.7, n_max_tracks=64, batch_size=4, bin_size=4, n_gt=3)
JF_SWEEP = dict(n_videos=12, exprs_per_video=4, t_range=(30, 120))      # per GPU; config-4-shaped (MeViS valid_u: mixed 360p-1080p)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sola_b200", choices=["sola_b200", "reference"])
    ap.add_argument("--tracks", type=int, default=CFG["n_tracks"], help="override for quick local checks only")
    ap.add_argument("--frames", type=int, default=CFG["n_frames"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-wc", action="store_true",
                    help="stage the e2e logits in WRITE-COMBINED pinned host memory (cudaHostAllocWriteCombined): DMA reads do not snoop the "
                         "CPU caches, which matters when 8 ranks stream 18.9 GB per step out of one socket")
    ap.add_argument("--no-fuse", action="store_true", help="run K1 and R1 as two kernels instead of the fused one")
    ap.add_argument("--overlap", action="store_true",
                    help="run the tail of a step (R2, K2 gather, K2 N x N, labels) on a second high-priority stream under the next video's K1+R1. "
                         "Measured (profiles/r3_overlap_experiment.json): 3.50 vs 3.63 ms per step, but K1+R1 slows from 2.93 to 3.49 ms because "
                         "K2's two 288-thread, 96-register CTAs per SM displace the five K1+R1 CTAs it needs for its loads in flight; off by default")
    ap.add_argument("--no-aux", action="store_true", help="issue every kernel of a step on one stream (no R2 beside K1+R1, no gather / labels beside K2)")
    ap.add_argument("--no-jf", action="store_true", help="skip the config-4 J&F sweep region")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the config-5 exchange slice (N > 1 only)")
    ap.add_argument("--jf-reps", type=int, default=20)
    ap.add_argument("--st-native", action="store_true",
                    help="N x N spatio-temporal IoU on the native 720x1280 planes instead of the 540x960 resized masklets the "
                         "reference's filter works on (generate_tokens_grid.py:248-250)")
    return ap.parse_args()


_WC_KEEP = []


def pinned_host(shape, dtype, write_combined=False):
    """Pinned host tensor; write_combined=True allocates it with cudaHostAlloc(cudaHostAllocWriteCombined) through libcudart."""
    if not write_combined:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), ctypes.c_uint(0x04))
    if rc != 0:
        raise MemoryError(f"cudaHostAlloc(write-combined, {n} bytes) failed with {rc}")
    buf = (ctypes.c_char * n).from_address(ptr.value)
    _WC_KEEP.append((rt, ptr, buf))
    return torch.frombuffer(buf, dtype=dtype).view(shape)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 50` running beside the bench; samples are time-stamped on arrival so that the summary can be
    restricted to the timed region (the recipe's clocks line, /opt/skills/guides/B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self._proc, self._t = index, [], None, None

    def _reader(self):
        try:
            for line in self._proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 6:
                    self.samples.append((time.perf_counter(), f))
        except Exception:
            pass

    def start(self):
        try:
            self._proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                           "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self._t = threading.Thread(target=self._reader, daemon=True)
            self._t.start()
        except Exception:
            self._proc = None
        return self

    def stop(self):
        if self._proc is not None:
            try:
                self._proc.terminate()
                self._proc.wait(timeout=3)
            except Exception:
                pass

    def summary(self, t0: float, t1: float):
        inside = [f for (t, f) in self.samples if t0 <= t <= t1]
        scope = "timed region"
        if not inside:                      # region shorter than the sampling period: take the samples closest to it
            near = sorted(self.samples, key=lambda tf: min(abs(tf[0] - t0), abs(tf[0] - t1)))[:3]
            inside, scope = [f for (_, f) in near], "nearest samples (timed region shorter than the 50 ms sampling period)"
        num = lambda x: float(x) if x.replace(".", "", 1).isdigit() else None
        sm = [v for v in (num(f[0]) for f in inside) if v is not None]
        mx = [v for v in (num(f[1]) for f in inside) if v is not None]
        reasons = sorted({n for f in inside for n, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "scope": scope}


# ---------------------------------------------------------------------------------------------------------------
# the step (product path)
# ---------------------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, device, seed, n_tracks, n_frames, fused=True, st_native=False, overlap=False, aux=True):
        import sola_b200 as S
        from sola_b200 import synth
        self.fused = fused
        self.st_native = st_native
        # the tail of video k (R2, K2 gather, K2 N x N, labels, read-backs: integer / latency bound, a few hundred MB) runs on a
        # second, high-priority stream UNDER the HBM-bound K1+R1 of video k+1; every in-flight video has its own output buffers
        self.overlap = overlap and fused
        self.tail_stream = torch.cuda.Stream(device=device, priority=-1) if self.overlap else None
        # within a video: R2 runs beside K1+R1, K2 gather + label counts beside K2 N x N (they only share read-only inputs)
        self.aux_stream = torch.cuda.Stream(device=device) if aux else None
        self.S, self.device = S, device
        self.N, self.T, self.H, self.W = n_tracks, n_frames, CFG["H"], CFG["W"]
        self.logits, prompts = synth.dedup_candidates(self.N, self.T, self.H, self.W, seed=seed, device=device, bin_size=CFG["bin_size"])
        self.prompt_meta = [{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"]} for p in prompts]
        self.prompt_masks_host = np.stack([p["segmentation"] for p in prompts])                       # (N, H, W) uint8
        self.prompt_masks_dev = torch.from_numpy(self.prompt_masks_host).to(device)
        self.oh, self.ow = S.packed.default_target_shape(self.H, self.W)
        n_slots = 2 if (self.overlap or aux) else 1                                                            # reused outputs, one set per in-flight video
        self.slot_packed = [S.PackedMasks.empty((self.N, self.T), self.H, self.W, device) for _ in range(n_slots)]
        self.slot_counts = [torch.empty((3, self.N * self.T), dtype=torch.int32, device=device) for _ in range(n_slots)]
        self.slot_resized = [S.PackedMasks.empty((self.N, self.T), self.oh, self.ow, device) for _ in range(n_slots)]
        self.packed, self.counts = self.slot_packed[0], self.slot_counts[0]                           # the last finished step's outputs
        self.k1_events = []
        # GT masklets of the video at the resized shape (what `gt_masklets` holds, generate_tokens_grid.py:112): G objects x T frames
        self.gt_masks = torch.stack([synth.blob_masklet(self.T, self.oh, self.ow, seed * 31 + g, device=device, fill=0.12 + 0.05 * g)
                                     for g in range(CFG["n_gt"])])                                    # (G, T, oh, ow) uint8
        self.gt_planes = S.pack_masks(self.gt_masks)
        self.k1_bytes = self.N * self.T * (self.H * self.W * 4 + self.H * ((self.W + 31) // 32) * 4 + 12)          # logits in, planes + 3 counts out
        if fused:
            self.k1_bytes += self.N * self.T * self.oh * ((self.ow + 31) // 32) * 4                                 # + resized planes out

    def make_jobs(self):
        from sola_b200 import dedup
        mk = lambda: dedup.VideoDedupJob(self.prompt_meta, self.T, device=self.device, mode="grid", st_on_resized=not self.st_native, bin_size=CFG["bin_size"],
                                         n_max_tracks=CFG["n_max_tracks"], batch_size=CFG["batch_size"], miou_thresh=CFG["miou_thresh"],
                                         tail_stream=self.tail_stream, aux_stream=self.aux_stream)
        self.jobs = [mk(), mk()]
        for j in self.jobs:
            j.set_gt_masklets(self.gt_planes)

    def enqueue(self, slot, logits=None, prompt_masks=None, record_k1=False):
        """Device half of one step (no synchronisation): K1+R1, R2, K2-gather, K2 N x N, label counts, async read-back."""
        job = self.jobs[slot]
        logits = self.logits if logits is None else logits
        prompt_masks = self.prompt_masks_dev if prompt_masks is None else prompt_masks
        S = self.S
        b = slot % len(self.slot_packed)
        out_p, out_c, out_r = self.slot_packed[b], self.slot_counts[b], self.slot_resized[b]
        if self.aux_stream is not None:
            job.enqueue_prompts(prompt_masks, (self.oh, self.ow))                                         # R2, beside K1+R1
        # the dominant kernel is the first launch of the step: bracket it with events on the launching stream
        if record_k1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.fused:
            packed, counts, resized = S.binarize_pack_resize(logits, 0.0, 1.0, out=out_p, counts_out=out_c, resized_out=out_r)   # K1 + R1
        else:
            packed, counts = S.binarize_pack_stability(logits, 0.0, 1.0, out=out_p, counts_out=out_c)                        # K1
        if record_k1:
            e1.record()
            self.k1_events.append((e0, e1))
        if not self.fused:
            resized = S.resize_bilinear_bin(packed)                                                                        # R1
        job._enqueue_tail(packed, resized, counts, prompt_masks)                                         # R2, K2 gather, K2 N x N, labels, read-backs

    def finish(self, slot):
        """Host half: wait for that step's read-back, replay both greedy loops, stability scores, label metrics."""
        r = self.jobs[slot].finish()
        b = slot % len(self.slot_packed)
        self.packed, self.counts = self.slot_packed[b], self.slot_counts[b]
        return {"tracked": r["tracked"], "filtered": r["filtered"], "kept_st": r["kept_spatiotemporal"], "stability": r["stability"],
                "inter": r["inter"], "labels": r.get("labels")}

    def step(self, logits=None, prompt_masks=None):
        self.enqueue(0, logits, prompt_masks)
        return self.finish(0)

    def run_steps(self, n, record_k1=False):
        """n steps, software-pipelined: the host post-processes step k while the GPU runs step k+1."""
        out, prev = None, None
        for k in range(n):
            slot = k & 1
            self.enqueue(slot, record_k1=record_k1)
            if prev is not None:
                out = self.finish(prev)
            prev = slot
        if prev is not None:
            out = self.finish(prev)
        return out


def checks(w: Workload, out) -> dict:
    """Size-independent properties at full size + oracle comparison on a sub-sample (untimed)."""
    from oracle import maskpath_oracle as O
    inter = out["inter"]
    assert np.array_equal(inter, inter.T), "intersection matrix not symmetric"
    area = np.diag(inter)
    c = w.counts.view(3, w.N, w.T).cpu().numpy()
    if w.st_native:
        assert np.array_equal(c[1].sum(1, dtype=np.int64), area), "diag(inter) != sum of K1 areas (checksum of checksums)"
    else:
        rz, r_area = w.S.resize_bilinear_bin(w.packed, want_area=True)                 # untimed: R1's own per-frame areas
        r_area = r_area.view(w.N, w.T).cpu().numpy().sum(1, dtype=np.int64)
        assert np.array_equal(r_area, area), "diag(inter) != sum of R1 areas (checksum of checksums)"
    assert (c[0] <= c[1]).all() and (c[1] <= c[2]).all(), "stability counts not nested"
    assert (inter <= np.minimum(area[:, None], area[None, :])).all()
    # sub-sample vs the oracle: tracks 0, 1 and the last two x 6 frames spread over the video — planes and stability scores —, and the
    # label metrics of track 0 vs GT 0 over all frames (the GPU test suite holds the full-size entry-by-entry comparisons)
    ti = sorted({0, min(1, w.N - 1), max(w.N - 2, 0), w.N - 1})
    fi = sorted({0, 1, w.T // 3, w.T // 2, max(w.T - 2, 0), w.T - 1})
    sub = w.logits[ti][:, fi].float().cpu()
    planes = w.packed.words[ti][:, fi].cpu().numpy().view(np.uint32)
    assert np.array_equal(planes, O.pack_bits(sub.numpy() > 0)), "K1 planes differ from the oracle on the sub-sample"
    s = O.get_stability_score(sub.numpy())
    assert np.array_equal(np.nan_to_num(s, nan=-1), np.nan_to_num(out["stability"][ti][:, fi], nan=-1)), "stability differs"
    lab = out["labels"]
    rz0 = w.S.unpack_masks(w.S.resize_bilinear_bin(w.packed[0]), torch.float32).cpu()
    p_, r_, i_ = O.compute_mask_metrics(rz0, w.gt_masks[0].float().cpu())
    assert (float(p_), float(r_), float(i_)) == (float(lab["precision"][0, 0]), float(lab["recall"][0, 0]), float(lab["iou"][0, 0])), \
        "label metrics differ from the oracle's compute_mask_metrics"
    return {"tracked": len(out["tracked"]), "filtered": len(out["filtered"]), "kept_spatiotemporal": len(out["kept_st"]),
            "labels": f"{lab['iou'].shape[0]} tracks x {lab['iou'].shape[1]} GT objects, mean IoU {float(lab['iou'].mean()):.4f}"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's own CPU path, a FULL config-2 step timed once
# ---------------------------------------------------------------------------------------------------------------
class _Clock:
    def __init__(self):
        self.t = {}

    def timed(self, key, fn):
        def wrapped(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            self.t[key] = self.t.get(key, 0.0) + time.perf_counter() - t0
            return r
        return wrapped


def cpu_reference_inputs(n_tracks, n_frames, seed=1234 + 2, gen_device=None):
    """Synthetic SAM2 logits + prompts + GT masklets on the HOST (generation is never timed; it may run on the GPU when there is one)."""
    from oracle import maskpath_oracle as O
    from sola_b200 import synth
    H, W = CFG["H"], CFG["W"]
    dev = gen_device if gen_device is not None else "cpu"
    logits, prompts = synth.dedup_candidates(n_tracks, n_frames, H, W, seed=seed, device=dev, bin_size=CFG["bin_size"])
    if logits.is_cuda:
        host = torch.empty(logits.shape, dtype=logits.dtype)
        for i in range(0, n_tracks, 8):
            host[i:i + 8] = logits[i:i + 8].cpu()
        logits = host
        torch.cuda.empty_cache()
    oh, ow = O.default_target_shape(H, W)
    gts = [synth.blob_masklet(n_frames, oh, ow, seed * 31 + g, device="cpu", fill=0.12 + 0.05 * g).float() for g in range(CFG["n_gt"])]
    return logits, prompts, gts


def cpu_reference_step(logits, prompts, gts, st_pair_sample=0):
    """The reference's grid loop (generate_tokens_grid.py:133-292 via oracle/greedy_oracle.py) with every dense operation the loop runs
    per tracked batch: per-frame binarise + cat (:215-224), stability score per plane (prompt_generator.py:169-186; named by BASELINE
    config 2), reshape_masklet (:248-250), the label metrics against every GT object (:253-264) and the greedy suppression pairs
    (:266-278: nearest resize + compute_mask_iou).  Returns (seconds, seconds per stage, seconds per dead-code volume pair, result)."""
    from oracle import greedy_oracle as GO
    from oracle import maskpath_oracle as O
    n_frames = int(logits.shape[1])
    clk = _Clock()

    def binarise_and_stability(frame_idx, batch):
        out = {}
        for p in batch:
            i = p["prompt_id"]
            frames = [O.binarize(logits[i, t][None]) for t in range(n_frames)]                 # :219 per frame
            out[i] = torch.cat(frames, 0)                                                      # :224
            _ = [O.get_stability_score(logits[i, t].numpy()) for t in range(n_frames)]         # prompt_generator.py:169 per plane
        return out

    def reshape_and_label(m):
        r = clk.timed("reshape_masklet", O.reshape_masklet)(m)
        for g in gts:                                                                          # :256-264
            clk.timed("label_metrics", O.compute_mask_metrics)(r, g)
        return r

    class Impl:
        compute_mask_iou = staticmethod(clk.timed("greedy_pairs", O.compute_mask_iou))
        reshape_masklet = staticmethod(reshape_and_label)

    t0 = time.perf_counter()
    res = GO.grid_greedy([dict(p) for p in prompts], n_frames, clk.timed("binarise_cat_stability", binarise_and_stability),
                         bin_size=CFG["bin_size"], n_max_tracks=CFG["n_max_tracks"], batch_size=CFG["batch_size"], miou_thresh=CFG["miou_thresh"], impl=Impl)
    t_step = time.perf_counter() - t0
    t_pair = None
    if st_pair_sample:
        # the dead-code volume IoU (seg_utils.compute_masklet_iou has no caller in the reference): timed on a few pairs, reported apart
        a = O.reshape_masklet((logits[0] > 0).float())
        b = O.reshape_masklet((logits[1] > 0).float())
        t0 = time.perf_counter()
        for _ in range(st_pair_sample):
            O.compute_masklet_iou(a, b, "cpu")
        t_pair = (time.perf_counter() - t0) / st_pair_sample
    stages = dict(clk.t)
    stages["greedy_other (nearest resize, control flow)"] = max(0.0, t_step - sum(clk.t.values()))
    return t_step, stages, t_pair, res


REF_MAX_STEPS = 3


def reference_line(args, n_tracks, n_frames, workload):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cores = torch.get_num_threads()
    gen = torch.device("cuda", 0) if torch.cuda.is_available() else None
    logits, prompts, gts = cpu_reference_inputs(n_tracks, n_frames, gen_device=gen)
    # warm-up: the thread pool / allocator on a 2-track slice (untimed); then up to REF_MAX_STEPS FULL steps, each timed
    cpu_reference_step(logits[:2, : min(n_frames, 8)], [{**p, "frame_idx": 0} for p in prompts[:2]], [g[:8] for g in gts])
    n_steps = max(1, min(args.steps, REF_MAX_STEPS))
    runs = []
    for k in range(n_steps):
        t_step, stages, t_pair, res = cpu_reference_step(logits, prompts, gts, st_pair_sample=6 if k == 0 else 0)
        runs.append((t_step, stages))
        if k == 0:
            pair_s = t_pair
    t_mean = float(np.mean([r[0] for r in runs]))
    stages = {k: float(np.mean([r[1][k] for r in runs])) for k in runs[0][1]}
    n_pairs = n_tracks * (n_tracks - 1) // 2
    value = n_tracks * n_frames / t_mean
    with_pairs = n_tracks * n_frames / (t_mean + n_pairs * pair_s)
    sample = (f"FULL steps, each timed: {n_tracks} candidate masklets x {n_frames} frames x {CFG['H']}x{CFG['W']} through the reference grid loop "
              f"({len(res['tracked'])} tracked by SAM2, {len(res['filtered'])} filtered before tracking); step times (s): "
              + ", ".join(f"{r[0]:.2f}" for r in runs) + "; mean stages (s): " + ", ".join(f"{k} {v:.2f}" for k, v in stages.items()))
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps,
            "warmup": 1, "ms_per_step": t_mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload,
                       "note": "oracle port of the reference's CPU path (the reference is Python; /root/reference does not exist on the GPU box), "
                               f"torch-CPU / numpy ops on all host threads; FULL steps are timed, --steps honoured as min(steps, {REF_MAX_STEPS}) and "
                               "--warmup as one untimed 2-track run, because a step takes seconds; one rank regardless of --gpus. The reference "
                               "tracks (and therefore binarises / resizes / labels) only the candidates that survive the greedy filter; the unit "
                               "count is the video's candidate masklet-frames, as in the product arm. The headline excludes the N x N "
                               "compute_masklet_iou pairs, which are dead code in the reference (SURVEY.md §0); `with_dead_code_st_pairs` adds "
                               "them (per-pair time measured on a sample)",
                       "requested_steps": args.steps, "requested_warmup": args.warmup},
            "step_s": [r[0] for r in runs], "stage_s": stages,
            "with_dead_code_st_pairs": {"value": with_pairs, "unit": UNIT, "pairs": n_pairs, "s_per_pair": pair_s,
                                        "extrapolated": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def cpu_baseline_sample(n_tracks, n_frames):
    """The product run's `cpu_baseline` leg: the same reference step (full 64-candidate loop, ~6 s on 16 cores), timed once after a
    short warm-up."""
    logits, prompts, gts = cpu_reference_inputs(n_tracks, n_frames, gen_device=torch.device("cuda", torch.cuda.current_device()))
    cpu_reference_step(logits[:2, : min(n_frames, 8)], [{**p, "frame_idx": 0} for p in prompts[:2]], [g[:8] for g in gts])
    t_step, stages, _, res = cpu_reference_step(logits, prompts, gts)
    value = n_tracks * n_frames / t_step
    sample = (f"one FULL config-2 step: {n_tracks} candidate masklets x {n_frames} frames x {CFG['H']}x{CFG['W']} through the reference grid loop "
              f"({len(res['tracked'])} tracked, {len(res['filtered'])} filtered) = {t_step:.1f} s; stages (s): "
              + ", ".join(f"{k} {v:.2f}" for k, v in stages.items()))
    return value, sample


# ---------------------------------------------------------------------------------------------------------------
# second timed region: config-4-shaped J&F sweep through the fused kernel
# ---------------------------------------------------------------------------------------------------------------
def jf_region(device, rank, world, reps, peak):
    """Two passes over the same config-4-shaped sweep: (1) J + F exactly as the reference defines them (evaluator.py:227-247: region
    counts only — HBM bound), (2) the same plus the north-star's boundary F (seg2bmap + disk dilation + match counting — integer /
    shared-memory bound).  Each pass is ONE launch of the fused kernel per sweep."""
    import torch.distributed as dist
    import sola_b200 as S
    from oracle import boundary_oracle as BO
    from oracle import maskpath_oracle as O
    from sola_b200 import evaluator, packed as P, sharding, synth
    units = synth.mevis_like_sweep(JF_SWEEP["n_videos"], JF_SWEEP["exprs_per_video"], 1234 + 4 + 1000 * rank, device,
                                   t_range=JF_SWEEP["t_range"], pack=S.pack_masks)
    pairs = [(p, g) for _, _, p, g in units]
    k_chk = min(range(len(pairs)), key=lambda i: pairs[i][0].words.numel())
    pu, gu = (S.unpack_masks(x, torch.uint8).cpu().numpy() for x in pairs[k_chk])
    stage, roofs = {}, {}
    for mode, with_boundary in (("J+F (reference definitions)", False), ("J+F+boundary-F", True)):
        plan = P.JFSweepPlan(pairs, with_boundary=with_boundary)
        buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device=device)
        host = torch.empty((7, plan.total_frames), dtype=torch.int32).pin_memory()

        def sweep_once():
            """ONE launch for the whole sweep, one read-back, the reference's formulas per unit on the host, the integer audit."""
            plan.run(buf)
            host.copy_(buf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            J, F, Fb, tot = evaluator.sweep_metrics_from_counts(host.numpy(), plan.offsets, plan.frames, with_boundary)
            return J.tolist(), F.tolist(), (Fb.tolist() if with_boundary else []), tot

        for _ in range(3):
            Js, Fs, Fbs, tot = sweep_once()
        # parity inside the run: one unit against the oracles (untimed)
        assert Js[k_chk] == float(O.compute_J(torch.from_numpy(pu).float(), torch.from_numpy(gu).float())), "J differs from the oracle on the checked unit"
        assert abs(Fs[k_chk] - O.compute_F(torch.from_numpy(pu).float(), torch.from_numpy(gu).float())) < 1e-6, "F differs from the oracle"
        if with_boundary:
            assert abs(Fbs[k_chk] - BO.boundary_f_masklet(pu, gu)) < 1e-12, "boundary F differs from the oracle on the checked unit"
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # (a) kernel only, CUDA events on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.run(buf)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        # (b) the whole sweep incl. read-back, host formulas and (N > 1) the final NCCL sum of the accumulators
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_sw = max(3, reps // 4)
        for _ in range(n_sw):
            Js, Fs, Fbs, tot = sweep_once()
            red = sharding.allreduce_jf(sum(Js), sum(Fs), sum(0.5 * (a + b) for a, b in zip(Js, Fs)), len(Js), tot, device=device)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / n_sw, k_ms * 1e-3], dtype=torch.float64, device=device)
        frames = torch.tensor([float(plan.total_frames)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(frames, op=dist.ReduceOp.SUM)
        sweep_s, kern_s, total_frames = float(dt[0]), float(dt[1]), float(frames[0])
        ach = plan.algorithmic_bytes / (k_ms * 1e-3) / 1e9
        stage[mode] = {"masklet_frames_per_s": total_frames / sweep_s, "kernel_only_masklet_frames_per_s": total_frames / kern_s,
                       "ms_per_sweep": sweep_s * 1e3, "kernel_ms": kern_s * 1e3, "launches_per_sweep": len(plan.launches),
                       "mean_J": red["mean_J"], "mean_F": red["mean_F"], "int_totals": [int(x) for x in red["int_totals"]]}
        if with_boundary:
            stage[mode]["mean_F_boundary_rank0"] = float(np.mean(Fbs))
        roofs[mode] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                       "bytes_per_launch": int(plan.algorithmic_bytes), "ms_per_launch": k_ms, "work_items": int(plan.n_items)}
    stage["workload"] = (f"config4-shaped J&F sweep: {len(pairs)} (video, expression) units per GPU, {plan.total_frames} frame pairs per GPU, mixed "
                         f"360p-1080p, {JF_SWEEP['t_range'][0]}-{JF_SWEEP['t_range'][1]} frames, bit-packed in HBM, {world} GPU(s)")
    stage["timed"] = "kernel: CUDA events on the launching stream; sweep: host clock around launch + read-back + host formulas" + \
                     (" + NCCL all-reduce of the accumulators" if world > 1 else "")
    stage["parity"] = "one unit checked against oracle.compute_J / compute_F / boundary_oracle inside the run (untimed)"
    r1 = roofs["J+F (reference definitions)"]
    r1["kernel"] = "jf_fused_kernel, region mode: J and F as evaluator.py:227-247 defines them (TMA-staged tiles, popcounts)"
    r1["note"] = "algorithmic bytes = both bit-packed planes of every frame pair read once"
    r2 = roofs["J+F+boundary-F"]
    r2["kernel"] = "jf_fused_kernel, boundary mode: + seg2bmap, sparse disk dilation and match counting in shared memory (north-star kernel 3)"
    r2["note"] = ("same algorithmic bytes; this mode is bound by instruction issue / shared-memory wavefronts, not HBM (ncu: profiles/), so the HBM "
                  "fraction is the distance to the floor the region mode reaches, not a utilisation claim")
    return stage, r1, r2


def cfg5_region(device, rank, world):
    """BASELINE config 5 slice (N > 1): 32 tracks per GPU x 200 frames of 1080p-derived 540x960 planes; the N x N matrix through the
    one-kernel exchange + K2 (TMA loads over peer memory), the peer-memory pull pipelined with K2 and the NCCL all-to-all, all checked
    against the single-rank matrix on rank 0."""
    import torch.distributed as dist
    import sola_b200 as S
    from sola_b200 import packed as P, sharding, synth
    n_local, T, H, W = 32, 200, 540, 960
    peers = sharding.PeerPlanes(n_local, T, H, W, device)
    for i in range(n_local):
        gidx = rank * n_local + i
        m = synth.blob_masklet(T, H, W, 500 + gidx % max(1, (n_local * world) // 3), device=device, fill=0.10 + 0.02 * (gidx % 5))
        peers.local.words[i] = S.pack_masks(m).words
    torch.cuda.synchronize()
    dist.barrier()

    def timed(fn, reps=3):
        best, out = None, None
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return best, out
    tma_ms, m_tma = timed(lambda: peers.pairwise_inter_matrix("tma"))
    pull_ms, m_pull = timed(lambda: peers.pairwise_inter_matrix("pull"))
    a2a_ms, m_a2a = timed(lambda: sharding.pairwise_inter_matrix_sharded(peers.local, split="words"))
    gathered = torch.empty((n_local * world, T, H, peers.local.Wp), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, peers.local.words.contiguous())
    identical = None
    if rank == 0:
        alone = S.pairwise_inter_matrix(S.PackedMasks(gathered, H, W))
        identical = bool(torch.equal(alone, m_tma) and torch.equal(alone, m_pull) and torch.equal(alone, m_a2a))
    words = T * H * peers.local.Wp
    remote_bytes = (world - 1) * n_local * (words // world) * 4            # what one rank pulls over NVLink
    return {"workload": f"config5-shaped slice: {n_local * world} tracks x {T} frames x {H}x{W} planes, {n_local} tracks per GPU",
            "one_kernel_peer_tma_ms": tma_ms, "peer_pull_pipelined_ms": pull_ms, "nccl_all_to_all_ms": a2a_ms,
            "identical_to_single_rank": identical, "nvlink_GBps_per_rank_pull": remote_bytes / (pull_ms * 1e-3) / 1e9,
            "note": "exchange + K2 over this rank's word slice + all-reduce of the int64 matrix, max over ranks, best of 3"}


# ---------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_tracks, n_frames = args.tracks, args.frames
    workload = f"config2: grid-prompt track dedup + label metrics, {n_tracks} masklets x {n_frames} frames x {CFG['H']}x{CFG['W']} fp32 logits, " \
               f"stability score, miou_thresh {CFG['miou_thresh']}, {CFG['n_gt']} GT masklets (one video per GPU per step)"

    if args.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(reference_line(args, n_tracks, n_frames, workload)))
        return

    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=device)
    import sola_b200 as S
    S.load_library()
    from sola_b200 import sharding
    # N > 1: one process per GPU, each bound to its GPU's local CPUs before any pinned allocation (the N = 1 run keeps every core:
    # its cpu_baseline leg must see the whole host)
    cpus = sharding.bind_to_gpu_cpus(local_rank) if (world > 1 and not os.environ.get("SOLA_BENCH_NO_AFFINITY")) else []

    w = Workload(device, seed=1234 + 2 + 1000 * rank, n_tracks=n_tracks, n_frames=n_frames, fused=not args.no_fuse, st_native=args.st_native,
                 overlap=args.overlap, aux=not args.no_aux)
    w.make_jobs()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local_rank).start()
    out = w.run_steps(max(args.warmup, 3))
    info = checks(w, out)
    barrier()
    launches0 = S.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")
    t_wall0 = time.perf_counter()
    ev0.record()
    for j in w.jobs:
        j.timing_events = []
    out = w.run_steps(args.steps, record_k1=True)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    torch.cuda.nvtx.range_pop()
    time.sleep(0.12)                         # let the last in-flight sample arrive
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = S.launch_count() - launches0
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in w.k1_events]))
    k2_ms = float(np.mean([a.elapsed_time(b) for j in w.jobs for a, b in (j.timing_events or [])]))
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    units = world * n_tracks * n_frames * args.steps
    value = units / (max_ms * 1e-3)
    clocks = clk.summary(t_wall0, t_wall1)

    # ---- e2e: logits + prompt masks + GT masks start in pinned host memory; H2D in the timed region; results read back -------
    e2e = None
    if not args.no_e2e:
        try:
            need = w.logits.numel() * 4 * world
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = None
            if avail is not None and need > 0.7 * avail:
                raise MemoryError(f"pinned staging for {world} ranks needs {need / 1e9:.0f} GB, host has {avail / 1e9:.0f} GB available")
            chunk = 4
            host_chunks = []
            for s in range(0, n_tracks, chunk):
                h = pinned_host((min(chunk, n_tracks - s), n_frames, CFG["H"], CFG["W"]), torch.float32, args.e2e_wc)
                h.copy_(w.logits[s:s + chunk])
                host_chunks.append(h)
            host_prompts = torch.from_numpy(w.prompt_masks_host).pin_memory()
            host_gt = w.gt_masks.cpu().pin_memory()
            dev_logits = torch.empty_like(w.logits)
            copy_stream = torch.cuda.Stream(device=device)
            h2d_ms = []

            def e2e_step():
                # chunked H2D on a copy stream; K1 of chunk c runs while chunk c+1 is still crossing PCIe
                evs = []
                with torch.cuda.stream(copy_stream):
                    c0 = torch.cuda.Event(enable_timing=True)
                    c0.record(copy_stream)
                    for c, h in enumerate(host_chunks):
                        dev_logits[c * chunk: c * chunk + h.shape[0]].copy_(h, non_blocking=True)
                        e = torch.cuda.Event(enable_timing=(c == len(host_chunks) - 1))
                        e.record(copy_stream)
                        evs.append(e)
                    h2d_ms.append((c0, evs[-1]))
                    pm = host_prompts.to(device, non_blocking=True)
                    gt = host_gt.to(device, non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(copy_stream)
                    evs.append(e)
                cview = w.counts.view(3, n_tracks, n_frames)
                for c, h in enumerate(host_chunks):
                    torch.cuda.current_stream().wait_event(evs[c])
                    sl = slice(c * chunk, c * chunk + h.shape[0])
                    _, cc = S.binarize_pack_stability(dev_logits[sl], 0.0, 1.0, out=w.packed[sl])       # K1 on the chunk that just landed
                    cview[:, sl] = cc
                torch.cuda.current_stream().wait_event(evs[-1])
                w.jobs[0].set_gt_masklets(S.pack_masks(gt))                                              # GT masks packed on the device
                w.jobs[0].enqueue_after_k1(w.packed, w.counts.view(3, n_tracks, n_frames), pm)          # R1, R2, K2-gather, K2, labels, read-backs
                return w.finish(0)

            e2e_step()
            barrier()
            n_e2e = max(1, min(args.e2e_steps, args.steps))
            h2d_ms.clear()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_e2e):
                res = e2e_step()
            b.record()
            barrier()
            te = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            # per-rank H2D bandwidth of the logits (explains the N > 1 e2e numbers: all ranks stream from one host's memory)
            bw = torch.tensor([w.logits.numel() * 4 / (np.mean([x.elapsed_time(y) for x, y in h2d_ms]) * 1e-3) / 1e9], dtype=torch.float64, device=device)
            bw_all = [torch.zeros_like(bw) for _ in range(world)] if world > 1 else [bw]
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
                dist.all_gather(bw_all, bw)
            h2d = w.logits.numel() * 4 + host_prompts.numel() + host_gt.numel()
            G = CFG["n_gt"]
            d2h = int(3 * n_tracks * n_frames * 4 + n_tracks * n_tracks * 8 + 3 * n_tracks * n_tracks * 4
                      + (n_tracks * G * n_frames + n_tracks * n_frames + G * n_frames) * 4)
            e2e = {"value": world * n_tracks * n_frames * n_e2e / (float(te.item()) * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h, "steps": n_e2e, "host_cpus_bound": len(cpus),
                   "h2d_GBps_per_rank": [round(float(x), 1) for x in bw_all], "host_staging": "write-combined pinned" if args.e2e_wc else "pinned",
                   "note": "pinned host fp32 logits + uint8 prompt masks + uint8 GT masks copied H2D every step (chunked, overlapped with K1); "
                           "stability counts, IoU count matrices, the N x N intersection matrix and the label count tables read back"}
            for j in w.jobs:
                j.set_gt_masklets(w.gt_planes)
            del host_chunks, dev_logits
        except Exception as ex:  # e.g. pinned allocation refused on a small host
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:200]}

    peak, peak_src = peaks()
    jf_stage = roofline_jf = roofline_jfb = cfg5 = None
    launches_jf = 0
    if not args.no_jf:
        try:
            l0 = S.launch_count()
            jf_stage, roofline_jf, roofline_jfb = jf_region(device, rank, world, args.jf_reps, peak)
            launches_jf = S.launch_count() - l0
        except Exception as ex:
            jf_stage = {"error": repr(ex)[:300]}
    if world > 1 and not args.no_cfg5:
        try:
            cfg5 = cfg5_region(device, rank, world)
        except Exception as ex:
            cfg5 = {"error": repr(ex)[:300]}
    clk.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fused_flag = w.fused
    achieved = w.k1_bytes / (k1_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload, "l2_policy": "inputs larger than L2 (18.9 GB fp32 logits per step vs 126 MB L2)",
                   "kept_sets": info, "parallelism": f"video-sharded x{world}, no data-path collective",
                   "n_x_n_planes": "native 720x1280" if args.st_native else "540x960 resized masklets (what the reference filter compares)"},
        "clocks": clocks,
        "stage_ms": ({"dominant (K1+R1 fused), timed while the previous video's tail runs beside it": k1_ms,
                      "K2 N x N (int pipe, carry-save, TMA-staged), on the tail stream under the next video's K1+R1": k2_ms,
                      "step time not covered by K1+R1 (tail not hidden + gaps)": max_ms / args.steps - k1_ms,
                      "streams": "2 (--overlap: tail of video k under K1+R1 of video k+1)"} if w.overlap else
                     {"dominant (K1+R1 fused)" if w.fused else "dominant (K1)": k1_ms, "K2 N x N (int pipe, carry-save, TMA-staged)": k2_ms,
                      "everything else incl. label counts and gaps": max_ms / args.steps - k1_ms - k2_ms,
                      "streams": ("2: R2 runs beside K1+R1; K2 gather, label counts and read-backs beside K2 N x N (both kernel times are "
                                  "measured with those neighbours running)") if w.aux_stream is not None else "1"}),
        "gpu_launches": int(launches),
        "gpu_launches_jf_region": int(launches_jf),
        "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": ("fused_pack_resize_kernel<float> (K1+R1: binarise+pack+stability+resize)" if w.fused
                                                else "pack_flat_kernel<float, THRESH3> (K1 binarise+pack+stability)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "bytes_per_launch": w.k1_bytes, "ms_per_launch": k1_ms, "share_of_step": k1_ms / (max_ms / args.steps),
                     "traffic": None},
    }
    if jf_stage is not None:
        line["jf_stage"] = jf_stage
    if roofline_jf is not None:
        roofline_jf["peak_source"] = roofline_jfb["peak_source"] = peak_src
        line["roofline_jf"] = roofline_jf
        line["roofline_jf_boundary"] = roofline_jfb
    if cfg5 is not None:
        line["cfg5"] = cfg5
    try:
        # secondary roofline: K2 is integer-pipe work.  Denominator = alu pipe at 64 lanes/clk/SM (B300_MICROARCH.md; LOP3 measured 64-84
        # by tools/microbench_int.cu) x 148 SMs x the SM clock sampled during the timed region / 2.5 LOP3 per pair-word (4 AND + 6
        # compressor LOP3 per 4 words).  `frac` counts pair-words DENSELY; the kernel skips all-zero operand quads, so on the object-like
        # bench masklets it is an upper bound on pipe utilisation — `executed_frac_of_dense` says how much of the dense work was run.
        oh_, ow_ = (CFG["H"], CFG["W"]) if args.st_native else (w.oh, w.ow)
        words_ = n_frames * oh_ * ((ow_ + 31) // 32)
        pair_words = (n_tracks * (n_tracks + 1) // 2) * words_
        sm_mhz = (line["clocks"] or {}).get("sm_mhz") or 1965.0
        if not (k2_ms > 0):
            raise ValueError("no K2 timing events")
        peak_k2 = 148 * 64 * sm_mhz * 1e6 / 2.5
        ach_k2 = pair_words / (k2_ms * 1e-3)
        planes_ = (w.packed if args.st_native else S.resize_bilinear_bin(w.packed)).words.view(n_tracks, -1, 4)
        nz_quads = (planes_ != 0).any(dim=-1).float().mean().item()            # share of (row, k-quad) operands that are not skipped
        line["roofline_k2"] = {"bound": "int-alu", "kernel": "pair_iou_st_ring_kernel (K2 N x N AND-popcount, carry-save)",
                               "achieved": ach_k2 / 1e12, "peak": peak_k2 / 1e12, "unit": "T pair-words/s", "frac": ach_k2 / peak_k2,
                               "executed_frac_of_dense": nz_quads, "frac_on_executed_pair_words": ach_k2 * nz_quads / peak_k2,
                               "peak_source": "148 SMs x 64 alu lanes/clk x sampled SM clock / 2.5 LOP3 per pair-word",
                               "pair_words_per_launch": int(pair_words), "ms_per_launch": k2_ms}
    except Exception as ex:
        line["roofline_k2"] = {"error": repr(ex)[:200]}
    if not args.no_cpu_baseline and world == 1:
        del w
        torch.cuda.empty_cache()
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        v, sample = cpu_baseline_sample(n_tracks, n_frames)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample}
    traffic_path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")       # from the committed `ncu --set full` capture
    if os.path.isfile(traffic_path) and (n_tracks, n_frames) == (CFG["n_tracks"], CFG["n_frames"]):
        with open(traffic_path) as f:
            tj = json.load(f)
            key = "fused_pack_resize_kernel<float>" if fused_flag else "pack_flat_kernel<float, THRESH3>"
            line["roofline"]["traffic"] = tj.get(key, {}).get("dram_bytes_per_launch")
            if "roofline_jf" in line:       # the ncu target (tools/ncu_targets.py jf_region / jf_boundary) runs rank 0's sweep of this bench
                line["roofline_jf"]["traffic"] = tj.get("jf_fused_kernel", {}).get("dram_bytes_per_launch")
                line["roofline_jf_boundary"]["traffic"] = tj.get("jf_fused_kernel[boundary]", {}).get("dram_bytes_per_launch")
    ncu_path = os.path.join(ROOT, "profiles", "r3_ncu_summary.json")                       # committed `ncu --set full` summaries
    if os.path.isfile(ncu_path) and "roofline_jf_boundary" in line:
        with open(ncu_path) as f:
            cap = json.load(f).get("jf_boundary", {})
        line["roofline_jf_boundary"]["ncu"] = {k: cap.get(k) for k in ("issue_active_pct", "smem_wavefronts_pct", "pipe_alu_pct", "pipe_xu_pct",
                                                                         "warp_instructions", "registers", "stalls_per_issue")}
        line["roofline_jf_boundary"]["binding"] = "instruction issue + shared-memory wavefronts (see ncu block); HBM is the floor, not the bound"
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
