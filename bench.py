#!/usr/bin/env python
"""bench.py — masklet-frames/s of the masklet-scoring hot path on BASELINE.json config 2
("Grid-prompt track dedup: 64 synthetic SAM2 masklets x 80 frames at 720x1280 with stability score and miou_thresh 0.7").

One step = one video's pass over the path, inputs already in HBM:
    K1  binarise + bit-pack + stability counts over 64 x 80 fp32 logit planes      (dominant, HBM bound)
    R1  bilinear resize to 540 x 960 + > 0.5 of all 5120 packed planes             (seg_utils.reshape_masklet)
    R2  nearest resize + pack of the 64 prompt masks
    G1  reference-order greedy filter: K2-gather IoU per tracked batch + host suppression
    K2  64 x 64 spatio-temporal intersection matrix of the resized masklets + index-order greedy (compute_masklet_iou semantics)
    one D2H of the stability counts -> float64 scores
`value` = masklet-frames processed by all ranks / max-over-ranks device time (weak scaling: one video per GPU per step, no
data-path collective).  `e2e` = the same step with the logits and prompt masks starting in pinned HOST memory (H2D inside the
timed region, results read back).  `--impl reference` times the oracle port of the reference's CPU path on the host cores.
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
CFG = dict(n_tracks=64, n_frames=80, H=720, W=1280, miou_thresh=0.7, n_max_tracks=64, batch_size=4, bin_size=4)


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
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2], help="videos in flight on separate CUDA streams")
    ap.add_argument("--tail-priority", action="store_true",
                    help="run everything after the fused K1+R1 (R2, K2 gather, K2 N x N, read-backs) on a HIGH-priority side stream, so the "
                         "INT-bound K2 of video k shares the SMs with the HBM-bound K1+R1 of video k+1")
    ap.add_argument("--no-fuse", action="store_true", help="run K1 and R1 as two kernels instead of the fused one")
    ap.add_argument("--st-native", action="store_true",
                    help="N x N spatio-temporal IoU on the native 720x1280 planes instead of the 540x960 resized masklets the "
                         "reference's filter works on (generate_tokens_grid.py:248-250)")
    return ap.parse_args()


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
    def __init__(self, device, seed, n_tracks, n_frames, n_streams=1, fused=True, st_native=False, tail_priority=False):
        import sola_b200 as S
        self.n_streams = n_streams
        self.tail_stream = torch.cuda.Stream(device=device, priority=-1) if tail_priority else None
        self.fused = fused
        self.st_native = st_native
        from sola_b200 import synth
        self.S, self.device = S, device
        self.N, self.T, self.H, self.W = n_tracks, n_frames, CFG["H"], CFG["W"]
        self.logits, prompts = synth.dedup_candidates(self.N, self.T, self.H, self.W, seed=seed, device=device, bin_size=CFG["bin_size"])
        self.prompt_meta = [{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"]} for p in prompts]
        self.prompt_masks_host = np.stack([p["segmentation"] for p in prompts])                       # (N, H, W) uint8
        self.prompt_masks_dev = torch.from_numpy(self.prompt_masks_host).to(device)
        self.packed = S.PackedMasks.empty((self.N, self.T), self.H, self.W, device)                   # reused outputs
        self.counts = torch.empty((3, self.N * self.T), dtype=torch.int32, device=device)
        self.k1_events = []
        oh, ow = S.packed.default_target_shape(self.H, self.W)
        self.k1_bytes = self.N * self.T * (self.H * self.W * 4 + self.H * ((self.W + 31) // 32) * 4 + 12)          # logits in, planes + 3 counts out
        if fused:
            self.k1_bytes += self.N * self.T * oh * ((ow + 31) // 32) * 4                                           # + resized planes out

    def make_jobs(self):
        from sola_b200 import dedup
        mk = lambda: dedup.VideoDedupJob(self.prompt_meta, self.T, device=self.device, mode="grid", st_on_resized=not self.st_native, bin_size=CFG["bin_size"],
                                         n_max_tracks=CFG["n_max_tracks"], batch_size=CFG["batch_size"], miou_thresh=CFG["miou_thresh"])
        self.jobs = [mk(), mk()]
        # one stream per in-flight video: the HBM-bound K1 of video k+1 overlaps the ALU/XU-bound R1 + K2 of video k
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.jobs] if self.n_streams > 1 else [None, None]
        two = self.n_streams > 1 or self.tail_stream is not None              # two videos in flight need two sets of output buffers
        self.packed_slots = [self.packed, self.S.PackedMasks.empty((self.N, self.T), self.H, self.W, self.device)] if two else [self.packed, self.packed]
        self.counts_slots = [self.counts, torch.empty_like(self.counts)] if two else [self.counts, self.counts]

    def enqueue(self, slot, logits=None, prompt_masks=None, record_k1=False):
        """Device half of one step (no synchronisation): K1, R1, R2, K2-gather, K2 N x N, async read-back."""
        job = self.jobs[slot]
        logits = self.logits if logits is None else logits
        prompt_masks = self.prompt_masks_dev if prompt_masks is None else prompt_masks
        if self.streams[slot] is not None:
            with torch.cuda.stream(self.streams[slot]):
                self._enqueue(job, slot, logits, prompt_masks, record_k1)
        else:
            self._enqueue(job, slot, logits, prompt_masks, record_k1)

    def _enqueue(self, job, slot, logits, prompt_masks, record_k1):
        packed_out, counts_out = self.packed_slots[slot], self.counts_slots[slot]
        S = self.S
        # the dominant kernel is the first launch of the step: bracket it with events on the launching stream
        if record_k1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.fused:
            packed, counts, resized = S.binarize_pack_resize(logits, 0.0, 1.0, out=packed_out, counts_out=counts_out)      # K1 + R1
        else:
            packed, counts = S.binarize_pack_stability(logits, 0.0, 1.0, out=packed_out, counts_out=counts_out)            # K1
        if record_k1:
            e1.record()
            self.k1_events.append((e0, e1))
        if not self.fused:
            resized = S.resize_bilinear_bin(packed)                                                                        # R1
        if self.tail_stream is not None:
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(self.tail_stream):
                self.tail_stream.wait_event(done)
                job._enqueue_tail(packed, resized, counts, prompt_masks)
            return
        job._enqueue_tail(packed, resized, counts, prompt_masks)                                                           # R2, K2 gather, K2 N x N, read-backs

    def finish(self, slot):
        """Host half: wait for that step's read-back, replay both greedy filters, stability scores."""
        r = self.jobs[slot].finish()
        return {"tracked": r["tracked"], "filtered": r["filtered"], "kept_st": r["kept_spatiotemporal"], "stability": r["stability"],
                "inter": r["inter"]}

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
        _, r_area = w.S.resize_bilinear_bin(w.packed, want_area=True)                  # untimed: R1's own per-frame areas
        r_area = r_area.view(w.N, w.T).cpu().numpy().sum(1, dtype=np.int64)
        assert np.array_equal(r_area, area), "diag(inter) != sum of R1 areas (checksum of checksums)"
    assert (c[0] <= c[1]).all() and (c[1] <= c[2]).all(), "stability counts not nested"
    assert (inter <= np.minimum(area[:, None], area[None, :])).all()
    # sub-sample vs the oracle: 2 tracks x 3 frames of planes + their pair intersection
    sub = w.logits[:2, :3].float().cpu()
    planes = w.packed.words[:2, :3].cpu().numpy().view(np.uint32)
    assert np.array_equal(planes, O.pack_bits(sub.numpy() > 0)), "K1 planes differ from the oracle on the sub-sample"
    s = O.get_stability_score(sub.numpy())
    assert np.array_equal(np.nan_to_num(s, nan=-1), np.nan_to_num(out["stability"][:2, :3], nan=-1)), "stability differs"
    return {"tracked": len(out["tracked"]), "filtered": len(out["filtered"]), "kept_spatiotemporal": len(out["kept_st"])}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's own CPU path), bounded sample scaled to the full step
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_step(logits_cpu: torch.Tensor, prompts, n_pairs_st: int):
    """The reference's operations on a sample: per-track linear stages, greedy IoU pairs, spatio-temporal pairs.
    Returns seconds (t_linear_per_track, t_gather_pair, t_st_pair).  Works on CPU tensors (the CPU arm) and on CUDA tensors
    (the reference's own GPU path: generic ATen launches with a .item() sync per scalar)."""
    from oracle import maskpath_oracle as O
    S_, T = logits_cpu.shape[:2]
    on_gpu = logits_cpu.is_cuda
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    sync()
    t0 = time.perf_counter()
    masklets, resized = [], []
    for i in range(S_):
        frames = [O.binarize(logits_cpu[i, t][None]) for t in range(T)]                  # generate_tokens_grid.py:219 per frame
        m = torch.cat(frames, 0)                                                         # :224
        masklets.append(m)
        _ = [O.get_stability_score(logits_cpu[i, t].cpu().numpy()) for t in range(T)]    # prompt_generator.py:169 (numpy, per plane)
        resized.append(O.reshape_masklet(m))                                             # seg_utils.py:145
    sync()
    t_lin = (time.perf_counter() - t0) / S_
    t0 = time.perf_counter()
    n_g = 0
    for i in range(S_):
        for p in prompts[: 2 * S_]:
            pm = O.resize_prompt_nearest(p["segmentation"], resized[i].shape[1], resized[i].shape[2])       # :271-272
            if on_gpu:
                pm = pm.to(logits_cpu.device)              # the reference does this H2D inside the loop (:271)
            O.compute_mask_iou(resized[i][p["frame_idx"]], pm)                                                # :273
            n_g += 1
    t_g = (time.perf_counter() - t0) / max(n_g, 1)
    t0 = time.perf_counter()
    n_st = 0
    for i in range(S_):
        for j in range(i + 1, S_):
            if n_st >= n_pairs_st:
                break
            O.compute_masklet_iou(resized[i], resized[j], resized[i].device)   # seg_utils.py:110 on the resized masklets (all that exist after grid :248-250)
            n_st += 1
    t_st = (time.perf_counter() - t0) / max(n_st, 1)
    return t_lin, t_g, t_st


def cpu_arm(n_tracks, n_frames, steps, warmup, sample_tracks=4, device="cpu"):
    from sola_b200 import synth
    logits, prompts = synth.dedup_candidates(sample_tracks, n_frames, CFG["H"], CFG["W"], seed=1234 + 2, device="cpu", bin_size=CFG["bin_size"])
    if device != "cpu":
        logits = logits.to(device)
    n_pairs = sample_tracks * (sample_tracks - 1) // 2
    times = []
    for it in range(warmup + steps):
        t = cpu_reference_step(logits, prompts, n_pairs)
        if it >= warmup:
            times.append(t)
    t_lin, t_g, t_st = (float(np.mean([x[k] for x in times])) for k in range(3))
    N = n_tracks
    # full step on the CPU: N linear stages, ~N*N/2 greedy pairs (every tracked masklet vs the remaining prompts), N(N-1)/2 volume pairs
    full = N * t_lin + (N * (N - 1) // 2) * t_g + (N * (N - 1) // 2) * t_st
    value = N * n_frames / full
    sample = (f"{sample_tracks} tracks x {n_frames} frames x {CFG['H']}x{CFG['W']} through binarise+cat+stability+reshape_masklet "
              f"({t_lin * 1e3:.0f} ms/track), {2 * sample_tracks * sample_tracks} greedy mask-IoU pairs ({t_g * 1e6:.0f} us/pair), "
              f"{n_pairs} compute_masklet_iou pairs ({t_st * 1e3:.0f} ms/pair); scaled to {N} tracks: "
              f"{N}*t_track + {N * (N - 1) // 2}*(t_gather + t_masklet_pair) = {full:.1f} s/step")
    return value, full, sample


def jf_stage(device):
    """The J&F half of the metric (BASELINE configs 1 / 4), untimed with respect to the step above and informative only: the drop-in
    Evaluator.compute_J + compute_F call on config 1 (30 x 480 x 854 fp32 masklets resident on the device, one read-back per call),
    the K3 count kernel on a 320-frame 720p batch, and the boundary-F extension."""
    import sola_b200 as S
    from sola_b200 import evaluator, synth

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    p1, g1 = synth.jf_pair(30, 480, 854, seed=1235, device=device)
    p1f, g1f = p1.float(), g1.float()
    ms1 = timed(lambda: evaluator.compute_JF(p1f, g1f), 50)
    T = 320
    a = (synth.smooth_logits(T, 720, 1280, 5, device=device, cell=120) > 0).float()
    b = (synth.smooth_logits(T, 720, 1280, 6, device=device, cell=120) > 0).float()
    ms3 = timed(lambda: S.frame_counts(a, b), 10)
    pa, pb = S.pack_masks(a), S.pack_masks(b)
    msb = timed(lambda: S.boundary_counts(pa, pb), 3)
    return {"config1_compute_J_and_F_call": {"ms": ms1, "masklet_frames_per_s": 30 / ms1 * 1e3, "shape": "30 x 480 x 854 fp32, incl. read-back"},
            "K3_counts_fp32_720p": {"GBps": 2 * a.numel() * 4 / ms3 / 1e6, "masklet_frames_per_s": T / ms3 * 1e3},
            "boundary_F_720p": {"masklet_frames_per_s": T / msb * 1e3}}


# ---------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_tracks, n_frames = args.tracks, args.frames
    workload = f"config2: grid-prompt track dedup, {n_tracks} masklets x {n_frames} frames x {CFG['H']}x{CFG['W']} fp32 logits, " \
               f"stability score, miou_thresh {CFG['miou_thresh']} (one video per GPU per step)"

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        value, full, sample = cpu_arm(n_tracks, n_frames, max(1, args.steps), max(0, args.warmup))
        cores = torch.get_num_threads()
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "note": "oracle port of the reference's CPU path (the reference is Python; /root/reference does "
                           "not exist on the GPU box), torch-CPU ops with all host threads; one rank regardless of --gpus"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
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

    w = Workload(device, seed=1234 + 2 + 1000 * rank, n_tracks=n_tracks, n_frames=n_frames, n_streams=args.streams, fused=not args.no_fuse, st_native=args.st_native, tail_priority=args.tail_priority)
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
    clk.stop()
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

    # ---- e2e: logits + prompt masks start in pinned host memory; H2D in the timed region; results read back -------
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
                h = torch.empty((min(chunk, n_tracks - s), n_frames, CFG["H"], CFG["W"]), dtype=torch.float32, pin_memory=True)
                h.copy_(w.logits[s:s + chunk])
                host_chunks.append(h)
            host_prompts = torch.from_numpy(w.prompt_masks_host).pin_memory()
            dev_logits = torch.empty_like(w.logits)
            copy_stream = torch.cuda.Stream(device=device)

            def e2e_step():
                # chunked H2D on a copy stream; K1 of chunk c runs while chunk c+1 is still crossing PCIe
                evs = []
                with torch.cuda.stream(copy_stream):
                    for c, h in enumerate(host_chunks):
                        dev_logits[c * chunk: c * chunk + h.shape[0]].copy_(h, non_blocking=True)
                        e = torch.cuda.Event()
                        e.record(copy_stream)
                        evs.append(e)
                    pm = host_prompts.to(device, non_blocking=True)
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
                w.jobs[0].enqueue_after_k1(w.packed, w.counts.view(3, n_tracks, n_frames), pm)          # R1, R2, K2-gather, K2, read-backs
                return w.finish(0)

            e2e_step()
            barrier()
            n_e2e = max(1, min(args.e2e_steps, args.steps))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_e2e):
                res = e2e_step()
            b.record()
            barrier()
            te = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            h2d = w.logits.numel() * 4 + host_prompts.numel()
            d2h = int(3 * n_tracks * n_frames * 4 + n_tracks * n_tracks * 8 + 3 * n_tracks * n_tracks * 4)
            e2e = {"value": world * n_tracks * n_frames * n_e2e / (float(te.item()) * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h, "steps": n_e2e, "host_cpus_bound": len(cpus),
                   "note": "pinned host fp32 logits + uint8 prompt masks copied H2D every step (chunked, overlapped with K1); "
                           "stability counts, IoU count matrices and the N x N intersection matrix read back"}
            del host_chunks, dev_logits
        except Exception as ex:  # e.g. pinned allocation refused on a small host
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    fused_flag = w.fused
    achieved = w.k1_bytes / (k1_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload, "l2_policy": "inputs larger than L2 (18.9 GB fp32 logits per step vs 126 MB L2)",
                   "kept_sets": info, "parallelism": f"video-sharded x{world}, no data-path collective",
                   "n_x_n_planes": "native 720x1280" if args.st_native else "540x960 resized masklets (what the reference filter compares)"},
        "clocks": clk.summary(t_wall0, t_wall1),
        "stage_ms": {"dominant (K1+R1 fused)" if w.fused else "dominant (K1)": k1_ms, "K2 N x N (int pipe, carry-save, TMA-staged)": k2_ms,
                     "everything else incl. gaps": max_ms / args.steps - k1_ms - k2_ms},
        "gpu_launches": int(launches),
        "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": ("fused_pack_resize_kernel<float> (K1+R1: binarise+pack+stability+resize)" if w.fused
                                                else "pack_flat_kernel<float, THRESH3> (K1 binarise+pack+stability)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "bytes_per_launch": w.k1_bytes, "ms_per_launch": k1_ms, "share_of_step": k1_ms / (max_ms / args.steps),
                     "traffic": None},
    }
    try:
        # secondary roofline: K2 is integer-pipe work.  Denominator = alu pipe at 64 lanes/clk/SM (B300_MICROARCH.md; LOP3 measured 64-84
        # by tools/microbench_int.cu) x 148 SMs x the SM clock sampled during the timed region / 2.5 LOP3 per pair-word (4 AND + 6
        # compressor LOP3 per 4 words).  Pair-words are counted densely; the kernel skips all-zero operand quads, which is why this
        # fraction on the object-like bench masklets (~0.8) is above the dense-data one (~0.6, profiles/r1_k2_variants.jsonl).
        oh_, ow_ = (CFG["H"], CFG["W"]) if args.st_native else S.packed.default_target_shape(CFG["H"], CFG["W"])
        words_ = n_frames * oh_ * ((ow_ + 31) // 32)
        pair_words = (n_tracks * (n_tracks + 1) // 2) * words_
        sm_mhz = (line["clocks"] or {}).get("sm_mhz") or 1965.0
        if not (k2_ms > 0):
            raise ValueError("no K2 timing events")
        peak_k2 = 148 * 64 * sm_mhz * 1e6 / 2.5
        ach_k2 = pair_words / (k2_ms * 1e-3)
        line["roofline_k2"] = {"bound": "int-alu", "kernel": "pair_iou_st_ring_kernel (K2 N x N AND-popcount, carry-save)",
                               "achieved": ach_k2 / 1e12, "peak": peak_k2 / 1e12, "unit": "T pair-words/s", "frac": ach_k2 / peak_k2,
                               "peak_source": "148 SMs x 64 alu lanes/clk x sampled SM clock / 2.5 LOP3 per pair-word",
                               "pair_words_per_launch": int(pair_words), "ms_per_launch": k2_ms}
    except Exception as ex:
        line["roofline_k2"] = {"error": repr(ex)[:200]}
    if world == 1:
        try:
            line["jf_stage"] = jf_stage(device)
        except Exception as ex:
            line["jf_stage"] = {"error": repr(ex)[:200]}
    if not args.no_cpu_baseline and world == 1:
        v, full, sample = cpu_arm(n_tracks, n_frames, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample}
        # second, informative baseline: the same reference operations on CUDA tensors — generic ATen kernels plus
        # one .item() sync per scalar, which is how the reference actually runs this path on a GPU
        try:
            del w                                   # release the bench buffers so ATen's temporaries do not fight the allocator
            torch.cuda.empty_cache()
            v2, full2, sample2 = cpu_arm(n_tracks, n_frames, 2, 2, device=device)
            line["gpu_aten_baseline"] = {"value": v2, "unit": UNIT, "kind": "port on CUDA tensors (ATen)", "sample": sample2}
        except Exception as ex:
            line["gpu_aten_baseline"] = {"value": None, "error": repr(ex)[:200]}
    traffic_path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")       # from the committed `ncu --set full` capture
    if os.path.isfile(traffic_path) and (n_tracks, n_frames) == (CFG["n_tracks"], CFG["n_frames"]):
        with open(traffic_path) as f:
            key = "fused_pack_resize_kernel<float>" if fused_flag else "pack_flat_kernel<float, THRESH3>"
            line["roofline"]["traffic"] = json.load(f).get(key, {}).get("dram_bytes_per_launch")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
