"""K3 / J&F throughput (diagnostic; bench.py's headline is config 2): per-frame counts from fp32, uint8 and packed masklets,
boundary-F counts, and the config-1 J&F call, with inputs resident in HBM."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import evaluator, synth

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
T, H, W = 2560, 720, 1280                       # 32 masklets x 80 frames, 9.4 GB per fp32 operand
pred8 = (synth.smooth_logits(T, H, W, 1, device="cuda", cell=120) > 0).to(torch.uint8)
gt8 = (synth.smooth_logits(T, H, W, 2, device="cuda", cell=120) > 0).to(torch.uint8)
predf, gtf = pred8.float(), gt8.float()
ms = timed(lambda: S.frame_counts(predf, gtf))
out["K3_f32_counts"] = {"ms": ms, "GBps": 2 * predf.numel() * 4 / ms / 1e6, "masklet_frames_per_s": T / ms * 1e3}
ms = timed(lambda: S.frame_counts(pred8, gt8))
out["K3_u8_counts"] = {"ms": ms, "GBps": 2 * pred8.numel() / ms / 1e6, "masklet_frames_per_s": T / ms * 1e3}
pp, gp = S.pack_masks(pred8).reshape_lead(1, T), S.pack_masks(gt8).reshape_lead(1, T)
ms = timed(lambda: S.frame_counts_packed(pp, gp))
out["K3_packed_counts"] = {"ms": ms, "GBps": 2 * pp.words.numel() * 4 / ms / 1e6, "masklet_frames_per_s": T / ms * 1e3}
ms = timed(lambda: S.boundary_counts(pp.reshape_lead(T), gp.reshape_lead(T)), reps=3)
out["boundary_counts_720p_r12"] = {"ms": ms, "masklet_frames_per_s": T / ms * 1e3}
del predf, gtf
# config 1: one (video, expression) 30 x 480 x 854, drop-in call incl. the single read-back
p1, g1 = synth.jf_pair(30, 480, 854, 3, device="cuda")
p1f, g1f = p1.float(), g1.float()
t0 = time.perf_counter()
for _ in range(200): evaluator.compute_JF(p1f, g1f)
dt = (time.perf_counter() - t0) / 200
out["config1_compute_JF_call_fp32"] = {"ms_per_call": dt * 1e3, "masklet_frames_per_s": 30 / dt}
t0 = time.perf_counter()
for _ in range(200): evaluator.compute_JF(p1, g1)
dt = (time.perf_counter() - t0) / 200
out["config1_compute_JF_call_u8"] = {"ms_per_call": dt * 1e3, "masklet_frames_per_s": 30 / dt}
print(json.dumps(out, indent=1))
