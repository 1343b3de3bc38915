"""Fused K1+R1 (and K1 alone) away from the bench's headline shape: bf16 logits and <= 480p inputs, on two kinds of synthetic logits —
  object : the bench's generator (synth.dedup_candidates: smooth object fields + 0.5 logit noise at 15-80 logit units per field unit —
           clean mask edges, what upsampled SAM2 logits look like);
  speckle: synth.smooth_logits(gain 6, noise 0.5): the noise is comparable to the edge slope, so every boundary is a band of random
           bits several pixels wide — an adversarial case for the resize phase (one evaluation per output word that straddles an edge).
GB/s counts the algorithmic bytes (logits in + native planes + resized planes out)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import synth


def timed(fn, reps=8):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def logits_of(kind, n_tracks, T, H, W, dtype):
    if kind == "object":
        x, _ = synth.dedup_candidates(n_tracks, T, H, W, seed=1236, device="cuda")
    else:
        x = synth.smooth_logits(n_tracks * T, H, W, 1, device="cuda", cell=80).view(n_tracks, T, H, W)
    return x.to(dtype)


out = {}
for name, (H, W, dtype, n_tracks) in {"bf16_720x1280": (720, 1280, torch.bfloat16, 32), "f32_720x1280": (720, 1280, torch.float32, 32),
                                      "f32_480x854": (480, 854, torch.float32, 48), "f32_480x864": (480, 864, torch.float32, 48),
                                      "bf16_480x854": (480, 854, torch.bfloat16, 48), "f32_360x640": (360, 640, torch.float32, 64),
                                      "f32_1080x1920": (1080, 1920, torch.float32, 12)}.items():
    for kind in ("object", "speckle"):
        x = logits_of(kind, n_tracks, 80, H, W, dtype)
        n = n_tracks * 80
        oh, ow = S.packed.default_target_shape(H, W)
        k1_bytes = x.numel() * x.element_size() + n * H * ((W + 31) // 32) * 4
        fused_bytes = k1_bytes + n * oh * ((ow + 31) // 32) * 4
        ms_k1 = timed(lambda: S.binarize_pack_stability(x))
        ms_f = timed(lambda: S.binarize_pack_resize(x))
        out[f"{name}/{kind}"] = {"K1_ms": ms_k1, "K1_GBps": k1_bytes / ms_k1 / 1e6, "fused_ms": ms_f, "fused_GBps": fused_bytes / ms_f / 1e6,
                                 "fused_frac_of_6555.8": fused_bytes / ms_f / 1e6 / 6555.8, "frames": n}
        del x
print(json.dumps(out, indent=1))
