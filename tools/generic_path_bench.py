"""Throughput of the generic (W % 32 != 0) paths at 480x854 — diagnostic."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import synth

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
n, H, W = 2560, 480, 854
x = synth.smooth_logits(n, H, W, 1, device="cuda", cell=80)
ms = timed(lambda: S.binarize_pack_stability(x))
out["K1_rows_480x854"] = {"ms": ms, "GBps": x.numel() * 4 / ms / 1e6}
p, _ = S.binarize_pack_stability(x)
ms = timed(lambda: S.resize_bilinear_bin(p))
out["R1_480x854_to_540x960"] = {"ms": ms, "frames_per_s": n / ms * 1e3}
ms = timed(lambda: S.binarize_pack_resize(x))
out["K1R1_unfused_fallback_480x854"] = {"ms": ms, "GBps": x.numel() * 4 / ms / 1e6}
xb = x.to(torch.bfloat16)
ms = timed(lambda: S.binarize_pack_stability(xb))
out["K1_rows_bf16_480x854"] = {"ms": ms, "GBps": xb.numel() * 2 / ms / 1e6}
y = synth.smooth_logits(n, 480, 864, 1, device="cuda", cell=80)
ms = timed(lambda: S.binarize_pack_resize(y))
out["K1R1_fused_480x864"] = {"ms": ms, "GBps": y.numel() * 4 / ms / 1e6}
z = synth.smooth_logits(1280, 720, 1280, 1, device="cuda", cell=80).to(torch.bfloat16)
ms = timed(lambda: S.binarize_pack_resize(z))
out["K1R1_fused_bf16_720p"] = {"ms": ms, "GBps": z.numel() * 2 / ms / 1e6}
print(json.dumps(out, indent=1))
