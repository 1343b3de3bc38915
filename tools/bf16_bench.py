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
n = 2560
z = synth.smooth_logits(n, 720, 1280, 1, device="cuda", cell=120, bias=0.9, gain=30.0).to(torch.bfloat16)
ms = timed(lambda: S.binarize_pack_stability(z)); out["K1_bf16_720p"] = {"ms": ms, "GBps": z.numel() * 2 / ms / 1e6}
p, _ = S.binarize_pack_stability(z)
ms = timed(lambda: S.resize_bilinear_bin(p)); out["R1_720p"] = {"ms": ms}
ms = timed(lambda: S.binarize_pack_resize(z)); out["fused_bf16_720p"] = {"ms": ms, "GBps": z.numel() * 2 / ms / 1e6}
zf = z.float()
ms = timed(lambda: S.binarize_pack_resize(zf)); out["fused_f32_720p"] = {"ms": ms, "GBps": zf.numel() * 4 / ms / 1e6}
ms = timed(lambda: S.binarize_pack_stability(zf)); out["K1_f32_720p"] = {"ms": ms, "GBps": zf.numel() * 4 / ms / 1e6}
del zf
y = synth.smooth_logits(1024, 1080, 1920, 1, device="cuda", cell=160, bias=0.9, gain=30.0)
ms = timed(lambda: S.binarize_pack_resize(y)); out["fused_f32_1080p"] = {"ms": ms, "GBps": y.numel() * 4 / ms / 1e6}
print(json.dumps(out, indent=1))
