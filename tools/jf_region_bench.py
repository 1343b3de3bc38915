"""Region mode (J and F as evaluator.py:227-247 defines them) of the fused J&F kernel: GB/s per shape and on the bench's mixed
MeViS-like sweep.  One JSON line; algorithmic bytes = both packed planes read once."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import packed as P, synth


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"build": os.environ.get("SOLA_EXTRA_NVCC_FLAGS", "default")}
for name, (T, H, W) in {"360x640": (4096, 360, 640), "480x854": (4096, 480, 854), "720x1280": (2560, 720, 1280), "1080x1920": (1024, 1080, 1920)}.items():
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    wp = (W + 31) // 32
    pw = torch.randint(-2**31, 2**31 - 1, (T, H, wp), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    gw = torch.randint(-2**31, 2**31 - 1, (T, H, wp), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    if W % 32:
        pw[..., -1] &= (1 << (W % 32)) - 1; gw[..., -1] &= (1 << (W % 32)) - 1
    plan = P.JFSweepPlan([(P.PackedMasks(pw, H, W), P.PackedMasks(gw, H, W))], with_boundary=False)
    buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device="cuda")
    ms = timed(lambda: plan.run(buf))
    out[name] = {"ms": round(ms, 4), "GBps": round(plan.algorithmic_bytes / ms / 1e6, 1), "bands": plan.bands[0], "smem": plan.smem_bytes}
    del pw, gw
units = synth.mevis_like_sweep(12, 4, 1238, "cuda", t_range=(30, 120), pack=S.pack_masks)
plan = P.JFSweepPlan([(p, g) for _, _, p, g in units], with_boundary=False)
buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device="cuda")
ms = timed(lambda: plan.run(buf))
out["bench_sweep"] = {"ms": round(ms, 4), "GBps": round(plan.algorithmic_bytes / ms / 1e6, 1), "items": plan.n_items, "frames": plan.total_frames}
print(json.dumps(out))
