"""Fused J&F kernel (csrc/jf_fused.cu) throughput on packed planes resident in HBM: uniform shapes of BASELINE configs 1 / 4 / 5 and a
mixed-shape MeViS-like sweep, on object-like prediction errors and on speckled ones (boundary pixels everywhere: the dense worst case).
Prints one JSON object; GB/s counts the algorithmic bytes (both packed planes read once)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import packed as P, synth


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def unit_batch(T, H, W, seed, speckle, chunk=64):
    ps, gs = [], []
    for s in range(0, T, chunk):
        p, g = synth.object_pair(min(chunk, T - s), H, W, seed + s, "cuda", speckle=speckle)
        ps.append(S.pack_masks(p).words); gs.append(S.pack_masks(g).words)
    return P.PackedMasks(torch.cat(ps), H, W), P.PackedMasks(torch.cat(gs), H, W)


out = {}
quick = "--quick" in sys.argv
only_auto = "--auto-only" in sys.argv
for name, (T, H, W) in {"480x854": (4096, 480, 854), "720x1280": (2560, 720, 1280), "1080x1920": (1024, 1080, 1920), "360x640": (4096, 360, 640)}.items():
    if quick: T //= 8
    for kind, speckle in (("object", 0.0), ("speckle2pct", 0.02)):
        pp, gp = unit_batch(T, H, W, 11, speckle)
        nbytes = 2 * pp.words.numel() * 4
        for wb in (True, False):
            plan = P.JFSweepPlan([(pp, gp)], with_boundary=wb)
            buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device="cuda")
            ms = timed(lambda: plan.run(buf), reps=5 if speckle else 20)
            out[f"{name}/{kind}/{'J+F+boundary' if wb else 'J+F'}"] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frames_per_s": T / ms * 1e3,
                                                                      "bands": plan.bands[0], "smem": plan.smem_bytes}
        for occ in (() if only_auto else (3, 2, 1)):                # forced tile class (CTAs per SM) in boundary mode
            try:
                plan = P.JFSweepPlan([(pp, gp)], with_boundary=True, ctas_per_sm=occ)
                buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device="cuda")
                ms = timed(lambda: plan.run(buf), reps=5 if speckle else 20)
                out[f"{name}/{kind}/J+F+boundary/ctas{occ}"] = {"ms": ms, "frames_per_s": T / ms * 1e3, "bands": plan.bands[0], "smem": plan.smem_bytes}
            except Exception as ex:
                out[f"{name}/{kind}/J+F+boundary/ctas{occ}"] = {"error": repr(ex)[:120]}
        if kind == "object":
            ms = timed(lambda: S.frame_counts_packed(pp.reshape_lead(1, T), gp.reshape_lead(1, T)), reps=20)
            out[f"{name}/object/old_K3_packed_counts"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
        del pp, gp
units = synth.mevis_like_sweep(6 if quick else 24, 4, 1238, "cuda", t_range=(30, 120), pack=S.pack_masks)
plan = P.JFSweepPlan([(p, g) for _, _, p, g in units], with_boundary=True)
buf = torch.empty((7, plan.total_frames), dtype=torch.int32, device="cuda")
ms = timed(lambda: plan.run(buf), reps=10)
for occ in (() if only_auto else (3, 2)):
    pl = P.JFSweepPlan([(p, g) for _, _, p, g in units], with_boundary=True, ctas_per_sm=occ)
    b2 = torch.empty((7, pl.total_frames), dtype=torch.int32, device="cuda")
    m2 = timed(lambda: pl.run(b2), reps=10)
    out[f"mevis_like_sweep/J+F+boundary/ctas{occ}"] = {"ms": m2, "frames_per_s": pl.total_frames / m2 * 1e3, "launches": len(pl.launches)}
out["mevis_like_sweep/J+F+boundary"] = {"ms": ms, "GBps": plan.algorithmic_bytes / ms / 1e6, "frames_per_s": plan.total_frames / ms * 1e3,
                                        "units": plan.n_units, "frames": plan.total_frames, "items": plan.n_items}
print(json.dumps(out, indent=1))
