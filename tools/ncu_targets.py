"""ncu targets: `python tools/ncu_targets.py <target>` launches the named kernel a few times on a realistic input, nothing else of ours.
Targets: fused_bf16, fused_480p, fused_720p, k3_f32, labels, gather, k2_dense, k2_object, jf_region, jf_boundary, rle_decode"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import packed as P, synth, rle

t = sys.argv[1]
dev = "cuda"
if t in ("fused_bf16", "fused_720p", "fused_480p", "fused_480x864"):
    H, W = {"fused_480p": (480, 854), "fused_480x864": (480, 864)}.get(t, (720, 1280))
    logits, _ = synth.dedup_candidates(16, 80, H, W, seed=1236, device=dev)
    if t == "fused_bf16":
        logits = logits.to(torch.bfloat16)
    for _ in range(4):
        S.binarize_pack_resize(logits)
elif t == "k3_f32":
    a = (synth.smooth_logits(640, 720, 1280, 5, device=dev, cell=120) > 0).float()
    b = (synth.smooth_logits(640, 720, 1280, 6, device=dev, cell=120) > 0).float()
    for _ in range(4):
        S.frame_counts(a, b)
elif t in ("labels", "gather", "k2_object", "k2_dense"):
    N, T, H, W = 64, 80, 540, 960
    if t == "k2_dense":
        g = torch.Generator(device=dev); g.manual_seed(1)
        words = torch.randint(-2**31, 2**31 - 1, (N, T, H, 30), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
        tracks = P.PackedMasks(words, H, W)
    else:
        logits, prompts = synth.dedup_candidates(N, T, 720, 1280, seed=1236, device=dev)
        _, _, tracks = S.binarize_pack_resize(logits)
        del logits
    if t == "labels":
        gt = S.pack_masks(torch.stack([synth.blob_masklet(T, H, W, 40 + g, device=dev) for g in range(3)]))
        for _ in range(4):
            S.frame_counts_packed(tracks, gt)
    elif t == "gather":
        pm = S.resize_nearest(torch.from_numpy(np.stack([p["segmentation"] for p in prompts])).to(dev), H, W)
        fi = [p["frame_idx"] for p in prompts]
        for _ in range(4):
            S.gathered_inter(tracks, pm, fi)
    else:
        for _ in range(4):
            S.pairwise_inter_matrix(tracks)
elif t in ("jf_region", "jf_boundary"):
    units = synth.mevis_like_sweep(12, 4, 1238, dev, t_range=(30, 120), pack=S.pack_masks)
    plan = P.JFSweepPlan([(p, g) for _, _, p, g in units], with_boundary=(t == "jf_boundary"))
    for _ in range(4):
        plan.run()
elif t == "rle_decode":
    from oracle import rle_oracle as RO           # input fabrication only
    m = synth.blob_masklet(64, 480, 854, 3).numpy()
    rl = RO.encode_masklet(m)
    for _ in range(4):
        rle.decode_rle_masklets_merged([rl, rl])
else:
    raise SystemExit(f"unknown target {t}")
torch.cuda.synchronize()
print("ok", t)
