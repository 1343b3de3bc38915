"""Batched label counts (sola_frame_counts_packed: 64 tracks x 3 GT objects x 80 frames of 540x960 planes) alone: ms and GB/s
(algorithmic bytes = every track plane once + every GT plane once)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import synth

N, G, T, H, W = 64, 3, 80, 540, 960
logits, _ = synth.dedup_candidates(N, T, 720, 1280, seed=1236, device="cuda")
_, _, tracks = S.binarize_pack_resize(logits)
del logits
gt = S.pack_masks(torch.stack([synth.blob_masklet(T, H, W, 40 + g, device="cuda") for g in range(G)]))
for _ in range(3):
    out = S.frame_counts_packed(tracks, gt)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    S.frame_counts_packed(tracks, gt)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
nbytes = (tracks.words.numel() + gt.words.numel()) * 4
print(json.dumps({"build": os.environ.get("SOLA_EXTRA_NVCC_FLAGS", "default"), "ms": ms, "GBps": nbytes / ms / 1e6,
                  "checksum": [int(x.sum().item()) for x in out]}))
