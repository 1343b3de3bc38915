"""ncu target: a few launches of the fused J&F kernel on one 720p (default) batch of object-like masklets."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import packed as P, synth
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
speckle = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
ps, gs = [], []
for s in range(0, T, 64):
    p, g = synth.object_pair(min(64, T - s), H, W, 11 + s, "cuda", speckle=speckle)
    ps.append(S.pack_masks(p).words); gs.append(S.pack_masks(g).words)
pp, gp = P.PackedMasks(torch.cat(ps), H, W), P.PackedMasks(torch.cat(gs), H, W)
plan = P.JFSweepPlan([(pp, gp)], with_boundary=True)
for _ in range(4):
    plan.run()
torch.cuda.synchronize()
print("ok", plan.n_items, plan.bands)
