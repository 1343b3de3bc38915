"""Per-stage device time of one bench step (synchronised between stages; diagnostic, not a bench value)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import dedup
import bench

w = bench.Workload(torch.device("cuda", 0), seed=1236, n_tracks=64, n_frames=80)
w.make_jobs()
for _ in range(3):
    w.step()
torch.cuda.synchronize()

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return r, float(np.median(ts))

out = {}
(packed, counts), out["K1_binarize_pack_stability"] = timed(lambda: S.binarize_pack_stability(w.logits, 0.0, 1.0, out=w.packed, counts_out=w.counts))
_, out["K1R1_fused"] = timed(lambda: S.binarize_pack_resize(w.logits, 0.0, 1.0, out=w.packed, counts_out=w.counts))
resized, out["R1_resize_bilinear_packed"] = timed(lambda: S.resize_bilinear_bin(packed))
mk = lambda: dedup.TrackDedup(w.prompt_meta, w.T, mode="grid", prompt_masks=w.prompt_masks_dev, bin_size=4, n_max_tracks=64, batch_size=4, miou_thresh=0.7)
dd, out["R2_trackdedup_init_nearest"] = timed(mk)
_, out["G1_gather_one_launch_plus_host"] = timed(lambda: mk().run_offline(resized))
def per_batch():
    d = mk()
    while (b := d.next_batch()) is not None:
        d.submit_resized(b, resized[b])
    return d.result()
_, out["G1_per_batch_sessions"] = timed(per_batch)
_, out["K2_pair_iou_st_kernel_only"] = timed(lambda: S.pairwise_inter_matrix(packed))
_, out["K2_dedup_matrix_total"] = timed(lambda: dedup.dedup_matrix(packed, 0.7))
_, out["stability_d2h"] = timed(lambda: S.packed.stability_from_counts(counts))
_, out["full_step_sync"] = timed(lambda: w.step())
_, t10 = timed(lambda: w.run_steps(10), reps=3)
out["pipelined_step"] = t10 / 10
pred = S.unpack_masks(packed[0], torch.float32); gt = S.unpack_masks(packed[1], torch.float32)
_, out["K3_frame_counts_f32_80x720p"] = timed(lambda: S.frame_counts(pred, gt))
out["K3_GBps"] = 2 * pred.numel() * 4 / (out["K3_frame_counts_f32_80x720p"] * 1e-3) / 1e9
print(json.dumps(out, indent=1))
