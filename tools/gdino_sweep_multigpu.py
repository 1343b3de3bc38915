"""BASELINE config 3 shape: the GroundingDINO-prompt path (generate_tokens_gdino.py:155-206,288-300) — per (video, expression) unit
16 candidate masklets x 36 frames x 720x1280, stability_score_thresh 0.85, n_max_tracks 16, batch_size 4 — with the units dealt to
the ranks of one box by the reference's modulo rule and ONE all-reduce of integer audit totals at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gdino_sweep_multigpu.py

Per unit: K1 on the 16 prompt frames gives the stability scores the filter at :162 reads (in the reference they were stored at prompt
generation, generate_prompts_gdino.py:179), then one VideoDedupJob (fused K1+R1, R2, K2 gather, read-back) and the gdino greedy
replay on the host.  Rank 0 repeats the whole sweep alone (untimed) to show the sharded audit is identical."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S  # noqa: E402
from sola_b200 import dedup, sharding, synth  # noqa: E402

RULES = dict(bin_size=4, n_max_tracks=16, batch_size=4, miou_thresh=0.7, stability_score_thresh=0.85)


def run_unit(seed, n, T, H, W, device):
    logits, prompts = synth.dedup_candidates(n, T, H, W, seed=seed, device=device, bin_size=RULES["bin_size"])
    fidx = torch.as_tensor([p["frame_idx"] for p in prompts], device=device)
    _, c = S.binarize_pack_stability(logits[torch.arange(n, device=device), fidx], want_packed=False)     # (3, n): the prompt frames only
    stab = S.packed.stability_from_counts(c.cpu().numpy())
    meta = [{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"], "expression_id": "0", "stability_score": float(stab[k])}
            for k, p in enumerate(prompts)]
    masks = torch.from_numpy(np.stack([p["segmentation"] for p in prompts])).to(device)
    job = dedup.VideoDedupJob(meta, T, device=device, mode="gdino", expression_id="0", **RULES)
    job.enqueue(logits, masks)
    r = job.finish()
    return np.array([1, len(r["tracked"]), len(r["filtered"]), r["n_not_used"], sum(r["tracked"]), sum(r["filtered"]),
                     int(np.sum(np.asarray(r["inter"]).diagonal() % 1000003))], dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=48, help="(video, expression) units in the sweep (Ref-YouTube-VOS valid has ~800)")
    ap.add_argument("--tracks", type=int, default=16)
    ap.add_argument("--frames", type=int, default=36)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = sharding.init_process_group_from_env(device)
    mine = sharding.shard_indices(args.units, rank, world)
    run_unit(10_000, args.tracks, args.frames, 720, 1280, device)                    # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tot = np.zeros(7, dtype=np.int64)
    for u in mine:
        tot += run_unit(20_000 + u, args.tracks, args.frames, 720, 1280, device)
    t = torch.from_numpy(tot).to(device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                                     # the path's only collective
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = np.zeros(7, dtype=np.int64)
        for u in range(args.units):
            ref += run_unit(20_000 + u, args.tracks, args.frames, 720, 1280, device)
        got = t.cpu().numpy()
        frames = args.units * args.tracks * args.frames
        print(json.dumps({"workload": f"config3-shaped gdino sweep: {args.units} units x {args.tracks} masklets x {args.frames} frames x 720x1280",
                          "n_gpus": world, "units": int(got[0]), "tracked": int(got[1]), "filtered": int(got[2]), "not_used": int(got[3]),
                          "seconds_max_over_ranks": float(dt.item()), "masklet_frames_per_s": frames / float(dt.item()),
                          "audit_identical_to_1_rank": bool(np.array_equal(got, ref)),
                          "includes": "synthetic logit generation on the device and the per-unit host greedy (not a bench value)"}))
        assert np.array_equal(got, ref)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
