"""BASELINE config 4 shape: a MeViS-valid_u-like J&F sweep (many (video, expression) units of mixed length / resolution) sharded
over the ranks of one box, ONE NCCL all-reduce of the accumulators at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/jf_sweep_multigpu.py

Every rank also recomputes the full sweep's integer audit from the unit list's seeds on rank 0 (untimed) to show that the sharded
result is bit-identical for the integer accumulators and within 1e-12 for the float64 means."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sola_b200 import evaluator, sharding, synth  # noqa: E402

SHAPES = [(360, 640), (480, 854), (720, 1280), (1080, 1920)]


def unit_list(n_videos=24, seed=4):
    rng = np.random.default_rng(seed)
    units = []
    for v in range(n_videos):
        H, W = SHAPES[int(rng.integers(0, len(SHAPES)))]
        T = int(rng.integers(30, 121))
        for e in range(int(rng.integers(2, 7))):
            units.append({"video": f"v{v:03d}", "exp": str(e), "T": T, "H": H, "W": W, "seed": 1000 * v + e,
                          "none": bool(rng.random() < 0.03)})
    return units


def run(units, idx, device):
    sweep = evaluator.JFSweep(device)
    frames = 0
    for i in idx:
        u = units[i]
        pred, gt = synth.jf_pair(u["T"], u["H"], u["W"], u["seed"], device=device)        # stands in for decoded RLE masklets (uint8)
        sweep.add((u["video"], u["exp"]), None if u["none"] else pred, gt)
        frames += u["T"]
    results, totals = sweep.finish()
    sJ = sum(r["J"] for _, r in results)
    sF = sum(r["F"] for _, r in results)
    sJF = sum(r["JF"] for _, r in results)
    return sJ, sF, sJF, len(results), totals, frames


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = sharding.init_process_group_from_env(device)
    units = unit_list()
    costs = [u["T"] * u["H"] * u["W"] for u in units]
    idx = sharding.shard_balanced(costs, rank, world)
    run(units, idx[:2], device)                                    # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sJ, sF, sJF, n, totals, frames = run(units, idx, device)
    res = sharding.allreduce_jf(sJ, sF, sJF, n, totals, device)    # the path's only collective
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    fr = torch.tensor([frames], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    if rank == 0:
        ref = run(units, range(len(units)), device)                # single-rank reference, untimed
        ok_int = ref[4].tolist() == res["int_totals"].tolist() and ref[3] == res["n_units"]
        ok_f = abs(ref[0] / ref[3] - res["mean_J"]) < 1e-12 and abs(ref[1] / ref[3] - res["mean_F"]) < 1e-12
        print(json.dumps({"workload": "config4-shaped J&F sweep", "n_gpus": world, "units": res["n_units"], "masklet_frames": int(fr.item()),
                          "seconds_max_over_ranks": float(dt.item()), "masklet_frames_per_s": float(fr.item()) / float(dt.item()),
                          "mean_J": res["mean_J"], "mean_F": res["mean_F"], "mean_JF": res["mean_JF"],
                          "int_totals": res["int_totals"].tolist(), "integer_audit_identical_to_1_rank": ok_int,
                          "float_means_within_1e-12": ok_f, "includes": "synthetic mask generation on the device (not a bench value)"}))
        assert ok_int and ok_f
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
