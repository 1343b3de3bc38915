"""Worker of tests/test_gpu_multirank.py (one process per GPU under torchrun): BASELINE config 5's exchange paths on a small video.
Every variant must give the int64 N x N matrix the single-rank kernel gives on the gathered planes."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S  # noqa: E402
from sola_b200 import sharding  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank, world = sharding.init_process_group_from_env(device)
    n_local, T, H, W = 16, 6, 54, 96                     # 16 tracks per rank; 6 * 54 * 3 = 972 words per track (multiple of 4)
    rng = np.random.default_rng(1000 + rank)
    masks = rng.random((n_local, T, H, W)) > rng.random((n_local, 1, 1, 1))
    masks[0] = masks[1]                                   # a duplicate inside the rank
    results = {}
    peers = sharding.PeerPlanes(n_local, T, H, W, device, n_chunks=3)
    peers.local.words.copy_(S.pack_masks(masks).words)
    torch.cuda.synchronize()
    dist.barrier()
    gathered = torch.empty((n_local * world, T, H, peers.local.Wp), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, peers.local.words.contiguous())
    alone = S.pairwise_inter_matrix(S.PackedMasks(gathered, H, W))
    variants = {
        "nccl_words": lambda: sharding.pairwise_inter_matrix_sharded(peers.local, split="words"),
        "nccl_tiles": lambda: sharding.pairwise_inter_matrix_sharded(peers.local, split="tiles"),
        "peer_pull": lambda: peers.pairwise_inter_matrix("pull"),
        "peer_direct": lambda: peers.pairwise_inter_matrix("direct"),
    }
    variants["peer_tma"] = lambda: peers.pairwise_inter_matrix("tma")
    variants["peer_auto"] = lambda: peers.pairwise_inter_matrix()
    for name, fn in variants.items():
        try:
            m = fn()
            torch.cuda.synchronize()
            results[name] = bool(torch.equal(m, alone))
        except Exception as ex:                           # report, let the launcher decide
            results[name] = f"error: {ex!r}"[:300]
    # the J&F half: units dealt by the modulo rule, one all-reduce of the accumulators; integer audit identical on every rank
    from sola_b200 import evaluator, synth
    units = [synth.object_pair(4, 48, 85, 50 + u, device="cpu") for u in range(5)]
    mine = sharding.shard_indices(len(units), rank, world)
    sweep = evaluator.JFSweep(device, with_boundary=True)
    for u in mine:
        sweep.add(u, *units[u])
    res, tot = sweep.finish()
    red = sharding.allreduce_jf(sum(r["J"] for _, r in res), sum(r["F"] for _, r in res), sum(r["JF"] for _, r in res), len(res), tot, device=device)
    full_tot = np.zeros(3, np.int64)
    for p, g in units:
        p, g = p.numpy().astype(bool), g.numpy().astype(bool)
        full_tot += np.array([(p & g).sum(), p.sum(), g.sum()], dtype=np.int64)
    results["jf_allreduce_int_totals"] = bool(red["n_units"] == len(units) and np.array_equal(red["int_totals"], full_tot))
    if rank == 0:
        print("MULTIRANK_RESULT " + json.dumps({"world": world, **results}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
