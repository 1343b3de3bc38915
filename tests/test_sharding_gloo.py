"""CPU, world_size 2 over gloo: the N>1 path (round-robin units, local accumulators, ONE all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import maskpath_oracle as O
from sola_b200 import evaluator, sharding


def _units(n=7):
    rng = np.random.default_rng(42)
    out = []
    for u in range(n):
        T, H, W = int(rng.integers(2, 6)), int(rng.integers(8, 20)), int(rng.integers(8, 40))
        gt = (rng.random((T, H, W)) > 0.5).astype(np.uint8)
        pred = gt ^ (rng.random((T, H, W)) > 0.9).astype(np.uint8)
        out.append((pred, gt))
    return out


def _local_sums(units, idx):
    sJ = sF = sJF = 0.0
    tot = np.zeros(3, dtype=np.int64)
    for i in idx:
        c = O.jf_counts_exact(*units[i])                       # stands in for the device counts in this CPU test
        J, F = float(evaluator.J_from_counts(*c)), float(evaluator.F_from_counts(*c))
        sJ, sF, sJF = sJ + J, sF + F, sJF + (J + F) / 2
        tot += np.array([x.sum() for x in c], dtype=np.int64)
    return sJ, sF, sJF, tot


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = sharding.init_process_group_from_env(device=None)
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    units = _units()
    idx = sharding.shard_indices(len(units), rank, world)
    sJ, sF, sJF, tot = _local_sums(units, idx)
    res = sharding.allreduce_jf(sJ, sF, sJF, len(idx), tot)
    q.put((rank, res["mean_J"], res["mean_F"], res["mean_JF"], res["n_units"], res["int_totals"].tolist()))
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = _units()
    sJ, sF, sJF, tot = _local_sums(units, range(len(units)))
    n = len(units)
    for rank, mJ, mF, mJF, n_units, totals in got:
        assert n_units == n and totals == tot.tolist()                      # integer audit: bit-identical
        assert abs(mJ - sJ / n) < 1e-12 and abs(mF - sF / n) < 1e-12 and abs(mJF - sJF / n) < 1e-12
    assert got[0][1:] == got[1][1:]                                         # both ranks hold the same result


def test_allreduce_is_identity_without_process_group():
    res = sharding.allreduce_jf(1.5, 2.5, 2.0, 2, np.array([10, 20, 30]))
    assert res["mean_J"] == 0.75 and res["n_units"] == 2 and res["int_totals"].tolist() == [10, 20, 30]
