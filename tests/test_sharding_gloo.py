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


# ---- BASELINE config 5: one video's N x N matrix with the WORD axis partitioned over the ranks -----------------------------------
def _np_inter_words(w: torch.Tensor) -> torch.Tensor:
    """numpy stand-in for the K2 kernel in this CPU test: (N, words) int32 rows -> int64 (N, N) AND-popcounts."""
    bits = np.unpackbits(w.numpy().view(np.uint8), axis=1).astype(np.int64)
    return torch.from_numpy(bits @ bits.T)


def _cfg5_planes(rank, n_local=3, T=3, H=9, W=70):
    rng = np.random.default_rng(100 + rank)
    Wp = (W + 31) // 32
    words = rng.integers(0, 2 ** 32, size=(n_local, T, H, Wp), dtype=np.uint64).astype(np.uint32)
    words[..., -1] &= np.uint32((1 << (W % 32)) - 1)                       # pad bits stay zero
    return words.view(np.int32), H, W


def _worker_cfg5(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sharding.init_process_group_from_env(device=None)
    from sola_b200 import packed as P
    P.pairwise_inter_matrix_words = _np_inter_words                        # the exchange / partition / reduction logic is what is under test
    words, H, W = _cfg5_planes(rank)
    local = P.PackedMasks(torch.from_numpy(words.copy()), H, W)
    inter = sharding.pairwise_inter_matrix_sharded(local, split="words")
    q.put((rank, inter.numpy().tolist()))
    dist.destroy_process_group()


def test_two_rank_word_partitioned_matrix_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_cfg5, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allw = np.concatenate([_cfg5_planes(r)[0] for r in range(2)])          # rank-major track order, as the sharded path defines it
    exp = _np_inter_words(torch.from_numpy(allw.reshape(allw.shape[0], -1).copy())).numpy()
    assert np.array_equal(np.asarray(got[0]), exp) and got[0] == got[1]


def test_word_slices_cover_the_axis_once():
    for words, world in [(1, 2), (31, 2), (32, 2), (33, 8), (3456000, 8), (100, 3)]:
        sl = sharding.word_slices(words, world)
        assert sl[0][0] == 0 and sl[-1][1] == words and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert all(lo % 32 == 0 for lo, _ in sl)
