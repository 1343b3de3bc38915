"""CPU: host-side logic of the product (greedy state machine, count->metric formulas, sharding) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import greedy_oracle as GO
from oracle import maskpath_oracle as O
from sola_b200 import dedup, evaluator, utils, metric, sharding
from conftest import GREEDY_CASES, greedy_prompts, greedy_table


def _oracle_iou_rows(masklets_f32, batch, prompts):
    """What the device computes per batch, restated with the oracle: IoU(resized track[frame_k], nearest prompt_k)."""
    rows = np.zeros((len(batch), len(prompts)))
    for b, k in enumerate(batch):
        rt = O.reshape_masklet(masklets_f32[k])
        for j, p in enumerate(prompts):
            pm = O.resize_prompt_nearest(p["segmentation"], rt.shape[1], rt.shape[2])
            rows[b, j] = O.compute_mask_iou(rt[p["frame_idx"]], pm)
    return rows


@pytest.mark.parametrize("case", sorted(GREEDY_CASES))
def test_greedy_state_matches_golden(golden, case):
    mode, kw, n_frames = GREEDY_CASES[case]
    masklets, _, _ = greedy_table(golden)
    mt = torch.from_numpy(masklets).float()
    for eid in (("0", "1") if mode == "gdino" else (None,)):
        key = case if eid is None else f"{case}_exp{eid}"
        exp = golden.greedy[key]
        prompts = greedy_prompts(golden)
        st = dedup.GreedyState(prompts, n_frames, mode=mode, bin_size=4, expression_id=eid, **kw)
        while (batch := st.next_batch()) is not None:
            st.apply_iou_rows(batch, _oracle_iou_rows(mt, batch, prompts))
        res = st.result()
        for k in ("tracked", "filtered", "batches"):
            assert res[k] == exp[k], (key, k)
        for k in ("n_tracked", "n_filtered", "n_not_used"):
            if k in exp:
                assert res[k] == exp[k], (key, k)
        if "filtered_by" in exp:
            assert {str(a): b for a, b in res["filtered_by"].items()} == exp["filtered_by"]
        if "filtered_iou" in exp:
            assert {str(a): b for a, b in res["filtered_iou"].items()} == exp["filtered_iou"]      # float64 bit-equal


def test_greedy_state_random_tables_vs_oracle():
    """Random IoU tables (no pixels): the state machine and the restated script loops must agree on every field.
    The oracle loops are driven through their `impl` hook: a 'masklet' is a (T,1,1) tensor holding the member id, a
    prompt 'segmentation' is a 1x1 array holding the candidate id, and compute_mask_iou looks the pair up."""
    rng = np.random.default_rng(7)
    for trial in range(80):
        n = int(rng.integers(1, 40))
        T = int(rng.choice([20, 40, 260]))                       # every frame_idx below stays inside the masklet (out of range raises, as in the reference)
        bin_size = int(rng.choice([1, 4]))
        frame_idx = rng.integers(0, 5, n) * int(rng.choice([1, 2, 4]))
        iou = rng.random((n, n))
        iou[rng.random((n, n)) < 0.1] = 0.7                      # exactly-at-threshold entries must NOT suppress
        stab = rng.random(n) * 0.3 + 0.7
        stab[rng.random(n) < 0.1] = np.nan
        stab[rng.random(n) < 0.1] = 0.85
        mode = "grid" if trial % 2 == 0 else "gdino"
        kw = dict(bin_size=bin_size, n_max_tracks=int(rng.choice([3, 16, 64])), batch_size=int(rng.choice([1, 2, 4])), miou_thresh=0.7)

        def mk():
            return [{"prompt_id": k, "frame_idx": int(frame_idx[k]), "segmentation": np.full((1, 1), k, np.uint8),
                     "expression_id": "e", "stability_score": float(stab[k])} for k in range(n)]

        class TableImpl:
            reshape_masklet = staticmethod(lambda m: m)
            compute_mask_iou = staticmethod(lambda a, b: float(iou[int(a[0, 0]), int(b[0, 0])]))

        track_fn = lambda frame, batch: {p["prompt_id"]: torch.full((32, 1, 1), float(p["prompt_id"])) for p in batch}
        if mode == "grid":
            ro = GO.grid_greedy(mk(), T, track_fn, impl=TableImpl, **kw)
        else:
            ro = GO.gdino_greedy(mk(), "e", T, track_fn, stability_score_thresh=0.85, impl=TableImpl, **kw)
        st = dedup.GreedyState(mk(), T, mode=mode, expression_id="e", stability_score_thresh=0.85, **kw)
        while (batch := st.next_batch()) is not None:
            st.apply_iou_rows(batch, iou[batch])
        rs = st.result()
        for k in ("tracked", "filtered", "batches", "n_tracked", "n_filtered", "filtered_by", "filtered_iou", "not_used"):
            assert rs[k] == ro[k], (trial, mode, k)


def test_count_formulas_match_oracle():
    rng = np.random.default_rng(11)
    pred = (rng.random((9, 30, 45)) > 0.5).astype(np.uint8)
    gt = (rng.random((9, 30, 45)) > 0.4).astype(np.uint8)
    pred[2] = 0; gt[2] = 0; pred[3] = 0; gt[4] = 0
    c = O.jf_counts_exact(pred, gt)
    pt, gtt = torch.from_numpy(pred).float(), torch.from_numpy(gt).float()
    assert evaluator.J_from_counts(*c) == O.compute_J(pt, gtt)
    assert evaluator.F_from_counts(*c) == O.compute_F(pt, gtt)
    assert evaluator.F_from_counts(*O.jf_counts_exact(np.zeros_like(gt), gt)) == 0.0
    for got, exp in zip(utils.mask_metrics_from_counts(*c), O.compute_mask_metrics(pt, gtt, "none")):
        assert torch.equal(got, exp)
    rows = dedup.iou_from_counts([0, 5, 3], [0, 5, 4], [0, 5, 6])
    assert rows.tolist() == [1.0, 1.0, 3 / 7]


def test_metric_mirror(golden):
    gt_ids, corr = [3, 5, 9], [3, 3, 5, 5, 7, 9]
    preds, labels = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0, 0.0]), torch.tensor([1, 1, 0, 1, 1, 0])
    assert metric.recall_per_track(gt_ids, preds, labels, corr) == golden["x1_recall_per_track"].tolist()
    assert metric.recall_per_exp(gt_ids, preds, labels, corr) == golden["x1_recall_per_exp"][0]
    assert metric.recall_per_track(gt_ids, preds, labels, corr) == O.recall_per_track(gt_ids, preds, labels, corr)


def test_shard_partitions():
    for n, world in ((10, 1), (10, 2), (7, 4), (3, 8), (0, 2)):
        parts = [sharding.shard_indices(n, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(i % world == r for r, p in enumerate(parts) for i in p)       # the reference's modulo rule
    costs = [5, 1, 9, 3, 3, 7, 2, 8]
    parts = [sharding.shard_balanced(costs, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)


def test_pair_tile_partition_covers_upper_triangle_once():
    for n, world in ((256, 8), (64, 2), (130, 3), (1, 4)):
        tiles = sharding.pair_tile_owner(n, world)
        nt = (n + 63) // 64
        assert len(tiles) == nt * (nt + 1) // 2 and len({(a, b) for a, b, _ in tiles}) == len(tiles)
        assert all(a <= b for a, b, _ in tiles) and all(0 <= r < world for _, _, r in tiles)
        counts = np.bincount([r for _, _, r in tiles], minlength=world)
        assert counts.max() - counts.min() <= 1                                   # round-robin: balanced to within one tile


def test_word_slices_partition_the_word_axis():
    """K-split of BASELINE config 5: contiguous, 32-word aligned, covers [0, words) exactly once for any world size."""
    from sola_b200 import sharding
    for words in (4, 32, 100, 3240000, 12345676):
        for world in (1, 2, 3, 8):
            sl = sharding.word_slices(words, world)
            assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == words
            for (a, b), (c, d) in zip(sl, sl[1:]):
                assert b == c and a <= b
            assert all(a % 32 == 0 for a, _ in sl)
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 32 or words < 32 * world


def _bilinear_axis_fp32(dst, scale_f32, in_size):
    """ATen's (and csrc/resize_core.cuh's) source index: src = fmaf(dst + 0.5, scale, -0.5) clamped at 0; i1 = i0 + (i0 < in - 1).
    The product of a <=12-bit and a 24-bit significand is exact in float64, so one rounding to fp32 reproduces the fused multiply-add."""
    src = (np.asarray(dst, np.float64) + 0.5) * np.float64(scale_f32) - 0.5
    src = src.astype(np.float32)
    src = np.where(src >= 0, src, np.float32(0))
    i0 = src.astype(np.int64)
    return i0, i0 + (i0 < in_size - 1)


def test_r1_window_spans_at_most_three_words_below_scale_2():
    """Design invariant behind the branch-free phase A of R1 (resize_core.cuh, `win3`): for any scale factor < 1.99 the source pixels
    read by one 32-pixel output word lie in at most 3 consecutive source words — checked with the kernel's own fp32 index arithmetic
    over every width pair of a sweep that includes the BASELINE shapes."""
    wide = [540, 960, 854, 1280, 1920, 480, 720, 1080, 333, 1000, 2048]
    worst = 0
    for ow in list(range(1, 160, 3)) + wide:
        if ow < 160:
            widths = range(max(1, ow // 3), int(ow * 1.99) + 2)
        else:
            widths = sorted({max(1, int(ow * r)) for r in (0.3, 0.5, 0.889, 1.0, 4 / 3, 1.5, 1.9, 1.98, 1.989)})
        for W in widths:
            scale = np.float32(W) / np.float32(ow)
            if not scale < np.float32(1.99):
                continue
            c = np.arange((ow + 31) // 32)
            xa = _bilinear_axis_fp32(c * 32, scale, W)[0]
            xb = _bilinear_axis_fp32(np.minimum(c * 32 + 31, ow - 1), scale, W)[1]
            worst = max(worst, int(((xb >> 5) - (xa >> 5)).max()))
    assert worst <= 2


def test_bf16_packed_compare_bit_tricks():
    """Arithmetic identities the bf16 K1 path relies on (csrc/pack_core.cuh), emulated with numpy integers:
    (1) subtracting HSET2 bit masks (0xFFFF per passing half) from a 32-bit accumulator and decoding a + b as 2a + ((acc - a) >> 16);
    (2) PRMT 0x6420 + two masks + multiply by 0x01010101 puts the 8 predicate bits of a 16-byte vector, in element order, in the top byte;
    (3) flooring a threshold to bf16 does not change `x > t` for any bf16 x."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 400))
        lo, hi = rng.random(n) < rng.random(), rng.random(n) < rng.random()
        acc = np.uint32(0)
        for l, h in zip(lo, hi):
            mask = np.uint32((0xFFFF if l else 0) | ((0xFFFF if h else 0) << 16))
            acc = np.uint32((int(acc) - int(mask)) & 0xFFFFFFFF)
        acc_i = int(acc) - (1 << 32) if int(acc) >= (1 << 31) else int(acc)
        a = acc_i & 0xFFFF
        assert 2 * a + ((acc_i - a) >> 16) == int(lo.sum() + hi.sum())
    for _ in range(500):
        e = rng.random(8) < 0.5
        m = [(0xFFFF if e[2 * k] else 0) | ((0xFFFF if e[2 * k + 1] else 0) << 16) for k in range(4)]
        byte = lambda w, i: (w >> (8 * i)) & 0xFF
        prmt = lambda x, y: byte(x, 0) | (byte(x, 2) << 8) | (byte(y, 0) << 16) | (byte(y, 2) << 24)      # selector 0x6420
        t = (prmt(m[0], m[1]) & 0x08040201) | (prmt(m[2], m[3]) & 0x80402010)
        top = ((t * 0x01010101) & 0xFFFFFFFF) >> 24
        assert top == sum(int(b) << k for k, b in enumerate(e))
    # (3): every finite bf16 x against thresholds around the interesting values
    bits = np.arange(0, 1 << 16, dtype=np.uint32)
    x = (bits << 16).view(np.float32)
    x = x[np.isfinite(x)]
    for t in (0.0, 1.0, -1.0, 0.3, -0.7, 1.0000001, -1.0000001, 1e-40, -1e-40, 3.3e38, -3.3e38):
        t32 = np.float32(t)
        b = int(np.array([t32]).view(np.uint32)[0])
        h = b >> 16
        if (b & 0xFFFF) and (b >> 31):
            h += 1
        t_floor = np.array([(h & 0xFFFF) << 16], dtype=np.uint32).view(np.float32)[0]
        np.testing.assert_array_equal(x > t32, x > t_floor)


def test_fused_band_ownership_partitions_the_input_rows():
    """Design invariant of the fused K1+R1 kernel (csrc/fused_pack_resize.cu): band b of 32 output rows owns input rows
    [tlo, own_end) — every input row is owned by exactly one band (so the packed planes and the stability counts are written /
    counted once) and every row a band's outputs read lies inside the rows it streams, for up- and down-scaling alike."""
    TR = 32

    def i0_i1(dst, scale, n):
        src = np.float32((np.float64(dst) + 0.5) * np.float64(scale) - 0.5)
        src = src if src >= 0 else np.float32(0)
        a = int(src)
        return a, a + (1 if a < n - 1 else 0)

    for H in list(range(1, 100, 2)) + [180, 480, 540, 720, 1080, 1280, 2160]:
        for oh in (1, 5, 31, 32, 33, 64, 100, 540, 960):
            sy = np.float32(H) / np.float32(oh)
            nb = (oh + TR - 1) // TR
            cover = np.zeros(H, int)
            for b in range(nb):
                oy0 = b * TR
                nrows = min(TR, oh - oy0)
                ylo, yhi = i0_i1(oy0, sy, H)[0], i0_i1(oy0 + nrows - 1, sy, H)[1]
                tlo = 0 if b == 0 else ylo
                own_end = H if b == nb - 1 else i0_i1(oy0 + TR, sy, H)[0]
                yend = max(yhi, own_end - 1)
                assert tlo <= own_end
                cover[tlo:own_end] += 1
                for oy in (oy0, oy0 + nrows - 1):
                    a, b1 = i0_i1(oy, sy, H)
                    assert tlo <= a and b1 <= yend
            assert (cover == 1).all(), (H, oh)


def test_carry_save_accumulation_is_exact():
    """K2's Harley-Seal step (csrc/pair_iou.cu csa_quad): ones/twos bit-sliced counters + popcount of the weight-4 carry word give
    exactly sum(popcount(a & b)) over any number of k-quads."""
    rng = np.random.default_rng(1)
    maj = lambda a, b, c: (a & b) | (a & c) | (b & c)
    pop = lambda v: bin(int(v)).count("1")
    for _ in range(50):
        nq = int(rng.integers(1, 40))
        a = rng.integers(0, 1 << 32, size=(nq, 4), dtype=np.uint64)
        b = rng.integers(0, 1 << 32, size=(nq, 4), dtype=np.uint64) & rng.integers(0, 1 << 32, size=(nq, 4), dtype=np.uint64)
        ones = twos = 0
        acc = 0
        for q in range(nq):
            x = [int(a[q, k] & b[q, k]) for k in range(4)]
            t1, s1 = maj(ones, x[0], x[1]), ones ^ x[0] ^ x[1]
            t2, ones = maj(s1, x[2], x[3]), s1 ^ x[2] ^ x[3]
            f, twos = maj(twos, t1, t2), twos ^ t1 ^ t2
            acc += 4 * pop(f)
        assert acc + 2 * pop(twos) + pop(ones) == sum(pop(int(u & v)) for u, v in zip(a.ravel(), b.ravel()))


def test_prompt_frame_idx_out_of_range_raises():
    """masklets[prompt_id][frame_idx] raises in the reference (generate_tokens_grid.py:273); the kernel must never be asked to compare
    against a clamped, wrong frame."""
    import pytest
    prompts = [{"prompt_id": 0, "frame_idx": 0}, {"prompt_id": 1, "frame_idx": 8}]
    with pytest.raises(IndexError):
        dedup.GreedyState(prompts, 8, mode="grid")
    with pytest.raises(IndexError):
        dedup.GreedyState([{"prompt_id": 0, "frame_idx": -4}], 8, mode="grid")
    dedup.GreedyState(prompts, 9, mode="grid")


def test_label_metrics_from_counts_match_reference_formulas():
    """dedup.label_metrics_from_counts (the host half of the label step, generate_tokens_grid.py:253-264) against
    oracle.compute_mask_metrics on every (track, GT object) pair, incl. the four empty-case rules."""
    rng = np.random.default_rng(11)
    N, G, T, H, W = 4, 3, 6, 12, 20
    tracks = rng.random((N, T, H, W)) > 0.6
    gts = rng.random((G, T, H, W)) > 0.7
    tracks[0, 1] = False; gts[0, 1] = False            # both empty
    tracks[1, 2] = False                               # pred empty, gt not
    gts[1, 3] = False                                  # gt empty, pred not
    inter = (tracks[:, None] & gts[None]).sum((-2, -1)).astype(np.int32)
    lab = dedup.label_metrics_from_counts(inter, tracks.sum((-2, -1)).astype(np.int32), gts.sum((-2, -1)).astype(np.int32))
    for i in range(N):
        for g in range(G):
            p, r, u = O.compute_mask_metrics(torch.from_numpy(tracks[i]).float(), torch.from_numpy(gts[g]).float())
            assert (float(p), float(r), float(u)) == (float(lab["precision"][i, g]), float(lab["recall"][i, g]), float(lab["iou"][i, g]))


def test_sweep_metrics_vectorised_equal_the_per_unit_formulas():
    """evaluator.sweep_metrics_from_counts (all units of a sweep in one pass) is bit-equal to J_from_counts / F_from_counts /
    F_boundary_from_counts unit by unit — empty unions, empty boundaries, zero volumes, empty units, volumes above 2**24."""
    from sola_b200 import evaluator as E
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(1, 30))
        frames = rng.integers(0, 150, size=n) if trial % 3 else rng.integers(1, 9, size=n)
        offs = np.concatenate([[0], np.cumsum(frames)[:-1]])
        tot = int(frames.sum())
        hi = 2_000_000 if trial % 5 == 0 else 50
        pa, ga = rng.integers(0, hi, tot), rng.integers(0, hi, tot)
        inter = (np.minimum(pa, ga) * rng.random(tot)).astype(np.int64)
        for arr in (pa, ga):
            z = rng.random(tot) < 0.2
            arr[z] = 0
            inter[z] = 0
        a, b = rng.integers(0, 300, tot), rng.integers(0, 300, tot)
        a[rng.random(tot) < 0.2] = 0
        b[rng.random(tot) < 0.2] = 0
        c = np.stack([inter, pa, ga, a, b, (a * rng.random(tot)).astype(np.int64), (b * rng.random(tot)).astype(np.int64)]).astype(np.int32)
        if n > 2 and trial % 4 == 0:
            c[:, offs[1]: offs[1] + frames[1]] = 0                  # a unit with tp == 0 and empty unions everywhere
        J, F, Fb, totals = E.sweep_metrics_from_counts(c, offs, frames, True)
        assert np.array_equal(totals, c[:3].sum(axis=1, dtype=np.int64))
        for u in range(n):
            o, t = int(offs[u]), int(frames[u])
            if t == 0:
                continue
            sl = c[:, o: o + t]
            assert J[u] == E.J_from_counts(sl[0], sl[1], sl[2])
            assert F[u] == E.F_from_counts(sl[0], sl[1], sl[2])
            assert Fb[u] == E.F_boundary_from_counts(sl[3], sl[4], sl[5], sl[6])
