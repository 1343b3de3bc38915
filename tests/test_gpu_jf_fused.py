"""GPU parity: the fused J&F kernel (csrc/jf_fused.cu) — region counts vs oracle.jf_counts_exact (the exact twin of
evaluator.py:227-247) and boundary counts vs oracle/boundary_oracle.py (DAVIS definition; parity unpinned by the reference) —
on the BASELINE config 1 / 4 / 5 shapes (854-wide pad bits!), odd shapes, unaligned planes, multi-band frames and mixed-shape sweeps."""
import numpy as np
import pytest
import torch

from oracle import boundary_oracle as BO
from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


def _expected(pred, gt, bound_th=0.008, with_boundary=True):
    T = pred.shape[0]
    out = np.zeros((7, T), np.int64)
    out[:3] = np.stack(O.jf_counts_exact(pred, gt))
    if with_boundary:
        for t in range(T):
            out[3:, t] = BO.boundary_counts(pred[t], gt[t], bound_th)
    return out


def _pair(T, H, W, seed, flip=0.003, empty=1):
    from sola_b200 import synth
    pred, gt = synth.jf_pair(T, H, W, seed=seed, device="cpu", flip=flip, empty_frames=min(empty, T - 1) if T > 1 else 0)
    pred, gt = pred.numpy().copy(), gt.numpy().copy()
    pred[0, -3:, -5:] = 1                                # last row / last column / corner rules
    gt[0, :2, :2] = 1
    return pred, gt


# config 1 (480x854, r = 8, W % 32 = 22), config 4 shapes (360p ... 720p), config 5 (1080p, r = 18, 7 row bands), plus awkward ones:
# frame_words % 4 != 0 (54x96), one word column (100x33 -> 2), one row, one column, tall-narrow
SHAPES = [(3, 480, 854), (2, 720, 1280), (1, 1080, 1920), (2, 360, 640), (3, 54, 96), (2, 100, 33), (2, 35, 1000), (4, 48, 85),
          (3, 1, 70), (3, 70, 1), (2, 300, 40), (5, 7, 5)]


@pytest.mark.parametrize("T,H,W", SHAPES)
def test_fused_counts_vs_oracles(T, H, W):
    import sola_b200 as S
    from sola_b200 import packed as P
    pred, gt = _pair(T, H, W, seed=T * 1000 + H + W)
    if T > 2:
        pred[2] = 0                                      # n_fg == 0, n_gt > 0 rule; tp == 0 for that frame
    c = P.jf_boundary_counts(S.pack_masks(pred), S.pack_masks(gt)).cpu().numpy()
    assert np.array_equal(c, _expected(pred, gt)), (T, H, W)


@pytest.mark.parametrize("flip", [0.0, 0.02, 0.3])
def test_fused_dense_and_identical(flip):
    """flip = 0: pred == gt (every boundary pixel matches at distance 0: the early exit); 0.02 / 0.3: boundary pixels in nearly every word
    (the work queues run full)."""
    import sola_b200 as S
    from sola_b200 import packed as P
    pred, gt = _pair(2, 200, 333, seed=77, flip=flip, empty=0)
    c = P.jf_boundary_counts(S.pack_masks(pred), S.pack_masks(gt)).cpu().numpy()
    assert np.array_equal(c, _expected(pred, gt))
    if flip == 0.0:
        same = P.jf_boundary_counts(S.pack_masks(gt), S.pack_masks(gt)).cpu().numpy()
        assert np.array_equal(same[5], same[3]) and np.array_equal(same[6], same[4]) and np.array_equal(same[0], same[1])


def test_fused_radius_sweep_and_limits():
    import sola_b200 as S
    from sola_b200 import _lib, packed as P
    rng = np.random.default_rng(0)
    seg = np.zeros((1, 70, 90), np.uint8)
    seg[0, 20:50, 25:70] = 1
    seg[0] ^= (rng.random((70, 90)) > 0.98).astype(np.uint8)
    gt = np.roll(seg, 3, axis=2)
    for bound in (1, 2, 3, 5, 13, 31):
        c = P.jf_boundary_counts(S.pack_masks(seg), S.pack_masks(gt), bound_th=bound).cpu().numpy()
        assert np.array_equal(c, _expected(seg, gt, bound_th=bound)), bound
    with pytest.raises(_lib.SolaError):
        P.jf_boundary_counts(S.pack_masks(seg), S.pack_masks(gt), bound_th=40)


def test_fused_region_only_and_unaligned_planes():
    """radius < 0 (what Evaluator.compute_JF_metrics uses) and plane bases that are not 16-byte aligned (plain-load staging)."""
    import sola_b200 as S
    from sola_b200 import packed as P
    pred, gt = _pair(4, 54, 96, seed=5)                  # frame_words = 162
    pp, gp = S.pack_masks(pred), S.pack_masks(gt)
    c = P.jf_boundary_counts(pp, gp, with_boundary=False).cpu().numpy()
    assert np.array_equal(c, _expected(pred, gt, with_boundary=False))
    # shift both buffers by one word: frames 1.. of a (T+1)-frame buffer viewed from word 1
    def shifted(pm):
        flat = torch.zeros(pm.words.numel() + 1, dtype=torch.int32, device=pm.words.device)
        flat[1:] = pm.words.reshape(-1)
        v = flat[1:].view(pm.words.shape)
        assert v.data_ptr() % 16 != 0
        return P.PackedMasks(v, pm.H, pm.W)
    c2 = P.jf_boundary_counts(shifted(pp), shifted(gp)).cpu().numpy()
    assert np.array_equal(c2, _expected(pred, gt))


@pytest.mark.parametrize("T,H,W", SHAPES + [(2, 1080, 1920), (40, 9, 40), (3, 2160, 3840)])
def test_region_build_vs_oracle(T, H, W):
    """Sweeps without a boundary unit run the lean region-only build (small tiles, many CTAs per SM, 128-bit shared loads with scalar
    head / tail words): every shape class, aligned and word-shifted planes, one unit and the same unit inside a mixed sweep."""
    import sola_b200 as S
    from sola_b200 import packed as P
    pred, gt = _pair(T, H, W, seed=T * 7 + H + W)
    want = _expected(pred, gt, with_boundary=False)
    pp, gp = S.pack_masks(pred), S.pack_masks(gt)
    assert np.array_equal(P.jf_boundary_counts(pp, gp, with_boundary=False).cpu().numpy(), want)

    def shifted(pm, k):
        flat = torch.zeros(pm.words.numel() + k, dtype=torch.int32, device=pm.words.device)
        flat[k:] = pm.words.reshape(-1)
        return P.PackedMasks(flat[k:].view(pm.words.shape), pm.H, pm.W)
    for k in (1, 3):
        assert np.array_equal(P.jf_boundary_counts(shifted(pp, k), shifted(gp, k), with_boundary=False).cpu().numpy(), want), k
    other = _pair(3, 48, 85, seed=3)
    plan = P.JFSweepPlan([(S.pack_masks(other[0]), S.pack_masks(other[1])), (pp, gp), (shifted(pp, 2), gp)], with_boundary=False)
    assert len(plan.launches) == 1
    c = plan.run().cpu().numpy()
    assert np.array_equal(c[:, plan.offsets[0]: plan.offsets[0] + 3], _expected(*other, with_boundary=False))
    for u in (1, 2):
        assert np.array_equal(c[:, plan.offsets[u]: plan.offsets[u] + T], want)


def test_fused_mixed_shape_sweep_one_launch():
    """A MeViS-like sweep: units of different (T, H, W) in ONE launch; per-unit slices equal the per-unit oracles, and JFSweep's J / F /
    F_boundary equal the reference formulas on them."""
    import sola_b200 as S
    from sola_b200 import evaluator, packed as P
    shapes = [(5, 360, 640), (3, 480, 854), (2, 720, 1280), (4, 54, 96), (1, 1080, 1920), (6, 48, 85)]
    units = [_pair(T, H, W, seed=31 + k) for k, (T, H, W) in enumerate(shapes)]
    launches0 = S.launch_count()
    plan = P.JFSweepPlan([(S.pack_masks(p), S.pack_masks(g)) for p, g in units], with_boundary=True)
    base = S.launch_count()
    counts = plan.run().cpu().numpy()
    assert S.launch_count() - base == len(plan.launches) == 1, "the whole mixed-shape sweep must be one kernel launch"
    forced = P.JFSweepPlan([(S.pack_masks(p), S.pack_masks(g)) for p, g in units], with_boundary=True, ctas_per_sm=3)     # other tile class, same counts
    assert np.array_equal(forced.run().cpu().numpy(), counts)
    assert counts.shape == (7, sum(s[0] for s in shapes))
    for k, (p, g) in enumerate(units):
        o, T = plan.offsets[k], plan.frames[k]
        assert np.array_equal(counts[:, o:o + T], _expected(p, g)), shapes[k]
    sweep = evaluator.JFSweep(with_boundary=True)
    for k, (p, g) in enumerate(units):
        sweep.add(k, p, g)
    sweep.add("none", None, units[0][1])
    results, totals = sweep.finish()
    assert results[-1][1] == {"J": 0.0, "F": 0.0, "JF": 0.0}
    for k, (p, g) in enumerate(units):
        rec = results[k][1]
        assert rec["J"] == O.compute_J(torch.from_numpy(p).float(), torch.from_numpy(g).float())
        assert abs(rec["F"] - O.compute_F(torch.from_numpy(p).float(), torch.from_numpy(g).float())) < 1e-6      # north-star tolerance
        assert abs(rec["F_boundary"] - BO.boundary_f_masklet(p, g)) < 1e-12
    exp_tot = sum(np.stack(O.jf_counts_exact(p, g)).sum(1) for p, g in units)
    assert np.array_equal(totals, exp_tot)


def test_compute_JF_all_matches_separate_paths():
    from sola_b200 import evaluator
    pred, gt = _pair(6, 120, 214, seed=9)
    J, F, Fb = evaluator.compute_JF_all(pred, gt)
    assert J == evaluator.compute_J(pred, gt) and F == evaluator.compute_F(pred, gt)
    assert abs(Fb - BO.boundary_f_masklet(pred, gt)) < 1e-12
