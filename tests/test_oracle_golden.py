"""CPU: the restated oracle against the golden vectors produced by the unmodified reference (oracle/gen_golden.py),
and — when /root/reference is mounted — against the live reference functions on fresh random inputs."""
import warnings

import numpy as np
import pytest
import torch

from oracle import boundary_oracle as BO
from oracle import greedy_oracle as GO
from oracle import maskpath_oracle as O
from oracle import ref_shim as R
from conftest import GREEDY_CASES, greedy_prompts, greedy_table


def _f(x):
    return torch.from_numpy(np.ascontiguousarray(x)).float()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_stability_and_binarise(golden, tag):
    logits = golden[f"stab_{tag}_logits"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array([O.get_stability_score(l) for l in logits])
        got2 = np.array([O.get_stability_score(l, 0.5, 0.25) for l in logits])
    np.testing.assert_array_equal(got, golden[f"stab_{tag}_score"])          # bit-exact incl. nan positions
    np.testing.assert_array_equal(got2, golden[f"stab_{tag}_score_t05_o025"])
    assert np.isnan(golden[f"stab_{tag}_score"]).sum() == 1                   # the all-negative plane
    packed = O.pack_bits(O.binarize(torch.from_numpy(logits)).numpy())
    np.testing.assert_array_equal(packed, golden[f"stab_{tag}_binarized_packed"])
    # batched form agrees with the per-plane form
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_array_equal(O.get_stability_score(logits), got)


def test_pack_roundtrip():
    rng = np.random.default_rng(0)
    for W in (1, 31, 32, 33, 70, 854):
        m = rng.random((3, 5, W)) > 0.5
        p = O.pack_bits(m)
        assert p.shape == (3, 5, (W + 31) // 32) and p.dtype == np.uint32
        np.testing.assert_array_equal(O.unpack_bits(p, W), m.astype(np.uint8))
        if W % 32:
            assert not (p[..., -1] >> (W % 32)).any()                          # pad bits are zero


def test_iou_family(golden):
    H, W = golden.meta["iou_shape"]
    A, B = _f(golden.masks("iou_A", W)), _f(golden.masks("iou_B", W))
    got = np.array([O.compute_mask_iou(a, b) for a, b in zip(A, B)])
    np.testing.assert_array_equal(got, golden["iou_mask_iou"])
    exp = golden["iou_mask_iou_torch"]
    for k, (a, b) in enumerate(zip(A, B)):
        if np.isnan(exp[k]):
            with pytest.raises(ZeroDivisionError):
                O.compute_mask_iou_torch(a, b)
        else:
            assert O.compute_mask_iou_torch(a, b) == exp[k]
        i, na, nb = O.iou_counts_exact(a.numpy(), b.numpy())
        assert O.iou_from_counts(i, na, nb) == golden["iou_mask_iou"][k]       # integer truth reproduces the fp32 path
    got = np.array([O.compute_masklet_iou(A[:6], B[:6]), O.compute_masklet_iou(A[6:7], B[6:7])])
    np.testing.assert_array_equal(got, golden["iou_masklet_iou"])


def test_mask_metrics_and_jf(golden):
    H, W = golden.meta["iou_shape"]
    pred, gt = _f(golden.masks("mm_pred", W)), _f(golden.masks("mm_gt", W))
    p, r, i = O.compute_mask_metrics(pred, gt, "none")
    np.testing.assert_array_equal(torch.stack([p, r, i]).numpy(), golden["mm_none"])
    p, r, i = O.compute_mask_metrics(pred, gt)
    np.testing.assert_array_equal(torch.stack([p, r, i]).numpy(), golden["mm_mean"])
    with pytest.raises(ValueError):
        O.compute_mask_metrics(pred, gt, "sum")
    c = O.jf_counts_exact(pred.numpy(), gt.numpy())
    p2, r2, i2 = O.mask_metrics_from_counts(*c)
    np.testing.assert_array_equal(torch.stack([p2, r2, i2]).numpy(), golden["mm_none"])
    z = torch.zeros_like(gt)
    J = np.array([O.compute_J(pred, gt), O.compute_J(z, gt), O.compute_J(gt, gt)])
    F = np.array([O.compute_F(pred, gt), O.compute_F(z, gt), O.compute_F(gt, gt)])
    np.testing.assert_array_equal(J, golden["jf_J"])
    np.testing.assert_array_equal(F, golden["jf_F"])
    assert O.J_from_counts(*c) == golden["jf_J"][0]
    assert O.F_from_counts(*c) == golden["jf_F"][0]
    assert O.F_from_counts(*O.jf_counts_exact(z.numpy(), gt.numpy())) == 0.0     # tp == 0 rule


def test_resizes(golden):
    land, port, sq = _f(golden.masks("rs_land_in", 128)), _f(golden.masks("rs_port_in", 72)), _f(golden.masks("rs_sq_in", 64))
    np.testing.assert_array_equal(O.pack_bits(O.reshape_masklet(land).numpy()), golden["rs_land_out"])
    np.testing.assert_array_equal(O.pack_bits(O.reshape_masklet(port).numpy()), golden["rs_port_out"])
    np.testing.assert_array_equal(O.pack_bits(O.reshape_masklet(land, (45, 80)).numpy()), golden["rs_land_out_45x80"])
    out = O.reshape_masklet(sq)
    assert tuple(out.shape) == (1, 960, 540)                                   # square goes to portrait (seg_utils.py:155)
    np.testing.assert_array_equal(O.pack_bits(out.numpy()), golden["rs_sq_out"])
    near = O.resize_prompt_nearest(land[0].numpy().astype(np.uint8), 540, 960)
    np.testing.assert_array_equal(O.pack_bits(near.numpy()), golden["rs_nearest_out"])
    # index restatement used by the GPU kernel's documentation
    ys, xs = O.nearest_source_index(540, 72), O.nearest_source_index(960, 128)
    np.testing.assert_array_equal(land[0].numpy()[ys][:, xs], near.numpy())


def test_partness_and_recall(golden):
    parts, full = _f(golden.masks("P_parts", 64)), _f(golden.masks("P_full", 64))
    np.testing.assert_array_equal(O.compute_P(parts, full).numpy(), golden["P_out"])
    gt_ids, corr = [3, 5, 9], [3, 3, 5, 5, 7, 9]
    preds, labels = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0, 0.0]), torch.tensor([1, 1, 0, 1, 1, 0])
    np.testing.assert_array_equal(np.array(O.recall_per_track(gt_ids, preds, labels, corr)), golden["x1_recall_per_track"])
    assert O.recall_per_exp(gt_ids, preds, labels, corr) == golden["x1_recall_per_exp"][0]


@pytest.mark.parametrize("case", sorted(GREEDY_CASES))
def test_greedy_oracle_vs_golden(golden, case):
    mode, kw, n_frames = GREEDY_CASES[case]
    masklets, _, _ = greedy_table(golden)
    mt = _f(masklets)
    track_fn = lambda frame, batch: {p["prompt_id"]: mt[p["prompt_id"]] for p in batch}
    for eid in (("0", "1") if mode == "gdino" else (None,)):
        key = case if eid is None else f"{case}_exp{eid}"
        exp = golden.greedy[key]
        if mode == "grid":
            res = GO.grid_greedy(greedy_prompts(golden), n_frames, track_fn, bin_size=4, **kw)
        else:
            res = GO.gdino_greedy(greedy_prompts(golden), eid, n_frames, track_fn, bin_size=4, **kw)
        for k in ("tracked", "filtered", "batches"):
            assert res[k] == exp[k], (key, k)
        if "filtered_by" in exp:
            assert {str(a): b for a, b in res["filtered_by"].items()} == exp["filtered_by"]


def test_jf_sweep_and_merges():
    rng = np.random.default_rng(3)
    gt = (rng.random((4, 20, 30)) > 0.6).astype(np.uint8)
    pred = (rng.random((4, 20, 30)) > 0.6).astype(np.uint8)
    out, mJ, mF, mJF = O.jf_sweep([("v", "0", pred, gt), ("v", "1", None, gt)])
    assert out["v"]["1"] == {"J": 0.0, "F": 0.0, "JF": 0.0}
    assert mJ == (out["v"]["0"]["J"] + 0.0) / 2 and mJF == (out["v"]["0"]["JF"]) / 2
    tracks = [(rng.random((4, 20, 30)) > 0.8).astype(np.uint8) for _ in range(4)]
    m = O.merge_selected_tracks(tracks, [0, 1, 0, 1])
    np.testing.assert_array_equal(m, tracks[1] | tracks[3])
    assert O.merge_selected_tracks(tracks, [0, 0, 0, 0]).sum() == 0
    assert O.merge_selected_tracks([], []) is None
    np.testing.assert_array_equal(O.merge_gt_objects(tracks[:3]), tracks[0] | tracks[1] | tracks[2])


def test_stability_filter_rule():
    assert O.stability_filter_keep(4, 0.85, 4, 0.85)            # == thresh kept
    assert O.stability_filter_keep(8, float("nan"), 4, 0.85)    # nan kept
    assert not O.stability_filter_keep(8, 0.84, 4, 0.85)
    assert not O.stability_filter_keep(6, 0.99, 4, 0.85)        # off-bin dropped


def test_boundary_oracle_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    for (H, W) in ((48, 85), (64, 64)):
        seg = np.zeros((H, W), np.uint8)
        seg[10:30, 20:60] = 1
        seg[5:12, 3:9] = 1
        seg ^= (rng.random((H, W)) > 0.97).astype(np.uint8)
        b = BO.seg2bmap(seg)
        for r in (0, 1, 3, 8):
            L = np.arange(-r, r + 1)
            X, Y = np.meshgrid(L, L)
            disk = ((X ** 2 + Y ** 2) <= r ** 2).astype(np.uint8)
            ref = cv2.dilate(b.astype(np.uint8), disk) if r > 0 else b.astype(np.uint8)
            np.testing.assert_array_equal(BO.dilate_disk(b, r).astype(np.uint8), ref)
    assert BO.bound_pix_for(480, 854) == 8 and BO.bound_pix_for(720, 1280) == 12 and BO.bound_pix_for(1080, 1920) == 18
    # rule table
    assert BO.f_from_boundary_counts(0, 5, 0, 0) == 0.0 and BO.f_from_boundary_counts(0, 0, 0, 0) == 1.0
    assert BO.f_from_boundary_counts(4, 0, 0, 0) == 0.0 and BO.f_from_boundary_counts(4, 4, 4, 4) == 1.0
    same = np.zeros((40, 50), np.uint8)
    same[10:20, 10:30] = 1
    assert BO.boundary_f_frame(same, same) == 1.0


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_oracle_vs_live_reference():
    """Fresh random inputs, oracle vs the imported reference — every function of SURVEY.md §8(a) that imports."""
    g = torch.Generator().manual_seed(123)
    for trial in range(4):
        H, W = [(36, 70), (64, 96), (50, 33), (24, 128)][trial]
        a = (torch.rand((5, H, W), generator=g) > 0.55).float()
        b = (torch.rand((5, H, W), generator=g) > 0.45).float()
        if trial == 1:
            a[2] = 0; b[2] = 0
            a[3] = 0
        assert O.compute_mask_iou(a[0], b[0]) == R.compute_mask_iou(a[0], b[0])
        assert O.compute_mask_iou(a[2], b[2]) == R.compute_mask_iou(a[2], b[2])
        assert O.compute_masklet_iou(a, b) == R.compute_masklet_iou(a, b, "cpu")
        assert O.compute_mask_iou_torch(a[0], b[0]) == R.compute_mask_iou_torch(a[0], b[0])
        assert O.compute_J(a, b) == R.compute_J(a, b) and O.compute_F(a, b) == R.compute_F(a, b)
        for red in ("mean", "none"):
            for x, y in zip(O.compute_mask_metrics(a, b, red), R.compute_mask_metrics(a, b, red)):
                assert torch.equal(x, y)
        assert torch.equal(O.reshape_masklet(a), R.reshape_masklet(a))
        assert torch.equal(O.reshape_masklet(a, (30, 40)), R.reshape_masklet(a, (30, 40)))
        np.testing.assert_array_equal(O.compute_P(a, b[0]).numpy(), R.compute_P(a, b[0]).numpy())   # nan (empty part) == nan
        logit = (torch.randn((3, H, W), generator=g) * 2).numpy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np.testing.assert_array_equal(O.get_stability_score(logit), R.get_stability_score(logit))
            np.testing.assert_array_equal(O.get_stability_score(logit[0], 0.3, 0.7), R.get_stability_score(logit[0], 0.3, 0.7))
