"""GPU parity: K3 counts and every drop-in built on them (IoU family, mask metrics, J, F, OR-merge, sweep)."""
import numpy as np
import pytest
import torch

from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


def _f(x):
    return torch.from_numpy(np.ascontiguousarray(x)).float()


def test_golden_iou_family(golden):
    from sola_b200 import seg_utils, utils
    H, W = golden.meta["iou_shape"]
    A, B = _f(golden.masks("iou_A", W)), _f(golden.masks("iou_B", W))
    got = np.array([seg_utils.compute_mask_iou(a.cuda(), b.cuda()) for a, b in zip(A, B)])
    np.testing.assert_array_equal(got, golden["iou_mask_iou"])                   # float64 bit-equal
    assert isinstance(seg_utils.compute_mask_iou(A[0], B[0]), float)             # CPU tensors are accepted (H2D), same value
    assert seg_utils.compute_mask_iou(A[0], B[0]) == golden["iou_mask_iou"][0]
    exp = golden["iou_mask_iou_torch"]
    for k, (a, b) in enumerate(zip(A, B)):
        if np.isnan(exp[k]):
            with pytest.raises(ZeroDivisionError):
                utils.compute_mask_iou_torch(a.cuda(), b.cuda())
        else:
            assert utils.compute_mask_iou_torch(a.cuda(), b.cuda()) == exp[k]
    got = np.array([seg_utils.compute_masklet_iou(A[:6], B[:6], "cuda"), seg_utils.compute_masklet_iou(A[6:7], B[6:7], "cuda")])
    np.testing.assert_array_equal(got, golden["iou_masklet_iou"])


def test_golden_metrics_J_F(golden):
    from sola_b200 import utils, evaluator
    H, W = golden.meta["iou_shape"]
    pred, gt = _f(golden.masks("mm_pred", W)).cuda(), _f(golden.masks("mm_gt", W)).cuda()
    p, r, i = utils.compute_mask_metrics(pred, gt, "none")
    assert p.dtype == torch.float32 and not p.is_cuda
    np.testing.assert_array_equal(torch.stack([p, r, i]).numpy(), golden["mm_none"])
    p, r, i = utils.compute_mask_metrics(pred, gt)
    np.testing.assert_array_equal(torch.stack([p, r, i]).numpy(), golden["mm_mean"])
    with pytest.raises(ValueError):
        utils.compute_mask_metrics(pred, gt, "sum")
    ev = evaluator.Evaluator()
    z = torch.zeros_like(gt)
    J = np.array([ev.compute_J(pred, gt), ev.compute_J(z, gt), ev.compute_J(gt, gt)])
    F = np.array([ev.compute_F(pred, gt), ev.compute_F(z, gt), ev.compute_F(gt, gt)])
    np.testing.assert_array_equal(J, golden["jf_J"])
    np.testing.assert_array_equal(F, golden["jf_F"])
    assert isinstance(J[0], np.float64) and isinstance(ev.compute_F(pred, gt), float)


@pytest.mark.parametrize("shape", [(30, 480, 854), (5, 33, 47), (3, 720, 1280), (1, 7, 3), (4, 540, 960)])
@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_counts_vs_oracle(shape, dtype):
    import sola_b200 as S
    from sola_b200 import evaluator
    rng = np.random.default_rng(sum(shape))
    gt = rng.random(shape) > 0.5
    pred = gt ^ (rng.random(shape) > 0.9)
    pred[0] = False; gt[0] = False
    if shape[0] > 2:
        pred[1] = False
    cast = (lambda m: m.astype(np.float32)) if dtype == "f32" else (lambda m: m.astype(np.uint8))
    c = S.frame_counts(cast(pred), cast(gt)).cpu().numpy()
    exp = O.jf_counts_exact(pred, gt)
    for k in range(3):
        np.testing.assert_array_equal(c[k], exp[k])
    pt, gtt = _f(pred), _f(gt)
    J, F = evaluator.compute_JF(cast(pred), cast(gt))
    assert J == O.compute_J(pt, gtt)                                            # per-frame sums are exact in fp32 too
    assert abs(F - O.compute_F(pt, gtt)) < 1e-6                                 # reference fp32 volume sums may drift
    assert F == O.F_from_counts(*exp)


def test_unaligned_and_odd_sizes_take_scalar_path():
    import sola_b200 as S
    rng = np.random.default_rng(5)
    a = (rng.random((3, 11, 13)) > 0.5).astype(np.float32)        # frame_px = 143, not a multiple of 4
    b = (rng.random((3, 11, 13)) > 0.5).astype(np.float32)
    c = S.frame_counts(a, b).cpu().numpy()
    exp = O.jf_counts_exact(a, b)
    for k in range(3):
        np.testing.assert_array_equal(c[k], exp[k])
    base = torch.from_numpy(np.concatenate([[0.0], a.ravel()]).astype(np.float32)).cuda()
    a_off = base[1:].view(3, 11, 13)
    c = S.frame_counts(a_off, torch.from_numpy(b).cuda()).cpu().numpy()
    np.testing.assert_array_equal(c[0], exp[0])


@pytest.mark.parametrize("Na,Nb,T,H,W", [(3, 3, 2, 40, 128), (1, 1, 3, 16, 256), (7, 5, 2, 24, 96), (64, 3, 2, 54, 96), (2, 9, 1, 8, 130)])
def test_packed_batched_counts_track_tiles(Na, Nb, T, H, W):
    """The batched label counts tile the tracks (two per CTA) and the objects (four per CTA): odd counts on both axes, the 128-bit
    path (frame words a multiple of 4) and the scalar one, all-zero and all-one planes."""
    import sola_b200 as S
    rng = np.random.default_rng(Na * 100 + Nb)
    A = rng.random((Na, T, H, W)) > 0.55
    B = rng.random((Nb, T, H, W)) > 0.45
    A[0, 0] = True
    B[-1, -1] = False
    inter, area_a, area_b = S.frame_counts_packed(S.pack_masks(A), S.pack_masks(B))
    np.testing.assert_array_equal(inter.cpu().numpy(), (A[:, None] & B[None]).sum((-2, -1)))
    np.testing.assert_array_equal(area_a.cpu().numpy(), A.sum((-2, -1)))
    np.testing.assert_array_equal(area_b.cpu().numpy(), B.sum((-2, -1)))


def test_packed_batched_and_ragged_counts():
    import sola_b200 as S
    from sola_b200 import utils
    rng = np.random.default_rng(8)
    Na, Nb, T, H, W = 5, 6, 4, 45, 70
    A = rng.random((Na, T, H, W)) > 0.6
    B = rng.random((Nb, T, H, W)) > 0.5
    A[1, 2] = False; B[3, 2] = False; B[0] = False
    pa, pb = S.pack_masks(A), S.pack_masks(B)
    inter, area_a, area_b = S.frame_counts_packed(pa, pb)
    np.testing.assert_array_equal(inter.cpu().numpy(), (A[:, None] & B[None]).sum((-2, -1)))
    np.testing.assert_array_equal(area_a.cpu().numpy(), A.sum((-2, -1)))
    np.testing.assert_array_equal(area_b.cpu().numpy(), B.sum((-2, -1)))
    p, r, i = utils.compute_mask_metrics_batch(pa, pb)
    for ia in range(Na):
        for ib in range(Nb):
            e = O.compute_mask_metrics(_f(A[ia]), _f(B[ib]))
            assert torch.equal(torch.stack([p[ia, ib], r[ia, ib], i[ia, ib]]), torch.stack(list(e)))
    # ragged: three units of different shapes concatenated
    shapes = [(3, 20, 33), (2, 64, 64), (5, 9, 100)]
    aw, bw, offs, exp = [], [], [0], []
    for (t, h, w) in shapes:
        x, y = rng.random((t, h, w)) > 0.5, rng.random((t, h, w)) > 0.5
        px, py = O.pack_bits(x), O.pack_bits(y)
        aw.append(px.reshape(-1)); bw.append(py.reshape(-1))
        fw = px.shape[1] * px.shape[2]
        for k in range(t):
            offs.append(offs[-1] + fw)
        exp.append(O.jf_counts_exact(x, y))
    a_dev = torch.from_numpy(np.concatenate(aw).view(np.int32)).cuda()
    b_dev = torch.from_numpy(np.concatenate(bw).view(np.int32)).cuda()
    c = S.packed.frame_counts_ragged(a_dev, b_dev, torch.tensor(offs, dtype=torch.int64).cuda()).cpu().numpy()
    for k in range(3):
        np.testing.assert_array_equal(c[k], np.concatenate([e[k] for e in exp]))


def test_or_merge_matches_dataloader_rules():
    import sola_b200 as S
    from sola_b200 import dataloader_ops as D
    rng = np.random.default_rng(2)
    tracks = [(rng.random((4, 30, 50)) > 0.8).astype(np.uint8) for _ in range(5)]
    packed = S.pack_masks(np.stack(tracks))
    for preds in ([0, 1, 0, 1, 1], [1, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0.2, 1.0, 0.0, 2.0, 1.0]):
        # the reference selects with `preds[i] > 0` (dataloader.py:339); its `< 1` skip (:324) only saves file reads
        exp = np.asarray(O.merge_selected_tracks(tracks, [1 if q > 0 else 0 for q in preds])) != 0
        got = D.merge_selected_tracks(packed, preds)
        np.testing.assert_array_equal(S.unpack_masks(got, torch.uint8).cpu().numpy(), exp.astype(np.uint8))
    assert D.merge_selected_tracks(None, []) is None
    got = D.merge_gt_objects(packed[:3])
    np.testing.assert_array_equal(S.unpack_masks(got, torch.uint8).cpu().numpy(), (tracks[0] | tracks[1] | tracks[2]))


def test_jf_sweep_and_evaluator_json(tmp_path):
    """compute_JF_metrics end to end against the oracle sweep, through a stand-in for the reference's dataset object."""
    import json
    from sola_b200 import evaluator
    rng = np.random.default_rng(4)
    units = {}
    for v in range(2):
        for e in range(3):
            T, H, W = int(rng.integers(2, 6)), 36 + 4 * v, 50 + 3 * e
            gt = (rng.random((T, H, W)) > 0.5).astype(np.uint8)
            pred = None if (v, e) == (1, 1) else (gt ^ (rng.random((T, H, W)) > 0.85)).astype(np.uint8)
            if (v, e) == (0, 2):
                pred = np.zeros_like(gt)
            units[(f"v{v}", f"{e}")] = (pred, gt)

    class DS:
        def set_video(self, vid): self.vid = vid
        def get_gt_masklet(self, vid, eid): return units[(vid, eid)][1]
        def get_sam2_masklet(self, video_id, expression_id, preds, root_types, prompt_types, sam2_anno_ids): return units[(video_id, expression_id)][0]

    class Loader: dataset = DS()
    pred_dict = {}
    for (vid, eid) in units:
        pred_dict.setdefault(vid, {})[eid] = {"expression": f"exp {eid}", "pred": [1], "root_type": [], "prompt_type": [], "sam2_anno_id": []}
    ev = evaluator.Evaluator(loader_dict={"valid": Loader()}, pred_dict=pred_dict, eval_output_dir=str(tmp_path), eval_weight_epoch=3)
    ev.compute_JF_metrics()
    exp_out, mJ, mF, mJF = O.jf_sweep([(v, e, p, g) for (v, e), (p, g) in units.items()])
    assert ev.metrics["mean_J"] == mJ and abs(ev.metrics["mean_F"] - mF) < 1e-6 and abs(ev.metrics["mean_JF"] - mJF) < 1e-6
    saved = json.load(open(tmp_path / "valid_JF_metrics_3epoch.json"))
    for (v, e) in units:
        assert saved[v][e]["expression"] == f"exp {e}"
        assert saved[v][e]["J"] == exp_out[v][e]["J"]
        assert abs(saved[v][e]["F"] - exp_out[v][e]["F"]) < 1e-6
    assert saved["v1"]["1"] == {"expression": "exp 1", "J": 0.0, "F": 0.0, "JF": 0.0}


@pytest.mark.parametrize("shape", [(30, 480, 854), (5, 33, 47), (3, 720, 1280), (1, 7, 3)])
def test_jf_accumulators_abi(shape):
    """sola_jf_f32 / _u8 / _packed (the J&F entry SURVEY.md §8(b) names): per-frame inter / union and exact tp / fp / fn, and the
    (J, F) they give against the oracle's restatement of evaluator.py:227-247."""
    import sola_b200 as S
    from sola_b200 import evaluator, synth
    T, H, W = shape
    pred, gt = synth.jf_pair(T, H, W, seed=77 + T, device="cpu")
    p, g = pred.numpy().astype(bool), gt.numpy().astype(bool)
    inter_ref = (p & g).reshape(T, -1).sum(1)
    uni_ref = (p | g).reshape(T, -1).sum(1)
    tot_ref = [int((p & g).sum()), int((p & ~g).sum()), int((~p & g).sum())]
    pf, gf = pred.float(), gt.float()
    for a, b in ((pf.cuda(), gf.cuda()), (pred.cuda(), gt.cuda()), (S.pack_masks(pred), S.pack_masks(gt))):
        inter, uni, tot = S.packed.jf_accumulators(a, b)
        np.testing.assert_array_equal(inter.cpu().numpy(), inter_ref)
        np.testing.assert_array_equal(uni.cpu().numpy(), uni_ref)
        assert tot.cpu().tolist() == tot_ref
        J, F = evaluator.jf_from_accumulators(inter, uni, tot)
        assert J == O.compute_J(pf, gf)
        assert abs(F - O.compute_F(pf, gf)) < 1e-6            # north-star tolerance; equal whenever the reference's fp32 sums are exact
    # empty unit and all-empty prediction (tp == 0 -> F = 0; union == 0 frames -> J_t = 1)
    z = torch.zeros((4, 20, 40), dtype=torch.uint8, device="cuda")
    inter, uni, tot = S.packed.jf_accumulators(z, z)
    assert evaluator.jf_from_accumulators(inter, uni, tot) == (1.0, 0.0) and tot.cpu().tolist() == [0, 0, 0]


def test_compute_J_then_compute_F_share_one_pass():
    """evaluator.py:201-202 calls compute_J and compute_F on the same tensors: one kernel launch, and any in-place edit invalidates."""
    import sola_b200 as S
    from sola_b200 import evaluator, synth
    p, g = synth.jf_pair(5, 40, 70, seed=2, device="cuda")
    pf, gf = p.float(), g.float()
    n0 = S.launch_count()
    J = evaluator.compute_J(pf, gf)
    n1 = S.launch_count()
    F = evaluator.compute_F(pf, gf)
    assert S.launch_count() == n1 and n1 > n0
    assert J == O.compute_J(pf.cpu(), gf.cpu()) and abs(F - O.compute_F(pf.cpu(), gf.cpu())) < 1e-6
    assert pf[2].any()
    pf[2].zero_()                                                   # in-place write -> version bump -> recomputed
    J2 = evaluator.compute_J(pf, gf)
    assert S.launch_count() > n1 and J2 == O.compute_J(pf.cpu(), gf.cpu()) and J2 != J
