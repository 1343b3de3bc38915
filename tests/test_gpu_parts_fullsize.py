"""GPU: part-mask suppression (compute_P) and full-size BASELINE shapes through size-independent properties + oracle sub-samples."""
import warnings

import numpy as np
import pytest
import torch

from oracle import boundary_oracle as BO
from oracle import greedy_oracle as GO
from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


def _nested_masks(n, H, W, seed):
    """Area-sorted masks with genuine part-of relations: boxes inside boxes plus noise."""
    rng = np.random.default_rng(seed)
    m = np.zeros((n, H, W), np.float32)
    for k in range(n):
        if k % 3 == 0 or k < 2:
            y, x = rng.integers(0, H // 2), rng.integers(0, W // 2)
            h, w = rng.integers(H // 6, H // 2), rng.integers(W // 6, W // 2)
        else:                                                    # mostly inside an earlier mask
            ys, xs = np.nonzero(m[rng.integers(0, k)])
            y, x = ys.min(), xs.min()
            h, w = max(2, (ys.max() - y) // 2 + rng.integers(0, 6)), max(2, (xs.max() - x) // 2 + rng.integers(0, 6))
        m[k, y:y + h, x:x + w] = 1
    m[n - 1] = 0                                                 # an empty mask: P is nan, never a part
    order = np.argsort(-m.sum((1, 2)), kind="stable")
    return m[order]


def test_compute_P_and_part_suppression(golden):
    from sola_b200 import utils
    parts = torch.from_numpy(golden.masks("P_parts", 64)).float()
    full = torch.from_numpy(golden.masks("P_full", 64)).float()
    np.testing.assert_array_equal(utils.compute_P(parts.cuda(), full.cuda()).cpu().numpy(), golden["P_out"])      # exact GEMV values
    for seed in range(4):
        m = _nested_masks(14, 90, 160, seed)
        mt = torch.from_numpy(m)
        got = utils.suppress_part_masks(m, 0.7, autocast_bf16=False)
        np.testing.assert_array_equal(got, GO.part_suppression(mt, 0.7).numpy())
        assert got.any() and not got.all()
        # what the reference computes on the GPU: GEMV under bf16 autocast (generate_prompts_grid.py:59)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref_P = O.compute_P(mt.cuda(), mt[1].cuda())
        np.testing.assert_array_equal(utils.compute_P(m, m[1], autocast_bf16=True).cpu().numpy(), ref_P.float().cpu().numpy())
        ref_part = torch.tensor([False] * 14)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for k in range(13):
                if ref_part[k]:
                    continue
                Pk = O.compute_P(mt.cuda(), mt[k].cuda()).cpu()
                ref_part[Pk > 0.7] = True
                ref_part[k] = False
        np.testing.assert_array_equal(utils.suppress_part_masks(m, 0.7, autocast_bf16=True), ref_part.numpy())


def test_config1_jf_full_size_vs_oracle():
    """BASELINE config 1: 1 video x 1 expression x 30 frames x 480x854 — J, F (reference definitions) and boundary F."""
    from sola_b200 import evaluator, synth
    pred, gt = synth.jf_pair(30, 480, 854, seed=1235, device="cpu")
    ev = evaluator.Evaluator()
    pf, gf = pred.float(), gt.float()
    J, F = ev.compute_J(pf.cuda(), gf.cuda()), ev.compute_F(pf.cuda(), gf.cuda())
    assert J == O.compute_J(pf, gf)                               # bit-equal: per-frame counts are exact in the reference too
    assert abs(F - O.compute_F(pf, gf)) < 1e-6                    # stated tolerance (volume sums < 2^24 here, so in fact equal)
    assert F == O.compute_F(pf, gf)
    c = evaluator.jf_counts(pred, gt)                             # uint8 inputs straight from the dataloader
    for k, e in enumerate(O.jf_counts_exact(pred.numpy(), gt.numpy())):
        np.testing.assert_array_equal(c[k], e)
    fb = evaluator.compute_F_boundary(pred, gt)
    sub = [0, 7, 29]
    import sola_b200 as S
    bc = S.boundary_counts(S.pack_masks(pred), S.pack_masks(gt)).cpu().numpy()
    for t in sub:
        assert tuple(bc[:, t]) == BO.boundary_counts(pred[t].numpy(), gt[t].numpy())
    assert 0.0 <= fb <= 1.0


def test_config2_scale_properties():
    """Config-2-shaped pass at 16 tracks x 80 frames x 720x1280 (4.7 GB of logits): identities that hold at any size."""
    import sola_b200 as S
    from sola_b200 import dedup, synth
    N, T, H, W = 16, 80, 720, 1280
    logits, prompts = synth.dedup_candidates(N, T, H, W, seed=77, device="cuda")
    packed, counts = S.binarize_pack_stability(logits)
    c = counts.cpu().numpy()
    assert (c[0] <= c[1]).all() and (c[1] <= c[2]).all()                                   # nested thresholds
    inter = S.pairwise_inter_matrix(packed).cpu().numpy()
    assert np.array_equal(inter, inter.T)
    np.testing.assert_array_equal(np.diag(inter), c[1].sum(1, dtype=np.int64))             # checksum of checksums: K2 diag == K1 areas
    # |A∩B| + |A∪B| == |A| + |B| through an independent kernel (OR-merge + area)
    a, b = 3, 5
    union = S.or_merge(packed[[a, b]])
    _, u_area = S.pack_masks(S.unpack_masks(union, torch.uint8), want_area=True)
    assert int(u_area.sum()) + int(inter[a, b]) == int(inter[a, a]) + int(inter[b, b])
    # oracle on a sub-sample of planes and one stability frame
    sub = logits[2, 10:12].cpu()
    np.testing.assert_array_equal(packed.words[2, 10:12].cpu().numpy().view(np.uint32), O.pack_bits(sub.numpy() > 0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_array_equal(S.packed.stability_from_counts(counts)[2, 10:12], O.get_stability_score(sub.numpy()))
    # R1 idempotence-style property: resizing to the same shape is the identity on a binary mask
    same = S.resize_bilinear_bin(packed[0, :4], (H, W))
    np.testing.assert_array_equal(same.numpy_u32(), packed[0, :4].numpy_u32())
    # whole-video job == batch-by-batch session
    meta = [{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"]} for p in prompts]
    pm = torch.from_numpy(np.stack([p["segmentation"] for p in prompts])).cuda()
    job = dedup.VideoDedupJob(meta, T, mode="grid", n_max_tracks=64, batch_size=4)
    job.enqueue(logits, pm)
    r = job.finish()
    dd = dedup.TrackDedup(prompts, T, mode="grid", n_max_tracks=64, batch_size=4)
    resized = S.resize_bilinear_bin(packed)
    while (batch := dd.next_batch()) is not None:
        dd.submit_resized(batch, resized[batch])
    r2 = dd.result()
    assert r["tracked"] == r2["tracked"] and r["filtered"] == r2["filtered"] and r["filtered_by"] == r2["filtered_by"]
    # one reference-exact IoU on the sub-sample: track 0 at prompt 1's frame vs prompt 1 (CUDA bilinear == torch-CUDA)
    f = prompts[1]["frame_idx"]
    m0 = (logits[0, f:f + 1] > 0).float()
    up = (torch.nn.functional.interpolate(m0[None], size=(540, 960), mode="bilinear") > 0.5)[0, 0].float()
    pmr = torch.nn.functional.interpolate(pm[1].float()[None, None], size=(540, 960), mode="nearest")[0, 0]
    assert r["iou_gather"][0, 1] == O.compute_mask_iou(up.cpu(), pmr.cpu())


def test_config5_shape_partial():
    """Config 5 geometry (1080x1920, r = 18) on a slice: 6 tracks x 12 frames."""
    import sola_b200 as S
    from sola_b200 import synth
    logits = synth.smooth_logits(6 * 12, 1080, 1920, seed=5, device="cuda", cell=120).view(6, 12, 1080, 1920)
    packed, counts = S.binarize_pack_stability(logits)
    ref = logits[1, 3].cpu().numpy()
    np.testing.assert_array_equal(packed.words[1, 3].cpu().numpy().view(np.uint32), O.pack_bits(ref > 0))
    inter = S.pairwise_inter_matrix(packed).cpu().numpy()
    flat = (logits > 0).view(6, -1)
    exp = (flat[:, None, :] & flat[None, :, :]).sum(-1).cpu().numpy()
    np.testing.assert_array_equal(inter, exp)
    bc = S.boundary_counts(packed[0, :2], packed[1, :2]).cpu().numpy()
    assert tuple(bc[:, 0]) == BO.boundary_counts((logits[0, 0] > 0).cpu().numpy(), (logits[1, 0] > 0).cpu().numpy())
    assert BO.bound_pix_for(1080, 1920) == 18
