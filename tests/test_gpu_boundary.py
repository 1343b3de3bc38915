"""GPU parity: boundary-F extension vs the numpy oracle of the DAVIS definition (parity unpinned by the reference)."""
import numpy as np
import pytest
import torch

from oracle import boundary_oracle as BO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,H,W", [(4, 48, 85), (3, 64, 64), (2, 480, 854), (2, 100, 33), (1, 720, 1280), (2, 35, 1000)])
def test_boundary_counts_vs_oracle(T, H, W):
    import sola_b200 as S
    from sola_b200 import evaluator, synth
    pred, gt = synth.jf_pair(T, H, W, seed=T * H + W, device="cpu", flip=0.003, empty_frames=1 if T > 1 else 0)
    pred, gt = pred.numpy(), gt.numpy()
    if T > 2:
        pred[2] = 0                                  # n_fg == 0, n_gt > 0 rule
    pred[0, -3:, -5:] = 1                            # touch the last row / column / corner rules
    gt[0, :2, :2] = 1
    c = S.boundary_counts(S.pack_masks(pred), S.pack_masks(gt)).cpu().numpy()
    for t in range(T):
        assert tuple(c[:, t]) == BO.boundary_counts(pred[t], gt[t]), t
    f = evaluator.compute_F_boundary(pred, gt)
    assert abs(f - BO.boundary_f_masklet(pred, gt)) < 1e-12


def test_boundary_radius_sweep_small():
    import sola_b200 as S
    rng = np.random.default_rng(0)
    seg = np.zeros((1, 70, 90), np.uint8)
    seg[0, 20:50, 25:70] = 1
    seg[0] ^= (rng.random((70, 90)) > 0.98).astype(np.uint8)
    gt = np.roll(seg, 3, axis=2)
    for bound in (1, 2, 5, 13, 31):
        c = S.boundary_counts(S.pack_masks(seg), S.pack_masks(gt), bound_th=bound).cpu().numpy()
        assert tuple(c[:, 0]) == BO.boundary_counts(seg[0], gt[0], bound_th=bound), bound
    from sola_b200 import _lib
    with pytest.raises(_lib.SolaError):
        S.boundary_counts(S.pack_masks(seg), S.pack_masks(gt), bound_th=40)
