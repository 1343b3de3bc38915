"""GPU parity: R1 (bilinear + > 0.5) must be bit-identical to torch-CUDA's F.interpolate, which is what the reference
executes on the GPU (seg_utils.py:158); R2 nearest likewise.  Golden vectors (torch-CPU) are compared with a tie budget."""
import numpy as np
import pytest
import torch

from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


def _torch_cuda_reshape(m_f32_cuda, target=None):
    T, H, W = m_f32_cuda.shape
    nh, nw = ((540, 960) if H < W else (960, 540)) if target is None else target
    up = torch.nn.functional.interpolate(m_f32_cuda[None], size=(nh, nw), mode="bilinear")
    return (up > 0.5)[0].float(), up[0]


def _blobs(n, H, W, seed):
    from sola_b200 import synth
    return (synth.smooth_logits(n, H, W, seed, device="cpu", cell=12) > 0).float()


@pytest.mark.parametrize("H,W,target", [(720, 1280, None), (1080, 1920, None), (480, 854, None), (1280, 720, None),
                                        (64, 64, None), (72, 128, (45, 80)), (37, 70, (50, 33)), (540, 960, None),
                                        (360, 640, None), (100, 100, (100, 100))])
def test_bilinear_bin_bit_identical_to_torch_cuda(H, W, target):
    import sola_b200 as S
    from sola_b200 import seg_utils
    m = _blobs(3, H, W, H + W).cuda()
    m[0, : H // 3] = 0
    m[1, H // 2:] = 1
    exp, raw = _torch_cuda_reshape(m, target)
    # (a) drop-in fp32 signature
    got = seg_utils.reshape_masklet(m, target)
    assert got.dtype == torch.float32 and got.shape == exp.shape and got.is_cuda
    n_ties = int(((raw - 0.5).abs() < 1e-6).sum())
    assert torch.equal(got, exp), f"{int((got != exp).sum())} pixels differ ({n_ties} within 1e-6 of 0.5)"
    # (b) packed -> packed
    packed = S.resize_bilinear_bin(S.pack_masks(m), target)
    np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(exp.cpu().numpy()))
    # (c) area output
    _, area = S.resize_bilinear_bin(S.pack_masks(m), target, want_area=True)
    np.testing.assert_array_equal(area.cpu().numpy(), exp.sum((-2, -1)).cpu().numpy().astype(np.int64))


def test_bilinear_general_float_input():
    """reshape_masklet on non-binary floats (not a reference use, but the signature allows it)."""
    from sola_b200 import seg_utils
    x = torch.rand(2, 90, 160, generator=torch.Generator().manual_seed(1)).cuda()
    exp, _ = _torch_cuda_reshape(x)
    assert torch.equal(seg_utils.reshape_masklet(x), exp)


def test_golden_cpu_vectors_with_tie_budget(golden):
    """torch-CPU golden outputs: identical except (possibly) pixels where the two ATen kernels round a tie differently."""
    import sola_b200 as S
    for key_in, W, key_out, target in (("rs_land_in", 128, "rs_land_out", None), ("rs_port_in", 72, "rs_port_out", None),
                                       ("rs_sq_in", 64, "rs_sq_out", None), ("rs_land_in", 128, "rs_land_out_45x80", (45, 80))):
        m = golden.masks(key_in, W)
        got = S.resize_bilinear_bin(S.pack_masks(m), target).numpy_u32()
        exp = golden[key_out]
        diff = int(np.unpackbits((got ^ exp).view(np.uint8)).sum())
        assert diff <= 1e-3 * exp.size * 32, (key_out, diff)


@pytest.mark.parametrize("H,W,oh,ow", [(72, 128, 540, 960), (720, 1280, 540, 960), (1080, 1920, 540, 960), (480, 854, 540, 960),
                                        (1280, 720, 960, 540), (33, 47, 21, 90)])
def test_nearest_matches_torch(H, W, oh, ow, golden):
    import sola_b200 as S
    rng = np.random.default_rng(H + ow)
    m = (rng.random((2, H, W)) > 0.5).astype(np.uint8)
    exp = torch.nn.functional.interpolate(torch.from_numpy(m).float().cuda()[None], size=(oh, ow), mode="nearest")[0]
    got = S.resize_nearest(m, oh, ow)
    np.testing.assert_array_equal(got.numpy_u32(), O.pack_bits(exp.cpu().numpy()))
    got2 = S.resize_nearest(S.pack_masks(m), oh, ow)
    np.testing.assert_array_equal(got2.numpy_u32(), got.numpy_u32())
    exp_cpu = torch.nn.functional.interpolate(torch.from_numpy(m).float()[None], size=(oh, ow), mode="nearest")[0]
    np.testing.assert_array_equal(got.numpy_u32(), O.pack_bits(exp_cpu.numpy()))
    if (H, W, oh, ow) == (72, 128, 540, 960):
        land0 = golden.masks("rs_land_in", 128)[0]
        np.testing.assert_array_equal(S.resize_nearest(land0, 540, 960).numpy_u32(), golden["rs_nearest_out"])
