"""GPU parity: K1 (binarise + pack + stability counts), mask packers / unpackers — vs the oracle and the golden vectors."""
import warnings

import numpy as np
import pytest
import torch

from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


def _check_k1(logits_cpu: torch.Tensor, thr=0.0, off=1.0):
    import sola_b200 as S
    x = logits_cpu.cuda()
    packed, counts = S.binarize_pack_stability(x, thr, off)
    ref = logits_cpu.float().numpy()
    t_mid, t_hi, t_lo = np.float32(thr), np.float32(thr + off), np.float32(thr - off)
    np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(ref > t_mid))            # bit-exact planes
    c = counts.cpu().numpy()
    lead_axes = (-2, -1)
    np.testing.assert_array_equal(c[0], (ref > t_hi).sum(lead_axes))
    np.testing.assert_array_equal(c[1], (ref > t_mid).sum(lead_axes))
    np.testing.assert_array_equal(c[2], (ref > t_lo).sum(lead_axes))
    return packed, counts


@pytest.mark.parametrize("tag", ["a", "b"])
def test_golden_stability_and_planes(golden, tag):
    from sola_b200 import prompt_generator as PG
    logits = golden[f"stab_{tag}_logits"]
    packed, _ = _check_k1(torch.from_numpy(logits))
    np.testing.assert_array_equal(packed.numpy_u32(), golden[f"stab_{tag}_binarized_packed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array([PG.get_stability_score(l) for l in logits])
        got2 = np.array([PG.PromptGenerator().get_stability_score(l, 0.5, 0.25) for l in logits])
        batched = PG.get_stability_score(logits)
    np.testing.assert_array_equal(got, golden[f"stab_{tag}_score"])                          # incl. the nan plane
    np.testing.assert_array_equal(got2, golden[f"stab_{tag}_score_t05_o025"])
    np.testing.assert_array_equal(batched, golden[f"stab_{tag}_score"])
    assert isinstance(got[0], np.float64)


@pytest.mark.parametrize("shape", [(3, 64, 96), (2, 5, 720, 1280), (1, 540, 960), (4, 33, 32), (2, 480, 854),
                                   (3, 37, 70), (1, 1, 1), (2, 7, 1000), (1, 1080, 1920), (5, 3, 17, 31), (4, 37, 70), (8, 5, 33), (2, 481, 854), (4, 64, 2000)])
def test_random_shapes_fp32(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 2.0
    x.view(-1)[::97] = 0.0
    x.view(-1)[5::101] = 1.0
    x.view(-1)[7::103] = -1.0
    x.view(-1)[11::107] = float("nan")
    _check_k1(x)
    _check_k1(x, 0.3, 0.7)


def test_unaligned_base_pointer_takes_row_path():
    g = torch.Generator().manual_seed(3)
    big = torch.randn(2 * 64 * 96 + 1, generator=g).cuda()
    x = big[1:].view(2, 64, 96)                      # 4-byte aligned only
    import sola_b200 as S
    packed, counts = S.binarize_pack_stability(x)
    ref = x.cpu().numpy()
    np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(ref > 0))
    np.testing.assert_array_equal(counts.cpu().numpy()[2], (ref > -1).sum((-2, -1)))


@pytest.mark.parametrize("shape", [(3, 64, 96), (2, 720, 1280), (2, 37, 70), (8, 37, 70), (4, 480, 854)])
def test_bf16_logits(shape):
    import sola_b200 as S
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(shape, generator=g) * 2).to(torch.bfloat16)
    x.view(-1)[::53] = 1.0
    x.view(-1)[1::59] = -1.0
    x.view(-1)[2::61] = 0.0
    packed, counts = S.binarize_pack_stability(x.cuda())
    ref = x.float().numpy()
    np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(ref > 0))
    c = counts.cpu().numpy()
    np.testing.assert_array_equal(c[0], (ref > 1).sum((-2, -1)))
    np.testing.assert_array_equal(c[2], (ref > -1).sum((-2, -1)))
    # thresholds that are NOT bf16-representable (the packed-compare count path floors them to bf16; must stay exact), incl. negatives
    for thr, off in ((0.3, 0.7), (-0.37, 0.11), (0.0, 1e-3), (2.0, 3.0)):
        packed, counts = S.binarize_pack_stability(x.cuda(), thr, off)
        c = counts.cpu().numpy()
        np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(ref > np.float32(thr)))
        np.testing.assert_array_equal(c[0], (ref > np.float32(thr + off)).sum((-2, -1)))
        np.testing.assert_array_equal(c[1], (ref > np.float32(thr)).sum((-2, -1)))
        np.testing.assert_array_equal(c[2], (ref > np.float32(thr - off)).sum((-2, -1)))


@pytest.mark.parametrize("shape", [(4, 64, 96), (3, 480, 854), (2, 33, 47), (1, 720, 1280)])
@pytest.mark.parametrize("dtype", ["f32", "u8", "bool"])
def test_pack_unpack_masks(shape, dtype):
    import sola_b200 as S
    rng = np.random.default_rng(sum(shape))
    m = rng.random(shape) > 0.6
    m[0, :2] = False
    src = {"f32": m.astype(np.float32), "u8": m.astype(np.uint8) * 255, "bool": m}[dtype]   # any nonzero byte is foreground
    packed, area = S.pack_masks(src, want_area=True)
    np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(m))
    np.testing.assert_array_equal(area.cpu().numpy(), m.sum((-2, -1)))
    for dt, npdt in ((torch.float32, np.float32), (torch.uint8, np.uint8)):
        back = S.unpack_masks(packed, dt).cpu().numpy()
        assert back.dtype == npdt
        np.testing.assert_array_equal(back, m.astype(npdt))


def test_counts_only_and_planes_only():
    import sola_b200 as S
    x = torch.randn(3, 64, 64).cuda()
    p, c = S.binarize_pack_stability(x, want_counts=False)
    assert c is None and p is not None
    p2, c2 = S.binarize_pack_stability(x, want_packed=False)
    assert p2 is None and c2.shape == (3, 3)
    np.testing.assert_array_equal(c2.cpu().numpy()[1], (x.cpu().numpy() > 0).sum((-2, -1)))


def test_empty_batch_and_errors():
    import sola_b200 as S
    from sola_b200 import _lib
    p, c = S.binarize_pack_stability(torch.zeros(0, 8, 8).cuda())
    assert p.words.shape == (0, 8, 1) and c.shape == (3, 0)
    with pytest.raises(_lib.SolaError):
        _lib.call("sola_binarize_pack_f32", None, 1, 8, 8, 0.0, 1.0, None, None, None, None, None)


def test_negative_zero_threshold():
    """thr - off can round to -0.0 (found by the hypothesis sweep): `x > -0.0` must behave like `x > 0.0`."""
    import sola_b200 as S
    x = torch.tensor([[0.0, -0.0, 1e-45, -1e-45, 1.0, -1.0, float("nan"), 0.0]]).repeat(4, 4)[None]
    for thr, off in ((0.0, 1.2e-211), (-0.0, 0.0), (0.0, 0.0)):
        packed, counts = S.binarize_pack_stability(x.cuda(), thr, off)
        ref = x.numpy()
        for k, t in enumerate((np.float32(thr + off), np.float32(thr), np.float32(thr - off))):
            np.testing.assert_array_equal(counts.cpu().numpy()[k], (ref > t).sum((1, 2)))
        np.testing.assert_array_equal(packed.numpy_u32(), O.pack_bits(ref > np.float32(thr)))
