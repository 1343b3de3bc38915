"""COCO-RLE codec: the oracle's self-consistency pins (CPU) and the GPU codec against the oracle."""
import numpy as np
import pytest
import torch

from oracle import rle_oracle as RO


def _cases():
    rng = np.random.default_rng(0)
    out = [np.zeros((4, 4), np.uint8), np.ones((4, 4), np.uint8), np.eye(5, dtype=np.uint8), (rng.random((37, 70)) > 0.5).astype(np.uint8)]
    blob = np.zeros((60, 94), np.uint8)
    blob[10:40, 20:70] = 1
    blob[25:30, 30:50] = 0
    out.append(blob)
    tall = np.zeros((200, 33), np.uint8)
    tall[5:190, 3:30] = 1
    out.append(tall)
    first = np.zeros((3, 3), np.uint8)
    first[0, 0] = 1                                            # first pixel set -> leading zero-length run
    out.append(first)
    return out


def test_oracle_hand_derived_strings():
    assert RO.encode(np.zeros((2, 2), np.uint8))["counts"] == "4"          # one zeros-run of 4
    assert RO.encode(np.ones((2, 2), np.uint8))["counts"] == "04"          # empty zeros-run, then 4 ones
    assert RO.mask_to_counts(np.array([[0, 1], [1, 1]], np.uint8)) == [1, 3]
    assert RO.encode(np.zeros((4, 4), np.uint8))["counts"] == "`0"          # 16 = 0b10000 needs the continuation char
    # delta coding kicks in from the 4th count on: counts [2,1,2,1,...] -> later entries stored as differences
    m = np.tile(np.array([0, 0, 1], np.uint8), 6).reshape(6, 3, order="F")
    c = RO.mask_to_counts(m)
    assert c == [2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1]
    assert RO.string_to_counts(RO.counts_to_string(c)) == c
    big = [0, 100000, 5, 70000, 123456, 1]
    assert RO.string_to_counts(RO.counts_to_string(big)) == big             # negative deltas / sign extension


def test_oracle_roundtrip():
    for m in _cases():
        r = RO.encode(m)
        assert r["size"] == list(m.shape) and isinstance(r["counts"], str)
        np.testing.assert_array_equal(RO.decode(r), m)
    masklet = np.stack([(np.random.default_rng(k).random((20, 30)) > 0.6).astype(np.uint8) for k in range(4)])
    rl = RO.encode_masklet(masklet)
    rl[2] = None                                               # missing frame -> zero mask (dataloader.py:364-368)
    dec = RO.decode_masklet(rl)
    np.testing.assert_array_equal(dec[[0, 1, 3]], masklet[[0, 1, 3]])
    assert dec[2].sum() == 0


def test_host_string_codec_matches_oracle():
    from sola_b200 import rle
    for m in _cases():
        c = RO.mask_to_counts(m)
        s = RO.counts_to_string(c)
        assert rle.counts_to_string(c) == s
        assert rle.string_to_counts(s).tolist() == c


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 37, 70), (3, 480, 854), (2, 64, 64), (5, 33, 100), (2, 720, 1280), (1, 1, 1), (2, 200, 33)])
def test_gpu_codec_vs_oracle(shape):
    import sola_b200 as S
    from sola_b200 import rle, synth
    T, H, W = shape
    m = (synth.smooth_logits(T, H, W, seed=H + W, device="cpu", cell=max(4, H // 5)) > 0).to(torch.uint8).numpy()
    m[0, 0, 0] = 1
    if T > 1:
        m[1] = 0
    exp = RO.encode_masklet(m)
    got = rle.encode_rle_masklet_torch(torch.from_numpy(m).float().cuda())            # the reference hands fp32 masks (seg_utils.py:100)
    assert got == exp
    assert rle.encode_rle_masklet_torch(S.pack_masks(m)) == exp
    dec = rle.decode_rle_masklet(exp)
    assert dec.dtype == np.uint8
    np.testing.assert_array_equal(dec, m)
    packed = rle.decode_rle_masklet_packed(exp)
    np.testing.assert_array_equal(packed.numpy_u32(), S.pack_masks(m).numpy_u32())
    with_gap = list(exp)
    if T > 2:
        with_gap[2] = None
        np.testing.assert_array_equal(rle.decode_rle_masklet(with_gap), RO.decode_masklet(with_gap))
    np.testing.assert_array_equal(rle.decode_rle_mask(exp[0]), m[0])


@pytest.mark.gpu
def test_gpu_bit_transpose_and_encode_overflow_fallback():
    import sola_b200 as S
    from sola_b200 import rle, _lib, packed as P
    rng = np.random.default_rng(3)
    m = (rng.random((3, 45, 70)) > 0.5).astype(np.uint8)
    p = S.pack_masks(m)
    out = torch.empty((3, 70, 2), dtype=torch.int32, device="cuda")
    _lib.call("sola_bit_transpose", p.words.data_ptr(), 3, 45, 70, out.data_ptr(), None)
    from oracle.maskpath_oracle import pack_bits
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), pack_bits(m.transpose(0, 2, 1)))
    # noisy mask with more transitions than the cap -> host fallback gives the same string
    assert rle.encode_rle_masklet_packed(p, cap=64) == RO.encode_masklet(m)


@pytest.mark.gpu
def test_gpu_get_sam2_masklet_packed_and_png_and_sidecar(tmp_path):
    """RLE lists -> selected OR-merge on the device == the reference's numpy path (oracle), PNG planes, packed sidecar round trip."""
    import sola_b200 as S
    from sola_b200 import dataloader_ops as D, evaluator
    from oracle import maskpath_oracle as O
    rng = np.random.default_rng(5)
    tracks = [(rng.random((4, 45, 70)) > 0.8).astype(np.uint8) for _ in range(5)]
    rles = [RO.encode_masklet(t) for t in tracks]
    for preds in ([0, 1, 0, 1, 1], [0, 0, 0, 0, 0], [1, 0, 0, 0, 0]):
        exp = np.asarray(O.merge_selected_tracks([RO.decode_masklet(r) for r in rles], preds)) != 0
        got = D.get_sam2_masklet_packed(rles, preds)
        np.testing.assert_array_equal(S.unpack_masks(got, torch.uint8).cpu().numpy(), exp.astype(np.uint8))
    assert D.get_sam2_masklet_packed([], []) is None
    merged = D.get_sam2_masklet_packed(rles, [0, 1, 0, 1, 1])
    png = D.png_planes(merged)
    ref = (tracks[1] | tracks[3] | tracks[4])
    np.testing.assert_array_equal(png, (ref * 255).astype(np.uint8))              # inference.py:90
    D.save_packed_masklet(str(tmp_path / "m.npz"), merged)
    back = D.load_packed_masklet(str(tmp_path / "m.npz"))
    np.testing.assert_array_equal(back.numpy_u32(), merged.numpy_u32())
    # J&F straight from the packed planes
    gt = S.pack_masks(tracks[0])
    c = S.frame_counts_packed(merged.reshape_lead(1, 4), gt.reshape_lead(1, 4))[0].cpu().numpy()[0, 0]
    np.testing.assert_array_equal(c, (ref & tracks[0]).sum((1, 2)))
