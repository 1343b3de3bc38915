"""GPU: the C ABI used directly — exactly the ctypes stub of INTEGRATION.md §2, no sola_b200 wrapper in between — plus
hypothesis-driven shape / threshold sweeps of K1 and K3 against first principles."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

LIB = os.path.join(ROOT, "sola_b200", "lib", "libsola_maskpath.so")


def _lib():
    from sola_b200 import _build
    _build.build()
    lib = ctypes.CDLL(LIB)
    lib.sola_last_error_string.restype = ctypes.c_char_p
    vp, ll, i, d = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_double
    lib.sola_binarize_pack_f32.argtypes = [vp, ll, i, i, d, d, vp, vp, vp, vp, vp]
    lib.sola_frame_counts_f32.argtypes = [vp, vp, ll, ll, vp, vp, vp, vp]
    return lib


def test_integration_stub_binarize_and_counts():
    lib = _lib()
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn((6, 96, 160), generator=g) * 2).cuda()
    n, H, W = logits.shape
    packed = torch.empty((n, H, (W + 31) // 32), dtype=torch.int32, device="cuda")
    counts = torch.empty((3, n), dtype=torch.int32, device="cuda")
    rc = lib.sola_binarize_pack_f32(logits.data_ptr(), n, H, W, 0.0, 1.0, packed.data_ptr(), counts[0].data_ptr(), counts[1].data_ptr(),
                                    counts[2].data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.sola_last_error_string()
    ref = logits.cpu().numpy()
    bits = np.unpackbits(packed.cpu().numpy().view(np.uint8).reshape(n, H, -1), axis=-1, bitorder="little")[..., :W]
    np.testing.assert_array_equal(bits, (ref > 0).astype(np.uint8))
    c = counts.cpu().numpy()
    np.testing.assert_array_equal(c[0], (ref > 1).sum((1, 2)))
    np.testing.assert_array_equal(c[2], (ref > -1).sum((1, 2)))
    a, b = (logits > 0).float(), (logits > 0.5).float()
    out = torch.empty((3, n), dtype=torch.int32, device="cuda")
    rc = lib.sola_frame_counts_f32(a.data_ptr(), b.data_ptr(), n, a[0].numel(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    np.testing.assert_array_equal(out.cpu().numpy(), torch.stack([(a * b).sum((1, 2)), a.sum((1, 2)), b.sum((1, 2))]).int().cpu().numpy())
    # error path: status code + message, nothing thrown
    assert lib.sola_binarize_pack_f32(None, 1, 4, 4, 0.0, 1.0, None, None, None, None, None) == -1
    assert b"null" in lib.sola_last_error_string()


def test_side_stream_and_no_implicit_sync():
    """Work is enqueued on the given stream only: results appear after synchronising THAT stream."""
    import sola_b200 as S
    s = torch.cuda.Stream()
    x = torch.randn(4, 720, 1280, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        packed, counts = S.binarize_pack_stability(x)
    s.synchronize()
    np.testing.assert_array_equal(counts.cpu().numpy()[1], (x > 0).sum((1, 2)).cpu().numpy())


def test_hypothesis_k1_k3_shapes():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import sola_b200 as S

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(1, 3), H=st.integers(1, 70), W=st.integers(1, 200), thr=st.floats(-2, 2), off=st.floats(0, 2),
           seed=st.integers(0, 10_000), wmul=st.sampled_from([0, 32]))
    def run(n, H, W, thr, off, seed, wmul):
        if wmul:
            W = max(32, (W // 32) * 32)                 # exercise the flat path as well as the row path
        g = torch.Generator().manual_seed(seed)
        x = torch.randn((n, H, W), generator=g) * 2
        x.view(-1)[::7] = float(np.float32(thr))        # values exactly at the thresholds
        x.view(-1)[3::11] = float(np.float32(thr + off))
        packed, counts = S.binarize_pack_stability(x.cuda(), thr, off)
        ref = x.numpy()
        t_mid, t_hi, t_lo = np.float32(thr), np.float32(thr + off), np.float32(thr - off)
        bits = np.unpackbits(packed.numpy_u32().view(np.uint8).reshape(n, H, -1), axis=-1, bitorder="little")
        np.testing.assert_array_equal(bits[..., :W], (ref > t_mid).astype(np.uint8))
        assert not bits[..., W:].any()                  # pad bits stay zero
        c = counts.cpu().numpy()
        np.testing.assert_array_equal(c[0], (ref > t_hi).sum((1, 2)))
        np.testing.assert_array_equal(c[2], (ref > t_lo).sum((1, 2)))
        a, b = (ref > t_mid), (ref > t_hi)
        k = S.frame_counts(a.astype(np.float32), b.astype(np.uint8).astype(np.float32)).cpu().numpy()
        np.testing.assert_array_equal(k[0], (a & b).sum((1, 2)))
        # |A ∩ B| + |A ∪ B| == |A| + |B|
        u = S.frame_counts((a | b).astype(np.uint8), (a | b).astype(np.uint8)).cpu().numpy()[1]
        np.testing.assert_array_equal(k[0] + u, k[1] + k[2])

    run()


def test_pure_c_client(tmp_path):
    """A plain C program (no Python, no torch) links the shared library through include/sola_maskpath.h and checks K1, K3 and the
    fused entry point against C loops."""
    import shutil
    import subprocess
    from sola_b200 import _build
    _build.build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.join(ROOT, "sola_b200", "lib")
    subprocess.run([nvcc, "-x", "c", os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-I", os.path.join(ROOT, "include"), "-L", libdir,
                    "-lsola_maskpath", "-Xlinker", f"-rpath={libdir}", "-o", exe], check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "c_abi_smoke ok" in res.stdout
