"""CPU: the C-ABI library builds, loads, and exports every symbol include/sola_maskpath.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "sola_maskpath.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sola_[a-z0-9_]+)\s*\(", text)))


def _build_digest():
    from sola_b200 import _build
    return _build.source_digest().encode()


def test_stale_library_is_refused(tmp_path, monkeypatch):
    """A library whose compiled-in source digest differs from csrc/ must not load silently (stale .so hazard)."""
    import pytest
    from sola_b200 import _build, _lib
    _build.build()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_build, "_source_digest", lambda: "0" * 64)
    monkeypatch.setattr(_build, "build", lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no nvcc")))
    monkeypatch.delenv("SOLA_ALLOW_STALE_LIB", raising=False)
    with pytest.raises(_lib.SolaError, match="stale|other sources"):
        _lib.load()
    monkeypatch.setenv("SOLA_ALLOW_STALE_LIB", "1")
    with pytest.warns(RuntimeWarning):
        assert _lib.load() is not None
    monkeypatch.setattr(_lib, "_lib", None)


def test_header_symbols_exported():
    from sola_b200 import _build, _lib
    path = _build.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # the ctypes table binds exactly the declared entry points
    assert sorted(_lib.SIGNATURES) == names


def test_library_identity_and_error_string():
    from sola_b200 import _lib
    lib = _lib.load()
    assert lib.sola_version() >= 100
    assert lib.sola_build_arch() == b"sm_100a"
    assert isinstance(lib.sola_launch_count(), int)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.sola_binarize_pack_f32(None, 1, 4, 4, 0.0, 1.0, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.sola_last_error_string()
    rc = lib.sola_pair_iou_st(None, 4, 16, None, None, None)
    assert rc == -1
    rc = lib.sola_jf_boundary_packed(16, 16, 1, 8, 8, 99, 16, None)
    assert rc == -3 and b"radius" in lib.sola_last_error_string()
    assert lib.sola_jf_boundary_packed(None, None, 1, 8, 8, 3, None, None) == -1 and b"null" in lib.sola_last_error_string()
    assert lib.sola_build_digest() == _build_digest()
    # the multi-GPU / J&F entry points validate the same way: status + message, never an exception or a crash
    assert lib.sola_pair_iou_st_rows(None, 4, 18, 0, 1, None, None) == -1 and b"multiple of 4" in lib.sola_last_error_string()
    assert lib.sola_pair_iou_st_rows(None, 4, 16, 2, 2, None, None) == -1 and b"partition" in lib.sola_last_error_string()
    assert lib.sola_pull_rows(None, 4, 2, 16, None, None) == -1 and b"aligned" in lib.sola_last_error_string()
    assert lib.sola_pull_rows(None, 0, 0, 16, None, None) == 0                       # nothing to do is not an error
    assert lib.sola_jf_f32(None, None, -1, 16, None, None, None, None) == -1
    assert lib.sola_jf_packed(None, None, 4, 16, None, None, None, None) == -1 and b"null" in lib.sola_last_error_string()
    assert lib.sola_pair_iou_st_part(1, 4, 16, 3, 2, 1, None) == -1


def test_no_cpu_fallback():
    """Without a CUDA device every product entry point must raise instead of computing on the host."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sola_b200
    from sola_b200 import seg_utils, evaluator, prompt_generator
    m = np.zeros((4, 4), np.float32)
    for fn in (lambda: seg_utils.compute_mask_iou(m, m), lambda: evaluator.compute_J(m[None], m[None]),
               lambda: prompt_generator.get_stability_score(m), lambda: sola_b200.pack_masks(m)):
        with pytest.raises(Exception):
            fn()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sola_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
