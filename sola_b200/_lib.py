"""ctypes binding of libsola_maskpath.so (the C ABI declared in include/sola_maskpath.h).

There is no CPU fallback: if the library is missing or was built for another architecture, loading raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

_P, _LL, _I, _D = C.c_void_p, C.c_longlong, C.c_int, C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "sola_version": [],
    "sola_last_error_string": [],
    "sola_build_arch": [],
    "sola_build_digest": [],
    "sola_launch_count": [],
    "sola_binarize_pack_f32": [_P, _LL, _I, _I, _D, _D, _P, _P, _P, _P, _P],
    "sola_binarize_pack_bf16": [_P, _LL, _I, _I, _D, _D, _P, _P, _P, _P, _P],
    "sola_threshold_pack_f32": [_P, _LL, _I, _I, _D, _P, _P, _P],
    "sola_pack_mask_f32": [_P, _LL, _I, _I, _P, _P, _P],
    "sola_pack_mask_u8": [_P, _LL, _I, _I, _P, _P, _P],
    "sola_unpack_f32": [_P, _LL, _I, _I, _P, _P],
    "sola_unpack_u8": [_P, _LL, _I, _I, _P, _P],
    "sola_unpack_u8_value": [_P, _LL, _I, _I, _I, _P, _P],
    "sola_frame_counts_f32": [_P, _P, _LL, _LL, _P, _P, _P, _P],
    "sola_frame_counts_u8": [_P, _P, _LL, _LL, _P, _P, _P, _P],
    "sola_frame_counts_packed": [_P, _P, _I, _I, _I, _LL, _P, _P, _P, _P],
    "sola_frame_counts_packed_ragged": [_P, _P, _P, _I, _P, _P, _P, _P],
    "sola_jf_f32": [_P, _P, _LL, _LL, _P, _P, _P, _P],
    "sola_jf_u8": [_P, _P, _LL, _LL, _P, _P, _P, _P],
    "sola_jf_packed": [_P, _P, _LL, _LL, _P, _P, _P, _P],
    "sola_or_merge": [_P, _P, _I, _LL, _P, _P],
    "sola_pair_iou_st": [_P, _I, _LL, _P, _P, _P],
    "sola_pair_iou_st_part": [_P, _I, _LL, _I, _I, _P, _P],
    "sola_pair_iou_st_rows": [_P, _I, _LL, _I, _I, _P, _P],
    "sola_pair_iou_st_peer": [_P, _I, _I, _LL, _I, _I, _P, _P],
    "sola_pair_iou_st_accumulate": [_P, _I, _LL, _P, _P],
    "sola_pull_rows": [_P, _I, _LL, _LL, _P, _P],
    "sola_pair_iou_gather": [_P, _P, _P, _I, _I, _I, _LL, _P, _P, _P, _P],
    "sola_resize_bilinear_bin_packed": [_P, _LL, _I, _I, _I, _I, _P, _P, _P],
    "sola_resize_bilinear_bin_f32": [_P, _LL, _I, _I, _I, _I, _P, _P, _P, _P],
    "sola_resize_nearest_u8": [_P, _LL, _I, _I, _I, _I, _P, _P, _P],
    "sola_resize_nearest_packed": [_P, _LL, _I, _I, _I, _I, _P, _P, _P],
    "sola_binarize_pack_resize_f32": [_P, _LL, _I, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P, _P],
    "sola_binarize_pack_resize_bf16": [_P, _LL, _I, _I, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P, _P],
    "sola_bit_transpose": [_P, _LL, _I, _I, _P, _P],
    "sola_rle_decode_runs": [_P, _P, _P, _LL, _LL, _I, _I, _P, _P, _P],
    "sola_rle_strings_to_runs": [_P, _P, _P, _LL, _LL, _P, _P, _P, _LL, _P],
    "sola_rle_encode_transitions": [_P, _LL, _I, _I, _P, _I, _P, _P, _P],
    "sola_jf_sweep_plan": [_P, _I, _P],
    "sola_jf_sweep": [_P, _I, _P, _P, _P],
    "sola_jf_boundary_packed": [_P, _P, _LL, _I, _I, _I, _P, _P],
}
_RESTYPES = {
    "sola_last_error_string": C.c_char_p,
    "sola_build_arch": C.c_char_p,
    "sola_build_digest": C.c_char_p,
    "sola_launch_count": C.c_ulonglong,
}


class SolaError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return os.environ.get("SOLA_MASKPATH_LIB", _build.LIB_PATH)


def load(build_if_missing: bool = True):
    """Load (building in-tree first if the .so is absent or stale and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if path == _build.LIB_PATH and build_if_missing and not _build.is_current():
            try:
                _build.build()
            except Exception as e:
                if not os.path.isfile(path):
                    raise SolaError(f"libsola_maskpath.so is missing and could not be built: {e}") from e
                # a library older than csrc/ may have another ABI or other numerics under the same symbol names: refuse it
                if os.environ.get("SOLA_ALLOW_STALE_LIB") != "1":
                    raise SolaError(f"csrc/ changed since {path} was built and the rebuild failed ({e}); refusing the stale library "
                                    "(set SOLA_ALLOW_STALE_LIB=1 to load it anyway)") from e
                import warnings
                warnings.warn(f"sola_b200: loading a STALE {path} (SOLA_ALLOW_STALE_LIB=1; rebuild failed: {e})", RuntimeWarning)
        if not os.path.isfile(path):
            raise SolaError(f"{path} not found — run `python -m sola_b200._build` (no CPU fallback exists)")
        lib = C.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        arch = lib.sola_build_arch().decode()
        if arch != "sm_100a":
            raise SolaError(f"library built for {arch}, expected sm_100a")
        if path == _build.LIB_PATH and os.environ.get("SOLA_ALLOW_STALE_LIB") != "1":
            built_from, have = lib.sola_build_digest().decode(), _build.source_digest()
            if built_from != have:
                raise SolaError(f"{path} was built from other sources (digest {built_from[:12]}, csrc/ is {have[:12]}); rebuild with "
                                "`python -m sola_b200._build --force` or set SOLA_ALLOW_STALE_LIB=1")
        _lib = lib
        return lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise SolaError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise SolaError(f"{name} failed (rc={rc}): {lib.sola_last_error_string().decode(errors='replace')}")


def launch_count() -> int:
    return int(load().sola_launch_count())
