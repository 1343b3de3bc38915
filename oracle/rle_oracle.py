"""TEST INFRASTRUCTURE ONLY — COCO run-length codec, restated in numpy / pure Python.

The reference calls `pycocotools.mask.encode / decode` (pycocotools==2.0.8, requirements.txt:30; call sites
track_generation/seg_utils.py:16,67,87,103, track_generation/utils.py:21,36,55, dataloader.py:360).  pycocotools is a
third-party dependency that is NOT vendored under /root/reference and is not installed in this image, so this file restates
its published algorithm (cocoapi `common/maskApi.c`: rleEncode, rleDecode, rleToString, rleFrString) and the parity of the
GPU codec is **unpinned by the reference**: there is no golden RLE string in the reference tree to check against.
Self-consistency pins used instead: decode(encode(m)) == m on random / blob / empty / full masks, hand-derived strings for
tiny masks (tests/test_rle.py), and agreement between this oracle and the GPU codec on every case.

Format: the mask is scanned in column-major (Fortran) order; `counts` are run lengths, alternating zeros-run, ones-run, starting
with a zeros-run (possibly 0).  The compressed string stores count[i] for i < 3 and count[i] - count[i-2] afterwards, as a
little-endian base-32 varint with a continuation bit (0x20) and sign extension from bit 0x10, offset by ASCII 48.
"""
from __future__ import annotations

from typing import List

import numpy as np


def mask_to_counts(mask: np.ndarray) -> List[int]:
    """rleEncode: (H, W) {0,1} -> uncompressed counts."""
    flat = np.asarray(mask).astype(bool).ravel(order="F")
    n = flat.size
    if n == 0:
        return []
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [n]])
    counts = np.diff(bounds).tolist()
    if flat[0]:
        counts = [0] + counts
    return counts


def counts_to_mask(counts, H: int, W: int) -> np.ndarray:
    """rleDecode: uncompressed counts -> (H, W) uint8."""
    vals = np.zeros(len(counts), dtype=np.uint8)
    vals[1::2] = 1
    flat = np.repeat(vals, np.asarray(counts, dtype=np.int64))
    assert flat.size == H * W, f"counts sum {flat.size} != {H}*{W}"
    return flat.reshape((H, W), order="F")


def counts_to_string(counts) -> str:
    """rleToString."""
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5                                   # arithmetic shift (Python ints), as in C on a signed long
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def string_to_counts(s) -> List[int]:
    """rleFrString."""
    if isinstance(s, bytes):
        s = s.decode("ascii")
    counts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def encode(mask: np.ndarray) -> dict:
    """pycocotools.mask.encode for one (H, W) mask; `counts` as str (the reference .decode('utf-8')s it, seg_utils.py:88)."""
    H, W = mask.shape
    return {"size": [int(H), int(W)], "counts": counts_to_string(mask_to_counts(mask))}


def decode(rle: dict) -> np.ndarray:
    """pycocotools.mask.decode for one RLE dict -> (H, W) uint8."""
    H, W = rle["size"]
    return counts_to_mask(string_to_counts(rle["counts"]), H, W)


def encode_masklet(masks: np.ndarray) -> List[dict]:
    """seg_utils.encode_rle_masklet_torch (seg_utils.py:93-106) without the torch -> numpy hop."""
    return [encode(m) for m in np.asarray(masks)]


def decode_masklet(rle_masklet) -> np.ndarray:
    """AlignDataset.rle_masklet_decode (dataloader.py:353-369): non-dict entries (missing frames) become zero masks of the
    last seen shape."""
    out, h, w = [], 0, 0
    for r in rle_masklet:
        if isinstance(r, dict):
            m = decode(r)
            h, w = m.shape
            out.append(m)
        else:
            out.append(None)
    out = [np.zeros((h, w), np.uint8) if m is None else m for m in out]
    return np.stack(out, axis=0)
