"""COCO-RLE <-> bit-packed planes: drop-ins for the pycocotools hops around the hot path.

  decode_rle_masklet / rle_masklet_decode   seg_utils.py:70-75, dataloader.py:353-369   (RLE list -> masks)
  encode_rle_masklet_torch                  seg_utils.py:93-106                        (masks -> RLE list)

The varint string <-> counts step is sequential and tiny (a few hundred counts per mask) and stays on the host; the pixel work
(run filling, row/column-major transposition, transition extraction) runs on the device on packed bits, so 1/32 of the bytes the
reference moves cross PCIe.  The string format follows cocoapi's rleToString / rleFrString (pycocotools 2.0.8 is not installed
here: parity of the codec is unpinned by the reference, see oracle/rle_oracle.py)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import packed as P


# ---- host: varint string <-> counts ----------------------------------------------------------------------------------

def string_to_counts(s) -> np.ndarray:
    if isinstance(s, str):
        s = s.encode("ascii")
    b = np.frombuffer(s, dtype=np.uint8).astype(np.int64) - 48
    counts: List[int] = []
    x, k = 0, 0
    for c in b.tolist():
        x |= (c & 0x1F) << (5 * k)
        k += 1
        if not (c & 0x20):
            if c & 0x10:
                x |= -1 << (5 * k)
            if len(counts) > 2:
                x += counts[-2]
            counts.append(x)
            x, k = 0, 0
    return np.asarray(counts, dtype=np.int64)


def counts_to_string(counts) -> str:
    counts = [int(c) for c in counts]
    out = bytearray()
    for i, c in enumerate(counts):
        x = c - counts[i - 2] if i > 2 else c
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return out.decode("ascii")


# ---- decode ----------------------------------------------------------------------------------------------------------

def decode_rle_masklet_packed(rle_masklet: Sequence, device=None) -> Optional[P.PackedMasks]:
    """List of COCO RLE dicts (non-dict entries = missing frames -> empty masks, dataloader.py:364-368) -> packed (T, H, Wp)."""
    sizes = [tuple(r["size"]) for r in rle_masklet if isinstance(r, dict)]
    if not sizes:
        return None
    H, W = sizes[-1]
    assert all(s == (H, W) for s in sizes), "all frames of a masklet share one size"
    dev = P._dev(device)
    plane_ids, starts, ends = [], [], []
    for t, r in enumerate(rle_masklet):
        if not isinstance(r, dict):
            continue
        c = string_to_counts(r["counts"])
        edges = np.concatenate([[0], np.cumsum(c)])
        assert edges[-1] == H * W, f"RLE of frame {t} covers {edges[-1]} pixels, expected {H * W}"
        s, e = edges[1:-1:2], edges[2::2]                       # ones-runs are the odd-indexed counts
        keep = e > s
        plane_ids.append(np.full(int(keep.sum()), t, dtype=np.int32))
        starts.append(s[keep].astype(np.int32))
        ends.append(e[keep].astype(np.int32))
    T = len(rle_masklet)
    n_runs = int(sum(len(s) for s in starts))
    out = P.PackedMasks.empty((T,), H, W, dev)
    Hp = (H + 31) // 32
    scratch = torch.empty((T, W, Hp), dtype=torch.int32, device=dev)
    if n_runs:
        rp = P.to_device(np.concatenate(plane_ids), device=dev)
        rs = P.to_device(np.concatenate(starts), device=dev)
        re = P.to_device(np.concatenate(ends), device=dev)
    else:
        rp = rs = re = None
    with torch.cuda.device(dev):
        _lib.call("sola_rle_decode_runs", P._ptr(rp), P._ptr(rs), P._ptr(re), n_runs, T, H, W, scratch.data_ptr(), out.words.data_ptr(), P._stream(out.words))
    return out


def decode_rle_masklet(rle_masklet: Sequence, device=None) -> np.ndarray:
    """Drop-in for seg_utils.decode_rle_masklet / AlignDataset.rle_masklet_decode: numpy uint8 (T, H, W)."""
    packed = decode_rle_masklet_packed(rle_masklet, device)
    if packed is None:
        return np.stack([np.zeros((0, 0), np.uint8) for _ in rle_masklet], axis=0)
    return P.unpack_masks(packed, torch.uint8).cpu().numpy()


def decode_rle_mask(rle_mask: dict, device=None) -> np.ndarray:
    """seg_utils.decode_rle_mask (seg_utils.py:64-67): one RLE dict -> numpy uint8 (H, W)."""
    return decode_rle_masklet([rle_mask], device)[0]


# ---- encode ----------------------------------------------------------------------------------------------------------

def encode_rle_masklet_packed(packed: P.PackedMasks, cap: int = 1 << 16) -> List[dict]:
    """Packed (T, H, Wp) -> list of {'size': [H, W], 'counts': str} (what encode_rle_masklet_torch returns)."""
    w = packed.words.contiguous()
    T, H, W = packed.n_frames, packed.H, packed.W
    Hp = (H + 31) // 32
    cap = int(min(cap, H * W))
    scratch = torch.empty((T, W, Hp), dtype=torch.int32, device=w.device)
    pos = torch.empty((T, cap), dtype=torch.int32, device=w.device)
    n = torch.empty((T,), dtype=torch.int32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("sola_rle_encode_transitions", w.data_ptr(), T, H, W, scratch.data_ptr(), cap, pos.data_ptr(), n.data_ptr(), P._stream(w))
    n_host = n.cpu().numpy()
    width = int(min(cap, max(1, n_host.max(initial=0))))
    pos_host = pos[:, :width].cpu().numpy()
    out = []
    N = H * W
    for t in range(T):
        k = int(n_host[t])
        if k > cap:                                            # pathological mask (> cap runs): host fallback from the packed bits
            bits = np.unpackbits(w[t].cpu().numpy().view(np.uint8).reshape(H, -1), axis=1, bitorder="little")[:, :W]
            flat = bits.ravel(order="F")
            tr = np.flatnonzero(np.diff(np.concatenate([[0], flat])) != 0)
        else:
            tr = pos_host[t, :k].astype(np.int64)
        counts = np.diff(np.concatenate([[0], tr, [N]]))
        out.append({"size": [H, W], "counts": counts_to_string(counts)})
    return out


def encode_rle_masklet_torch(masks) -> List[dict]:
    """Drop-in for seg_utils.encode_rle_masklet_torch (seg_utils.py:93-106): (N, H, W) {0,1} tensor -> RLE list.
    The masks are packed on the device (or passed already packed) and only run positions cross PCIe."""
    packed = masks if isinstance(masks, P.PackedMasks) else P.pack_masks(masks)
    return encode_rle_masklet_packed(packed)
