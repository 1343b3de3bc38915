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

def _string_to_counts_scalar(s) -> np.ndarray:
    """rleFrString, one character at a time (kept as the plain statement of the format; tests cross-check the vectorised parser)."""
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


def string_to_counts(s) -> np.ndarray:
    """rleFrString vectorised over the whole string (numpy): base-32 little-endian varints with a continuation bit (0x20) and sign
    extension from bit 0x10, then counts[i] += counts[i-2] for i > 2 as two strided running sums.  The reference decodes with
    pycocotools' C loop (dataloader.py:360); a per-character Python loop here would make the host the bottleneck of the GPU path."""
    if isinstance(s, str):
        s = s.encode("ascii")
    c = np.frombuffer(s, dtype=np.uint8).astype(np.int64) - 48
    if c.size == 0:
        return np.zeros(0, dtype=np.int64)
    last = (c & 0x20) == 0                                   # last character of each varint
    ends = np.flatnonzero(last)
    starts = np.concatenate([[0], ends[:-1] + 1])
    k = np.arange(c.size) - np.repeat(starts, ends - starts + 1)             # position of the character inside its varint
    x = np.add.reduceat((c & 0x1F) << (5 * k), starts)
    neg = (c[ends] & 0x10) != 0
    x = np.where(neg, x - (np.int64(1) << (5 * (ends - starts + 1))), x)      # x |= -1 << (5 * n_chars)
    if x.size > 3:                                                             # undo the delta coding against counts[i-2]
        x[3::2] = x[1] + np.cumsum(x[3::2])
        if x.size > 4:
            x[4::2] = x[2] + np.cumsum(x[4::2])
    return x


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

def _runs_of_masklets(rle_masklets: Sequence[Sequence], H: int, W: int):
    """Ones-runs of every frame of several RLE masklets that share planes (frame t of every masklet -> plane t): pinned host int32
    tensors (plane ids, starts, ends).  One call into the library's sequential C parser (`sola_rle_strings_to_runs`, ~1.5 ns per
    character) — the reference's pycocotools decode also parses in C, but then fills H*W bytes per frame on the host."""
    import ctypes as C
    bufs, planes = [], []
    for m in rle_masklets:
        for t, r in enumerate(m):
            if not isinstance(r, dict):
                continue
            assert tuple(r["size"]) == (H, W), "all frames of a masklet share one size"
            cstr = r["counts"]
            bufs.append(cstr.encode("ascii") if isinstance(cstr, str) else bytes(cstr))
            planes.append(t)
    n = len(bufs)
    cap = max(16, sum(len(b) for b in bufs) // 2 + n)            # a ones-run costs at least two characters (its count and the next zeros-run)
    run_p = torch.empty((cap,), dtype=torch.int32).pin_memory() if torch.cuda.is_available() else torch.empty((cap,), dtype=torch.int32)
    run_s, run_e = torch.empty_like(run_p), torch.empty_like(run_p)
    if run_p.is_pinned():
        run_s, run_e = run_s.pin_memory(), run_e.pin_memory()
    strings = (C.c_char_p * max(n, 1))(*bufs)
    lens = (C.c_longlong * max(n, 1))(*[len(b) for b in bufs])
    pids = (C.c_int * max(n, 1))(*planes)
    n_runs = C.c_longlong(0)
    _lib.call("sola_rle_strings_to_runs", C.cast(strings, C.c_void_p), C.cast(lens, C.c_void_p), C.cast(pids, C.c_void_p), n, H * W,
              run_p.data_ptr(), run_s.data_ptr(), run_e.data_ptr(), cap, C.byref(n_runs))
    k = int(n_runs.value)
    return run_p[:k], run_s[:k], run_e[:k]


def _fill_runs(run_p: torch.Tensor, run_s: torch.Tensor, run_e: torch.Tensor, T: int, H: int, W: int, dev) -> P.PackedMasks:
    n_runs = int(run_p.numel())
    out = P.PackedMasks.empty((T,), H, W, dev)
    Hp = (H + 31) // 32
    scratch = torch.empty((T, W, Hp), dtype=torch.int32, device=dev)
    rp = rs = re = None
    if n_runs:
        rp, rs, re = (t.to(dev, non_blocking=True) for t in (run_p, run_s, run_e))
    with torch.cuda.device(dev):
        _lib.call("sola_rle_decode_runs", P._ptr(rp), P._ptr(rs), P._ptr(re), n_runs, T, H, W, scratch.data_ptr(), out.words.data_ptr(), P._stream(out.words))
    return out


def _masklet_shape(rle_masklet: Sequence):
    sizes = [tuple(r["size"]) for r in rle_masklet if isinstance(r, dict)]
    return sizes[-1] if sizes else None


def decode_rle_masklet_packed(rle_masklet: Sequence, device=None) -> Optional[P.PackedMasks]:
    """List of COCO RLE dicts (non-dict entries = missing frames -> empty masks, dataloader.py:364-368) -> packed (T, H, Wp)."""
    shape = _masklet_shape(rle_masklet)
    if shape is None:
        return None
    H, W = shape
    dev = P._dev(device)
    return _fill_runs(*_runs_of_masklets([rle_masklet], H, W), len(rle_masklet), H, W, dev)


def decode_rle_masklets_merged(rle_masklets: Sequence[Sequence], device=None) -> Optional[P.PackedMasks]:
    """OR of several RLE masklets of one video (the selected SAM2 tracks of dataloader.py:339-344, or an expression's GT objects,
    dataloader.py:285-299) straight into ONE packed (T, H, Wp) masklet: the ones-runs of every track are filled into the same
    planes with atomicOr, so decode and merge are one launch and no per-track plane ever exists."""
    live = [m for m in rle_masklets if _masklet_shape(m) is not None]
    if not live:
        return None
    H, W = _masklet_shape(live[0])
    T = len(live[0])
    assert all(len(m) == T and _masklet_shape(m) == (H, W) for m in live), "tracks of one video share (T, H, W)"
    return _fill_runs(*_runs_of_masklets(live, H, W), T, H, W, P._dev(device))


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
