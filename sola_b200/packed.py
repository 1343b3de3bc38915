"""Batched device-side entry points over bit-packed mask planes (the reference has no equivalent; the
per-call drop-in wrappers in seg_utils.py / utils.py / evaluator.py / prompt_generator.py route here).

PyTorch is used for device memory and streams only; every computation is a kernel of libsola_maskpath.so
reached through the C ABI (include/sola_maskpath.h).  There is no CPU path: CPU inputs are copied to the
current CUDA device, and a missing library raises.

Packed layout: (..., H, Wp) int32 words (bit pattern of uint32), Wp = ceil(W/32), bit b of word w = pixel
32*w + b, pad bits zero.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def words_per_row(W: int) -> int:
    return (W + 31) // 32


def _dev(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _lib.SolaError("sola_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(x, device=None, dtype=None) -> torch.Tensor:
    """numpy / CPU tensor / CUDA tensor -> contiguous CUDA tensor (H2D copy when needed)."""
    if isinstance(x, np.ndarray):
        x = np.ascontiguousarray(x)
        if x.dtype == np.bool_:
            x = x.view(np.uint8)
        x = torch.from_numpy(x)
    if x.dtype == torch.bool:
        x = x.view(torch.uint8) if x.is_contiguous() else x.contiguous().view(torch.uint8)
    if not x.is_cuda:
        x = x.to(_dev(device), non_blocking=True)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass
class PackedMasks:
    """Bit-packed mask planes on the device: `words` is (..., H, Wp) int32."""
    words: torch.Tensor
    H: int
    W: int

    @property
    def Wp(self) -> int:
        return words_per_row(self.W)

    @property
    def frame_words(self) -> int:
        return self.H * self.Wp

    @property
    def lead_shape(self) -> Tuple[int, ...]:
        return tuple(self.words.shape[:-2])

    @property
    def n_frames(self) -> int:
        return int(np.prod(self.lead_shape, dtype=np.int64)) if self.lead_shape else 1

    @property
    def device(self) -> torch.device:
        return self.words.device

    def __getitem__(self, idx) -> "PackedMasks":
        w = self.words[idx]
        assert w.dim() >= 2 and w.shape[-2:] == self.words.shape[-2:], "index only the leading (track/frame) axes"
        return PackedMasks(w, self.H, self.W)

    def reshape_lead(self, *lead) -> "PackedMasks":
        return PackedMasks(self.words.reshape(*lead, self.H, self.Wp), self.H, self.W)

    def numpy_u32(self) -> np.ndarray:
        return self.words.cpu().numpy().view(np.uint32)

    @staticmethod
    def from_numpy_u32(arr: np.ndarray, H: int, W: int, device=None) -> "PackedMasks":
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int32)).to(_dev(device))
        return PackedMasks(t, H, W)

    @staticmethod
    def empty(lead: Sequence[int], H: int, W: int, device=None) -> "PackedMasks":
        return PackedMasks(torch.empty((*lead, H, words_per_row(W)), dtype=torch.int32, device=_dev(device)), H, W)


# ---------------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------------

def binarize_pack_stability(logits, mask_threshold: float = 0.0, threshold_offset: float = 1.0, *,
                            want_packed: bool = True, want_counts: bool = True,
                            out: Optional[PackedMasks] = None, counts_out: Optional[torch.Tensor] = None):
    """One pass over fp32 / bf16 logits (..., H, W): packed `logit > thr` planes plus, per frame,
    counts[0] = #(> thr+off), counts[1] = #(> thr), counts[2] = #(> thr-off)  (int32, device).
    Replaces `(out_mask_logits > 0.0).float()` + `torch.cat` (generate_tokens_grid.py:215-224) and
    PromptGenerator.get_stability_score (prompt_generator.py:169-186)."""
    x = to_device(logits)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    assert x.dim() >= 2, "logits must be (..., H, W)"
    H, W = int(x.shape[-2]), int(x.shape[-1])
    lead = tuple(x.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    packed = None
    if want_packed:
        packed = out if out is not None else PackedMasks.empty(lead, H, W, x.device)
        assert packed.words.is_contiguous() and packed.n_frames == n and (packed.H, packed.W) == (H, W)
    counts = None
    if want_counts:
        counts = counts_out if counts_out is not None else torch.empty((3, n), dtype=torch.int32, device=x.device)
        assert counts.is_contiguous() and counts.shape == (3, n) and counts.dtype == torch.int32
    fn = "sola_binarize_pack_f32" if x.dtype == torch.float32 else "sola_binarize_pack_bf16"
    with torch.cuda.device(x.device):
        _lib.call(fn, x.data_ptr(), n, H, W, float(mask_threshold), float(threshold_offset),
                  _ptr(packed.words) if packed is not None else None,
                  _ptr(counts[0]) if counts is not None else None,
                  _ptr(counts[1]) if counts is not None else None,
                  _ptr(counts[2]) if counts is not None else None, _stream(x))
    if counts is not None:
        counts = counts.view(3, *lead)
    return packed, counts


def binarize_pack_resize(logits, mask_threshold: float = 0.0, threshold_offset: float = 1.0, target_shape=None, *,
                         want_packed: bool = True, want_area: bool = False,
                         out: Optional[PackedMasks] = None, counts_out: Optional[torch.Tensor] = None,
                         resized_out: Optional[PackedMasks] = None):
    """K1 + R1 in one pass over the logits: (full-resolution packed planes | None, counts (3, ...), resized packed planes
    [, per-frame areas of the resized planes]).  Bit-identical to binarize_pack_stability followed by resize_bilinear_bin;
    the resize work hides under the HBM time of reading the logits."""
    x = to_device(logits)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    H, W = int(x.shape[-2]), int(x.shape[-1])
    oh, ow = default_target_shape(H, W) if target_shape is None else (int(target_shape[0]), int(target_shape[1]))
    lead = tuple(x.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    fusable = (W % 32 == 0) and (x.data_ptr() % 16 == 0)
    packed = None
    if want_packed or not fusable:
        packed = out if out is not None else PackedMasks.empty(lead, H, W, x.device)
        assert packed.words.is_contiguous() and packed.n_frames == n and (packed.H, packed.W) == (H, W), \
            f"out= must be contiguous ({n} frames, {H}x{W}); got {packed.n_frames} frames, {packed.H}x{packed.W}"
        assert packed.words.device == x.device
    resized = resized_out if resized_out is not None else PackedMasks.empty(lead, oh, ow, x.device)
    assert resized.words.is_contiguous() and resized.n_frames == n and (resized.H, resized.W) == (oh, ow)
    counts = counts_out if counts_out is not None else torch.empty((3, n), dtype=torch.int32, device=x.device)
    assert counts.is_contiguous() and counts.shape == (3, n) and counts.dtype == torch.int32
    area = torch.empty((n,), dtype=torch.int32, device=x.device) if want_area else None
    fn = "sola_binarize_pack_resize_f32" if x.dtype == torch.float32 else "sola_binarize_pack_resize_bf16"
    with torch.cuda.device(x.device):
        _lib.call(fn, x.data_ptr(), n, H, W, oh, ow, float(mask_threshold), float(threshold_offset),
                  _ptr(packed.words) if packed is not None else None, resized.words.data_ptr(),
                  counts[0].data_ptr(), counts[1].data_ptr(), counts[2].data_ptr(), _ptr(area), _stream(x))
    counts = counts.view(3, *lead)
    if want_area:
        return packed, counts, resized, area.view(*lead) if lead else area
    return packed, counts, resized


def stability_from_counts(counts) -> np.ndarray:
    """float64 hi/lo with numpy's 0/0 -> nan (prompt_generator.py:186)."""
    c = counts.cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    with np.errstate(divide="ignore", invalid="ignore"):
        return c[0].astype(np.int32) / c[2].astype(np.int32)


def pack_masks(mask, *, want_area: bool = False, threshold: Optional[float] = None):
    """{0,1} masks (fp32 / uint8 / bool, nonzero = foreground) of shape (..., H, W) -> PackedMasks
    (and int32 per-frame areas).  With `threshold`, fp32 input is binarised as `x > threshold` instead."""
    x = to_device(mask)
    if x.dtype not in (torch.float32, torch.uint8):
        x = x.float() if x.dtype.is_floating_point else x.to(torch.uint8)
    H, W = int(x.shape[-2]), int(x.shape[-1])
    lead = tuple(x.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    packed = PackedMasks.empty(lead, H, W, x.device)
    area = torch.empty((n,), dtype=torch.int32, device=x.device) if want_area else None
    with torch.cuda.device(x.device):
        if threshold is not None:
            assert x.dtype == torch.float32
            _lib.call("sola_threshold_pack_f32", x.data_ptr(), n, H, W, float(threshold), packed.words.data_ptr(), _ptr(area), _stream(x))
        elif x.dtype == torch.float32:
            _lib.call("sola_pack_mask_f32", x.data_ptr(), n, H, W, packed.words.data_ptr(), _ptr(area), _stream(x))
        else:
            _lib.call("sola_pack_mask_u8", x.data_ptr(), n, H, W, packed.words.data_ptr(), _ptr(area), _stream(x))
    if want_area:
        return packed, (area.view(*lead) if lead else area)
    return packed


def unpack_masks(packed: PackedMasks, dtype=torch.float32, one_value: int = 1) -> torch.Tensor:
    """PackedMasks -> (..., H, W) {0,1} tensor of fp32 or uint8 on the device (uint8: foreground = `one_value`, e.g. 255
    for the PNG planes of inference.py:89-91)."""
    assert dtype in (torch.float32, torch.uint8)
    w = packed.words.contiguous()
    out = torch.empty((*packed.lead_shape, packed.H, packed.W), dtype=dtype, device=w.device)
    with torch.cuda.device(w.device):
        if dtype == torch.uint8 and one_value != 1:
            _lib.call("sola_unpack_u8_value", w.data_ptr(), packed.n_frames, packed.H, packed.W, int(one_value), out.data_ptr(), _stream(w))
        else:
            _lib.call("sola_unpack_f32" if dtype == torch.float32 else "sola_unpack_u8",
                      w.data_ptr(), packed.n_frames, packed.H, packed.W, out.data_ptr(), _stream(w))
    return out


# ---------------------------------------------------------------------------------------------------------
# K3
# ---------------------------------------------------------------------------------------------------------

def frame_counts(a, b) -> torch.Tensor:
    """Raw {0,1} planes a, b of identical shape (T, H, W) or (H, W), fp32 or uint8 -> int32 (3, T) device
    tensor [inter, area_a, area_b] per frame, in ONE pass over both inputs."""
    # both operands on ONE device: a CUDA input's device wins, else the current device
    dev = next((t.device for t in (a, b) if isinstance(t, torch.Tensor) and t.is_cuda), None)
    a, b = to_device(a, device=dev), to_device(b, device=dev)
    if b.device != a.device:
        b = b.to(a.device)
    assert a.shape == b.shape, f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}"
    if a.dtype != b.dtype or a.dtype not in (torch.float32, torch.uint8):
        both_u8 = a.dtype == torch.uint8 and b.dtype == torch.uint8
        a, b = (a, b) if both_u8 else (a.float(), b.float())
    if a.dim() == 2:
        a, b = a[None], b[None]
    T = int(a.shape[0])
    frame_px = int(np.prod(a.shape[1:], dtype=np.int64))
    out = torch.empty((3, T), dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("sola_frame_counts_f32" if a.dtype == torch.float32 else "sola_frame_counts_u8",
                  a.data_ptr(), b.data_ptr(), T, frame_px, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _stream(a))
    return out


def frame_counts_packed(A: PackedMasks, B: PackedMasks):
    """A (Na, T, H, Wp), B (Nb, T, H, Wp) -> inter (Na, Nb, T), area_a (Na, T), area_b (Nb, T) int32 device tensors."""
    assert (A.H, A.W) == (B.H, B.W)
    aw, bw = A.words.contiguous(), B.words.contiguous()
    if aw.dim() == 3:
        aw = aw[None]
    if bw.dim() == 3:
        bw = bw[None]
    Na, T = int(aw.shape[0]), int(aw.shape[1])
    Nb = int(bw.shape[0])
    assert int(bw.shape[1]) == T, "frame counts differ"
    inter = torch.empty((Na, Nb, T), dtype=torch.int32, device=aw.device)
    area_a = torch.empty((Na, T), dtype=torch.int32, device=aw.device)
    area_b = torch.empty((Nb, T), dtype=torch.int32, device=aw.device)
    with torch.cuda.device(aw.device):
        _lib.call("sola_frame_counts_packed", aw.data_ptr(), bw.data_ptr(), Na, Nb, T, A.frame_words,
                  inter.data_ptr(), area_a.data_ptr(), area_b.data_ptr(), _stream(aw))
    return inter, area_a, area_b


def frame_counts_ragged(a_words: torch.Tensor, b_words: torch.Tensor, word_offsets: torch.Tensor) -> torch.Tensor:
    """Flat int32 word buffers + int64 device offsets [n_frames+1] -> int32 (3, n_frames) [inter, area_a, area_b]."""
    n = int(word_offsets.numel()) - 1
    out = torch.empty((3, max(n, 0)), dtype=torch.int32, device=a_words.device)
    if n <= 0:
        return out
    assert word_offsets.dtype == torch.int64 and word_offsets.is_cuda
    with torch.cuda.device(a_words.device):
        _lib.call("sola_frame_counts_packed_ragged", a_words.data_ptr(), b_words.data_ptr(), word_offsets.data_ptr(), n,
                  out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), _stream(a_words))
    return out


def jf_accumulators(pred, gt):
    """The J&F accumulators of one (video, expression) unit, on the device: (inter int32 [T], union int32 [T], int64 [tp, fp, fn]).
    pred / gt: raw {0,1} masklets (T, H, W) fp32 or uint8, or PackedMasks (T, H, Wp).  J = mean(union ? inter / union : 1)
    (evaluator.py:227-237); F from the exact volume sums (evaluator.py:239-247) — see evaluator.jf_from_accumulators."""
    if isinstance(pred, PackedMasks):
        assert isinstance(gt, PackedMasks) and (pred.H, pred.W) == (gt.H, gt.W)
        a, b = pred.words.contiguous(), gt.words.contiguous()
        T, per_frame, fn = pred.n_frames, pred.frame_words, "sola_jf_packed"
    else:
        a = to_device(pred)
        b = to_device(gt, device=a.device)
        assert a.shape == b.shape, f"shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}"
        if not (a.dtype == torch.uint8 and b.dtype == torch.uint8):
            a, b = a.float(), b.float()
        if a.dim() == 2:
            a, b = a[None], b[None]
        a, b = a.contiguous(), b.contiguous()
        T, per_frame = int(a.shape[0]), int(np.prod(a.shape[1:], dtype=np.int64))
        fn = "sola_jf_f32" if a.dtype == torch.float32 else "sola_jf_u8"
    inter = torch.empty((T,), dtype=torch.int32, device=a.device)
    uni = torch.empty((T,), dtype=torch.int32, device=a.device)
    totals = torch.empty((3,), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call(fn, a.data_ptr(), b.data_ptr(), T, per_frame, inter.data_ptr(), uni.data_ptr(), totals.data_ptr(), _stream(a))
    return inter, uni, totals


def or_merge(tracks: PackedMasks, select=None) -> PackedMasks:
    """tracks (K, T, H, Wp) -> (T, H, Wp): OR of the tracks with select[k] != 0 (all when select is None).
    With nothing selected the result is all-zero planes (dataloader.py:346-349)."""
    w = tracks.words.contiguous()
    K = int(w.shape[0])
    words = int(w[0].numel()) if K else 0
    out = PackedMasks(torch.empty(w.shape[1:], dtype=torch.int32, device=w.device), tracks.H, tracks.W)
    sel = None
    if select is not None:
        sel = to_device(np.asarray([1 if s else 0 for s in (select.tolist() if hasattr(select, "tolist") else select)], dtype=np.uint8), device=w.device)
        assert sel.numel() == K
    with torch.cuda.device(w.device):
        _lib.call("sola_or_merge", w.data_ptr(), _ptr(sel), K, words, out.words.data_ptr(), _stream(w))
    return out


# ---------------------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------------------

def pairwise_inter_matrix(tracks: PackedMasks) -> torch.Tensor:
    """tracks (N, T, H, Wp) -> int64 (N, N) spatio-temporal intersection counts; diagonal = areas.
    Exact-integer version of seg_utils.compute_masklet_iou (seg_utils.py:110-125) for every pair."""
    w = tracks.words.contiguous()
    N = int(w.shape[0])
    words = int(w[0].numel()) if N else 1
    inter = torch.empty((N, N), dtype=torch.int64, device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("sola_pair_iou_st", w.data_ptr(), N, words, inter.data_ptr(), None, _stream(w))
    return inter


def pairwise_inter_matrix_words(w: torch.Tensor) -> torch.Tensor:
    """(N, words) int32 rows of packed bits (any slice of the word axis of N tracks) -> int64 (N, N) intersection counts."""
    assert w.dim() == 2 and w.dtype == torch.int32 and w.is_cuda and w.is_contiguous()
    N = int(w.shape[0])
    inter = torch.empty((N, N), dtype=torch.int64, device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("sola_pair_iou_st", w.data_ptr(), N, max(int(w.shape[1]), 1), inter.data_ptr(), None, _stream(w))
    return inter


def pairwise_inter_matrix_part(tracks: PackedMasks, part: int, n_parts: int) -> torch.Tensor:
    """This part's share of the N x N intersection matrix (pair tiles part, part + n_parts, ...; zeros elsewhere)."""
    w = tracks.words.contiguous()
    N = int(w.shape[0])
    words = int(w[0].numel()) if N else 1
    inter = torch.empty((N, N), dtype=torch.int64, device=w.device)
    with torch.cuda.device(w.device):
        _lib.call("sola_pair_iou_st_part", w.data_ptr(), N, words, int(part), int(n_parts), inter.data_ptr(), _stream(w))
    return inter


def pairwise_inter_matrix_rows(row_ptrs: torch.Tensor, words_per_track: int, part: int = 0, n_parts: int = 1) -> torch.Tensor:
    """N x N intersection share from a device int64 table of per-track plane pointers (tracks may live on peer GPUs)."""
    assert row_ptrs.dtype == torch.int64 and row_ptrs.is_cuda and row_ptrs.is_contiguous()
    N = int(row_ptrs.numel())
    inter = torch.empty((N, N), dtype=torch.int64, device=row_ptrs.device)
    with torch.cuda.device(row_ptrs.device):
        _lib.call("sola_pair_iou_st_rows", row_ptrs.data_ptr(), N, int(words_per_track), int(part), int(n_parts), inter.data_ptr(), _stream(row_ptrs))
    return inter


def pairwise_inter_matrix_peer(bases: Sequence[int], n_local: int, words_per_track: int, part: int, n_parts: int, device) -> torch.Tensor:
    """One K2 launch whose TMA producer reads every rank's (n_local, words) planes in place over NVLink —
    `bases` are the peer-mapped device addresses of the ranks' buffers.  Returns this part's (N, N) int64 share."""
    import ctypes
    world = len(bases)
    arr = (ctypes.c_void_p * world)(*[int(b) for b in bases])
    N = world * int(n_local)
    inter = torch.empty((N, N), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.call("sola_pair_iou_st_peer", ctypes.cast(arr, ctypes.c_void_p), world, int(n_local), int(words_per_track), int(part), int(n_parts),
                  inter.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    return inter


def pull_rows(row_ptrs: torch.Tensor, word_lo: int, n_words: int, out: torch.Tensor) -> torch.Tensor:
    """out (N, n_words) int32 <- words [word_lo, word_lo + n_words) of every row of the pointer table (peer rows: over NVLink)."""
    assert row_ptrs.dtype == torch.int64 and row_ptrs.is_cuda and row_ptrs.is_contiguous()
    N = int(row_ptrs.numel())
    assert out.dtype == torch.int32 and out.is_contiguous() and out.numel() == N * n_words and out.device == row_ptrs.device
    with torch.cuda.device(out.device):
        _lib.call("sola_pull_rows", row_ptrs.data_ptr(), N, int(word_lo), int(n_words), out.data_ptr(), _stream(out))
    return out


def pairwise_inter_accumulate(w: torch.Tensor, inter: torch.Tensor) -> torch.Tensor:
    """inter (N, N) int64 += intersections over the (N, words) int32 rows `w` (a chunk of the word axis)."""
    assert w.dim() == 2 and w.dtype == torch.int32 and w.is_cuda and w.is_contiguous()
    N = int(w.shape[0])
    assert inter.shape == (N, N) and inter.dtype == torch.int64 and inter.is_contiguous() and inter.device == w.device
    with torch.cuda.device(w.device):
        _lib.call("sola_pair_iou_st_accumulate", w.data_ptr(), N, int(w.shape[1]), inter.data_ptr(), _stream(w))
    return inter


def iou_matrix_from_inter(inter) -> np.ndarray:
    """float64 IoU matrix from the integer intersection matrix: union = a_i + a_j - inter; empty union -> 1.0."""
    m = inter.cpu().numpy().astype(np.int64) if isinstance(inter, torch.Tensor) else np.asarray(inter, dtype=np.int64)
    area = np.diag(m)
    union = area[:, None] + area[None, :] - m
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = m / union
    iou[union == 0] = 1.0
    return iou


def gathered_inter(tracks: PackedMasks, prompts: PackedMasks, frame_idx) -> torch.Tensor:
    """tracks (N, T, h, wp), prompts (P, h, wp), frame_idx [P] -> int32 (3, N, P) device tensor:
    [0] inter[i, j] = |track_i[frame_idx[j]] & prompt_j|, [1] area of that track frame, [2] prompt area (row-broadcast)."""
    assert (tracks.H, tracks.W) == (prompts.H, prompts.W)
    tw, pw = tracks.words.contiguous(), prompts.words.contiguous()
    if tw.dim() == 3:
        tw = tw[None]
    N, T = int(tw.shape[0]), int(tw.shape[1])
    P = int(pw.shape[0])
    if not (isinstance(frame_idx, torch.Tensor) and frame_idx.is_cuda):
        # the reference indexes masklets[prompt_id][frame_idx] (generate_tokens_grid.py:273): out of range raises there, so it must
        # here too (device-resident index tensors are validated by their owner, dedup.GreedyState; the kernel clamps for memory safety)
        fi_host = np.asarray(frame_idx.cpu() if isinstance(frame_idx, torch.Tensor) else frame_idx, dtype=np.int64)
        if fi_host.size and (fi_host.min() < 0 or fi_host.max() >= T):
            raise IndexError(f"frame_idx out of range for masklets of {T} frames: min {fi_host.min()}, max {fi_host.max()}")
    fi = to_device(np.asarray(frame_idx, dtype=np.int32) if not isinstance(frame_idx, torch.Tensor) else frame_idx.to(torch.int32), device=tw.device)
    assert fi.numel() == P
    out = torch.empty((3, N, P), dtype=torch.int32, device=tw.device)
    area_p = torch.empty((P,), dtype=torch.int32, device=tw.device)
    with torch.cuda.device(tw.device):
        _lib.call("sola_pair_iou_gather", tw.data_ptr(), pw.data_ptr(), fi.data_ptr(), N, P, T, tracks.frame_words,
                  out[0].data_ptr(), out[1].data_ptr(), area_p.data_ptr(), _stream(tw))
    out[2] = area_p[None, :]
    return out


# ---------------------------------------------------------------------------------------------------------
# R1 / R2
# ---------------------------------------------------------------------------------------------------------

def default_target_shape(H: int, W: int) -> Tuple[int, int]:
    """seg_utils.py:154-155."""
    return (540, 960) if H < W else (960, 540)


def resize_bilinear_bin(packed: PackedMasks, target_shape=None, *, want_area: bool = False):
    """Packed (..., H, Wp) -> packed (..., oh, owp): `interpolate(bilinear) > 0.5` of seg_utils.reshape_masklet."""
    oh, ow = default_target_shape(packed.H, packed.W) if target_shape is None else (int(target_shape[0]), int(target_shape[1]))
    w = packed.words.contiguous()
    out = PackedMasks.empty(packed.lead_shape, oh, ow, w.device)
    n = packed.n_frames
    area = torch.empty((n,), dtype=torch.int32, device=w.device) if want_area else None
    with torch.cuda.device(w.device):
        _lib.call("sola_resize_bilinear_bin_packed", w.data_ptr(), n, packed.H, packed.W, oh, ow, out.words.data_ptr(), _ptr(area), _stream(w))
    return (out, area) if want_area else out


def resize_bilinear_bin_f32(x, target_shape=None, *, want_packed: bool = True, want_f32: bool = False):
    """fp32 planes (..., H, W) of arbitrary values -> packed and/or fp32 {0,1} planes at the target shape."""
    x = to_device(x, dtype=torch.float32)
    H, W = int(x.shape[-2]), int(x.shape[-1])
    oh, ow = default_target_shape(H, W) if target_shape is None else (int(target_shape[0]), int(target_shape[1]))
    lead = tuple(x.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    packed = PackedMasks.empty(lead, oh, ow, x.device) if want_packed else None
    f32 = torch.empty((*lead, oh, ow), dtype=torch.float32, device=x.device) if want_f32 else None
    with torch.cuda.device(x.device):
        _lib.call("sola_resize_bilinear_bin_f32", x.data_ptr(), n, H, W, oh, ow,
                  _ptr(packed.words) if packed is not None else None, _ptr(f32), None, _stream(x))
    return packed, f32


def resize_nearest(mask, oh: int, ow: int) -> PackedMasks:
    """uint8 / bool masks (..., H, W) or PackedMasks -> packed (..., oh, owp) with ATen's legacy 'nearest' indexing."""
    if isinstance(mask, PackedMasks):
        w = mask.words.contiguous()
        out = PackedMasks.empty(mask.lead_shape, oh, ow, w.device)
        with torch.cuda.device(w.device):
            _lib.call("sola_resize_nearest_packed", w.data_ptr(), mask.n_frames, mask.H, mask.W, oh, ow, out.words.data_ptr(), None, _stream(w))
        return out
    x = to_device(mask)
    if x.dtype != torch.uint8:
        x = (x != 0).to(torch.uint8)
    H, W = int(x.shape[-2]), int(x.shape[-1])
    lead = tuple(x.shape[:-2])
    n = int(np.prod(lead, dtype=np.int64)) if lead else 1
    out = PackedMasks.empty(lead, oh, ow, x.device)
    with torch.cuda.device(x.device):
        _lib.call("sola_resize_nearest_u8", x.data_ptr(), n, H, W, oh, ow, out.words.data_ptr(), None, _stream(x))
    return out


# ---------------------------------------------------------------------------------------------------------
# fused J & F (region counts + boundary-match counts)
# ---------------------------------------------------------------------------------------------------------

def bound_pix_for(H: int, W: int, bound_th: float = 0.008) -> int:
    """DAVIS rule: bound_th >= 1 is a pixel radius, else a fraction of the frame diagonal (ceil)."""
    import math
    return int(bound_th) if bound_th >= 1 else int(math.ceil(bound_th * math.sqrt(H * H + W * W)))


import ctypes as _C


class JfUnit(_C.Structure):
    """ctypes mirror of `sola_jf_unit` (include/sola_maskpath.h): 64 bytes."""
    _fields_ = [("pred", _C.c_void_p), ("gt", _C.c_void_p), ("out_off", _C.c_longlong), ("item0", _C.c_longlong),
                ("T", _C.c_int), ("H", _C.c_int), ("W", _C.c_int), ("radius", _C.c_int),
                ("band_rows", _C.c_int), ("n_bands", _C.c_int), ("reserved0", _C.c_int), ("reserved1", _C.c_int)]


class JfPlan(_C.Structure):
    """ctypes mirror of `sola_jf_plan`: 32 bytes."""
    _fields_ = [("n_items", _C.c_longlong), ("total_frames", _C.c_longlong), ("raw_cap", _C.c_int), ("bm_cap", _C.c_int),
                ("mask_steps", _C.c_int), ("reserved", _C.c_int)]


assert _C.sizeof(JfUnit) == 64 and _C.sizeof(JfPlan) == 32


class _JFLaunch:
    """One planned launch of the fused kernel over a list of (pred words, gt words, T, H, W, radius) units."""

    def __init__(self, units, device, ctas_per_sm: int = 0):
        n = len(units)
        arr = (JfUnit * max(n, 1))()
        for k, (pw, gw, T, H, W, radius) in enumerate(units):
            u = arr[k]
            u.pred, u.gt, u.T, u.H, u.W, u.radius = pw.data_ptr(), gw.data_ptr(), T, H, W, radius
        self.n_units, self.device = n, device
        self.plan = JfPlan()
        self.plan.reserved = int(ctas_per_sm)                  # 0 = let the library pick the tile class
        _lib.call("sola_jf_sweep_plan", _C.cast(arr, _C.c_void_p), n, _C.byref(self.plan))
        self.offsets = [arr[k].out_off for k in range(n)]
        self.bands = [(arr[k].band_rows, arr[k].n_bands) for k in range(n)]
        self.smem_bytes = (2 * self.plan.raw_cap + 2 * self.plan.bm_cap + 16 * self.plan.mask_steps) * 4
        self.units_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)[: 64 * n].clone().to(device) if n else None

    def run(self, out: torch.Tensor) -> None:
        if self.n_units == 0 or self.plan.total_frames == 0:
            return
        with torch.cuda.device(self.device):
            _lib.call("sola_jf_sweep", self.units_dev.data_ptr(), self.n_units, _C.byref(self.plan), out.data_ptr(),
                      torch.cuda.current_stream(self.device).cuda_stream)


JF_SMALL_TILE_BYTES = 100 * 1024          # the library's two-CTAs-per-SM tile budget (csrc/jf_fused.cu: JF_BUDGET_2CTA)


class JFSweepPlan:
    """A planned J&F sweep over units of different shapes through the fused kernel.

        plan = JFSweepPlan([(pred0, gt0), (pred1, gt1), ...], with_boundary=True)      # PackedMasks pairs, (T_u, H_u, Wp_u) each
        counts = plan.run()          # int32 (7, total_frames) on the device; plan.offsets[u] = first column of unit u, plan.frames[u] columns

    All units go into ONE launch (frames up to 1080p with the DAVIS disk fit the two-CTAs-per-SM tile).  Only frames whose tile needs a
    whole SM's shared memory (beyond ~1080p) form a second class, launched separately so that they do not halve the occupancy of every
    other unit — at most two launches per sweep.
    The plan keeps references to the planes; run() may be called repeatedly (the bench times it)."""

    def __init__(self, pairs: Sequence[Tuple["PackedMasks", "PackedMasks"]], with_boundary: bool = True, bound_th: float = 0.008,
                 ctas_per_sm: int = 0):
        self.pairs, self.device = [], None
        units = []
        for p, g in pairs:
            assert isinstance(p, PackedMasks) and isinstance(g, PackedMasks)
            assert (p.H, p.W) == (g.H, g.W) and p.n_frames == g.n_frames, "pred / gt shape mismatch"
            pw, gw = p.words.contiguous(), g.words.contiguous()
            assert pw.is_cuda and gw.device == pw.device
            if self.device is None:
                self.device = pw.device
            assert pw.device == self.device, "all units of a sweep must live on one device"
            self.pairs.append((pw, gw))
            units.append((pw, gw, p.n_frames, p.H, p.W, bound_pix_for(p.H, p.W, bound_th) if with_boundary else -1))
        self.n_units = len(units)
        self.frames = [u[2] for u in units]
        # class of every unit: does its tile fit the two-CTAs-per-SM budget?  (planned alone: a host-only call)
        small, big = [], []
        for k, u in enumerate(units):
            alone = _JFLaunch([(u[0], u[1], 1, u[3], u[4], u[5])], self.device)
            (small if alone.smem_bytes <= JF_SMALL_TILE_BYTES else big).append(k)
        self.launches, self.offsets, base = [], [0] * self.n_units, 0
        for group in (small, big):
            if not group:
                continue
            L = _JFLaunch([units[k] for k in group], self.device, ctas_per_sm)
            for j, k in enumerate(group):
                self.offsets[k] = base + L.offsets[j]
            self.launches.append((L, base))
            base += L.plan.total_frames
        self.total_frames = base
        self.n_items = sum(L.plan.n_items for L, _ in self.launches)
        self.smem_bytes = max([L.smem_bytes for L, _ in self.launches], default=0)
        self.bands = [None] * self.n_units
        for (L, _), group in zip(self.launches, [g for g in (small, big) if g]):
            for j, k in enumerate(group):
                self.bands[k] = L.bands[j]
        self.algorithmic_bytes = sum(2 * pw.numel() * 4 for pw, _ in self.pairs)
        self._parts = None

    def run(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = self.device if self.device is not None else _dev()
        if out is None:
            out = torch.empty((7, self.total_frames), dtype=torch.int32, device=dev)
        assert out.shape == (7, self.total_frames) and out.dtype == torch.int32 and out.is_contiguous()
        if len(self.launches) == 1:
            self.launches[0][0].run(out)
            return out
        # two launches: each writes its own (7, n) table (the kernel's row stride is its launch's frame count), then one strided copy each
        if self._parts is None:
            self._parts = [torch.empty((7, L.plan.total_frames), dtype=torch.int32, device=dev) for L, _ in self.launches]
        for (L, base), part in zip(self.launches, self._parts):
            L.run(part)
            out[:, base: base + L.plan.total_frames].copy_(part)
        return out


def jf_boundary_counts(pred: PackedMasks, gt: PackedMasks, bound_th: float = 0.008, with_boundary: bool = True) -> torch.Tensor:
    """One unit through the fused J&F kernel: int32 (7, n_frames) device tensor
    [|pred ∩ gt|, |pred|, |gt|, |b(pred)|, |b(gt)|, fg_match, gt_match] (rows 3..6 zero when with_boundary is False)."""
    assert (pred.H, pred.W) == (gt.H, gt.W) and pred.n_frames == gt.n_frames
    pw, gw = pred.words.contiguous(), gt.words.contiguous()
    n = pred.n_frames
    r = bound_pix_for(pred.H, pred.W, bound_th) if with_boundary else -1
    out = torch.empty((7, n), dtype=torch.int32, device=pw.device)
    with torch.cuda.device(pw.device):
        _lib.call("sola_jf_boundary_packed", pw.data_ptr(), gw.data_ptr(), n, pred.H, pred.W, r, out.data_ptr(), _stream(pw))
    return out


def boundary_counts(pred: PackedMasks, gt: PackedMasks, bound_th: float = 0.008) -> torch.Tensor:
    """int32 (4, n_frames) device tensor: n_fg, n_gt, fg_match, gt_match (DAVIS boundary measure)."""
    return jf_boundary_counts(pred, gt, bound_th)[3:7]
