"""Device versions of the two OR-merges of `dataloader.py` that feed J&F (AlignDataset.get_sam2_masklet :305-351,
get_gt_masklet :278-303), with the file I/O and RLE decoding left to the caller."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import packed as P


def merge_selected_tracks(masklets: Optional[P.PackedMasks], preds: Sequence[float]) -> Optional[P.PackedMasks]:
    """masklets (K, T, H, Wp) in directory order, preds[k] > 0 selects.  No tracks at all -> None (the reference
    returns None when the directories are empty); nothing selected -> all-zero planes of the tracks' shape."""
    if masklets is None or masklets.words.shape[0] == 0:
        return None
    preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
    return P.or_merge(masklets, select=(preds[: masklets.words.shape[0]] > 0))


def merge_gt_objects(masklets: P.PackedMasks) -> P.PackedMasks:
    """OR over the expression's GT objects."""
    return P.or_merge(masklets, select=None)


def get_sam2_masklet_packed(rle_masklets: Sequence, preds: Sequence[float], device=None) -> Optional[P.PackedMasks]:
    """Device version of AlignDataset.get_sam2_masklet (dataloader.py:305-351) once the per-track JSON files are read:
    `rle_masklets[k]` is track k's `sam2_masklet_info['rle']` list in directory order.  RLE decode, selection and OR-merge all
    happen on packed bits; the result feeds evaluator / JFSweep directly (no uint8 (T,H,W) arrays, no fp32 H2D copy)."""
    from . import rle
    if len(rle_masklets) == 0:
        return None
    preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
    selected = [k for k in range(len(rle_masklets)) if preds[k] > 0]
    if not selected:
        first = rle.decode_rle_masklet_packed(rle_masklets[0], device)          # shape donor: zeros of the first track's shape (:346-349)
        return P.PackedMasks(torch.zeros_like(first.words), first.H, first.W)
    planes = [rle.decode_rle_masklet_packed(rle_masklets[k], device) for k in selected]
    stacked = P.PackedMasks(torch.stack([p.words for p in planes]), planes[0].H, planes[0].W)
    return P.or_merge(stacked)


def png_planes(masklet: P.PackedMasks) -> np.ndarray:
    """(T, H, W) uint8 planes with foreground 255, as inference.py:89-91 writes them (`(mask * 255).astype(np.uint8)`)."""
    return P.unpack_masks(masklet, torch.uint8, one_value=255).cpu().numpy()


SIDECAR_SUFFIX = ".packed.npz"


def save_packed_masklet(path: str, masklet: P.PackedMasks) -> None:
    """Packed sidecar format: the uint32 planes + shape, 32x smaller than uint8 masks and loadable
    straight into the kernels without an RLE decode."""
    np.savez_compressed(path, words=masklet.numpy_u32(), H=masklet.H, W=masklet.W)


def sidecar_path(json_path: str) -> str:
    """`<id:05d>.json` (the per-track file generate_tokens_grid.py:280 writes) -> `<id:05d>.packed.npz` next to it."""
    return os.path.splitext(json_path)[0] + SIDECAR_SUFFIX


def write_sidecars(masklet_dir: str, device=None) -> int:
    """Stage-1 side: next to every per-track JSON of a `sam2_masklets/...` directory write the packed sidecar of its RLE masklet, so
    that stage 2 (AlignDatasetAdapter.get_sam2_masklet) skips the RLE parse.  Returns the number of sidecars written."""
    from . import rle
    n = 0
    for name in sorted(os.listdir(masklet_dir)):
        if not name.endswith(".json"):
            continue
        path = os.path.join(masklet_dir, name)
        with open(path, "r") as f:
            info = json.load(f)
        packed = rle.decode_rle_masklet_packed(info["rle"], device)
        if packed is not None:
            save_packed_masklet(sidecar_path(path), packed)
            n += 1
    return n


def load_packed_masklet(path: str, device=None) -> P.PackedMasks:
    z = np.load(path)
    return P.PackedMasks.from_numpy_u32(z["words"], int(z["H"]), int(z["W"]), device)


class AlignDatasetAdapter:
    """The J&F half of the reference's `AlignDataset` (dataloader.py:241-369) — same method names, same argument lists, same directory
    walk and selection rules — with every masklet decoded, merged and kept as BIT-PACKED planes on the device:

        set_video / load_gt_masklet      dataloader.py:241-259   GT objects of the video: RLE -> packed (cached per anno id)
        get_gt_masklet                   dataloader.py:278-303   OR over the expression's GT objects
        get_sam2_masklet                 dataloader.py:305-351   per-track JSON -> RLE -> OR over the tracks with preds > 0
        rle_masklet_decode               dataloader.py:353-369   missing frames -> zero planes

    The reference spends its evaluation time here: `json.load` + pycocotools decode (H*W bytes per frame on the host) + `np.logical_or`,
    then a 4 B/px fp32 upload per masklet (evaluator.py:199-200).  Here the host only parses the RLE strings into run lists (C loop in
    the library, `sola_rle_strings_to_runs`); the runs of ALL selected tracks are filled into the same planes on the device (decode
    and OR-merge are one launch) and the packed result goes straight into the fused J&F kernel (`evaluator.JFSweep`).
    Returns `PackedMasks` (T, H, Wp) where the reference returns uint8 / bool (T, H, W) arrays; `None` exactly where it returns None.
    Only the RLE-backed datasets are covered ("mevis" and anything whose GT lives in `mask_dict`); ref-davis reads PNG palettes
    (dataloader.py:260-274, image I/O: out of scope) and ref-ytbvos raises NotImplementedError like the reference."""

    def __init__(self, data_name: str, data_type: str, track_root: str, sam2_output_dirs: Sequence[str], meta: dict,
                 mask_dict: Optional[dict] = None, device=None, use_sidecars: bool = True):
        self.data_name, self.data_type = data_name, data_type
        self.track_root = track_root
        self.sam2_output_dirs = list(sam2_output_dirs)
        self.meta = meta
        self.mask_dict = mask_dict or {}
        self.device = P._dev(device)
        self.video_id = None
        self.cached_gt_masklet: Dict[str, P.PackedMasks] = {}
        self.use_sidecars = use_sidecars        # prefer `<id>.packed.npz` next to a track's JSON (write_sidecars) over its RLE strings

    @classmethod
    def from_dataset(cls, ds, device=None) -> "AlignDatasetAdapter":
        """Wrap a reference `AlignDataset` instance (reads its data_name / data_type / track_root / sam2_output_dirs / meta / mask_dict)."""
        return cls(ds.data_name, ds.data_type, ds.track_root, ds.sam2_output_dirs, ds.meta, getattr(ds, "mask_dict", None), device)

    # -- dataloader.py:241-249 ---------------------------------------------------------------------------------------
    def set_video(self, video_id):
        if self.video_id is None or self.video_id != video_id:
            self.video_id = video_id
            self.load_gt_masklet(video_id)
        else:
            raise NotImplementedError

    # -- dataloader.py:251-276 ---------------------------------------------------------------------------------------
    def load_gt_masklet(self, video_id):
        self.cached_gt_masklet = {}
        if self.data_name == "mevis":
            for _, expression_meta in self.meta["videos"][video_id]["expressions"].items():
                for gt_anno_id in expression_meta["anno_id"]:
                    gt_anno_id = str(gt_anno_id)
                    if gt_anno_id not in self.cached_gt_masklet:
                        self.cached_gt_masklet[gt_anno_id] = self.rle_masklet_decode(self.mask_dict[gt_anno_id])
        elif self.data_name == "ref-davis":
            raise NotImplementedError("ref-davis ground truth is PNG palettes (dataloader.py:260-274): image I/O is outside the hot path")
        else:
            raise ValueError(f"Invalid data_name: {self.data_name}")

    # -- dataloader.py:278-303 ---------------------------------------------------------------------------------------
    def get_gt_masklet(self, video_id, expression_id) -> Optional[P.PackedMasks]:
        assert self.video_id == video_id, f"video_id is not set: {self.video_id} != {video_id}"
        if self.data_name == "mevis" or self.data_name == "ref-davis":
            expression_meta = self.meta["videos"][video_id]["expressions"][expression_id]
            gt_anno_ids = expression_meta["obj_id"] if self.data_name == "ref-davis" else expression_meta["anno_id"]
            planes = []
            for gt_anno_id in gt_anno_ids:
                gt_anno_id = str(gt_anno_id)
                if gt_anno_id in self.cached_gt_masklet:
                    planes.append(self.cached_gt_masklet[gt_anno_id])
                else:
                    planes.append(self.rle_masklet_decode(self.mask_dict[gt_anno_id]))
            if not planes:
                return None
            if len(planes) == 1:
                return planes[0]
            return P.or_merge(P.PackedMasks(torch.stack([p.words for p in planes]), planes[0].H, planes[0].W))
        elif self.data_name == "ref-ytbvos":
            raise NotImplementedError
        raise ValueError(f"Invalid data_name: {self.data_name}")

    # -- dataloader.py:305-351 ---------------------------------------------------------------------------------------
    def get_sam2_masklet(self, video_id: str, expression_id: str, preds, root_types: list, prompt_types: list,
                         sam2_anno_ids: list) -> Optional[P.PackedMasks]:
        from . import rle
        preds = np.asarray(preds.cpu() if isinstance(preds, torch.Tensor) else preds)
        selected, selected_packed, zeros_shape, have_merged = [], [], None, False
        sam2_anno_idx = 0
        for sam2_output_dir in self.sam2_output_dirs:
            sam2_output_dir = os.path.join(self.track_root, sam2_output_dir)
            if "gdino" in sam2_output_dir:
                sam2_masklet_dir = os.path.join(sam2_output_dir, self.data_name, self.data_type, "sam2_masklets", video_id, expression_id)
            else:
                sam2_masklet_dir = os.path.join(sam2_output_dir, self.data_name, self.data_type, "sam2_masklets", video_id)
            for sam2_masklet_path in sorted(n for n in os.listdir(sam2_masklet_dir) if not n.endswith(SIDECAR_SUFFIX)):
                if preds[sam2_anno_idx] < 1 and have_merged:                       # :323-325 — unselected tracks are not even opened
                    sam2_anno_idx += 1
                    continue
                with open(os.path.join(sam2_masklet_dir, sam2_masklet_path), "r") as f:
                    info = json.load(f)
                root_type, prompt_type, sam2_anno_id = root_types[sam2_anno_idx], prompt_types[sam2_anno_idx], sam2_anno_ids[sam2_anno_idx]
                assert root_type == os.path.basename(sam2_output_dir), f"Invalid root_type: {root_type} != {os.path.basename(sam2_output_dir)}"
                assert prompt_type == info["prompt_type"], f"Invalid prompt_type: {prompt_type} != {info['prompt_type']}"
                assert sam2_anno_id == info["anno_id"], f"Invalid sam2_anno_id: {sam2_anno_id} != {info['anno_id']}"
                if preds[sam2_anno_idx] > 0:                                        # :339-344
                    side = sidecar_path(os.path.join(sam2_masklet_dir, sam2_masklet_path))
                    if self.use_sidecars and os.path.isfile(side):                  # packed planes written by stage 1: no RLE parse at all
                        selected_packed.append(load_packed_masklet(side, self.device))
                    else:
                        selected.append(info["rle"])
                elif not have_merged:                                               # :345-349 — zeros of the first track's shape
                    h, w = info["rle"][0]["size"]
                    zeros_shape = (len(info["rle"]), int(h), int(w))
                have_merged = True
                sam2_anno_idx += 1
        if selected or selected_packed:
            planes = list(selected_packed)
            if selected:
                planes.append(rle.decode_rle_masklets_merged(selected, self.device))   # decode + OR-merge of every selected RLE track: one launch
            if len(planes) == 1:
                return planes[0]
            return P.or_merge(P.PackedMasks(torch.stack([p.words for p in planes]), planes[0].H, planes[0].W))
        if zeros_shape is not None:
            t, h, w = zeros_shape
            return P.PackedMasks(torch.zeros((t, h, P.words_per_row(w)), dtype=torch.int32, device=self.device), h, w)
        return None

    # -- dataloader.py:353-369 ---------------------------------------------------------------------------------------
    def rle_masklet_decode(self, rle_masklet) -> Optional[P.PackedMasks]:
        from . import rle
        return rle.decode_rle_masklet_packed(rle_masklet, self.device)

    def get_frames(self, video_id):
        return self.meta["videos"][video_id]["frames"]
