"""TEST INFRASTRUCTURE ONLY — generates tests/golden/dataloader_golden.npz: a small on-disk SAM2-track tree in the reference's layout
(`<track_root>/<sam2_output_dir>/<data_name>/<data_type>/sam2_masklets/<video>[/<expression>]/<id:05d>.json`, dataloader.py:316-321) and
what the UNMODIFIED reference `AlignDataset.get_sam2_masklet` / `get_gt_masklet` (dataloader.py:278-351, imported through
oracle/ref_shim.py) return on it for a set of selection vectors.  pycocotools is not installed, so `mask_utils.decode` is the numpy
restatement of cocoapi's rleDecode / rleFrString (oracle/rle_oracle.py) — the codec itself stays unpinned; what this golden pins is the
directory walk, the selection / zero-fill / OR-merge rules and the missing-frame rule.

    python -m oracle.gen_golden_dataloader
"""
from __future__ import annotations

import json
import os
import tempfile
import types

import numpy as np

from . import ref_shim as R
from . import rle_oracle as RO
from .maskpath_oracle import pack_bits

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataloader_golden.npz")
DATA_NAME, DATA_TYPE = "mevis", "valid_u"
SAM2_DIRS = ["grid_tracks", "gdino_tracks"]


def _masklet(rng, T, H, W, fill):
    yy, xx = np.mgrid[0:H, 0:W]
    cy, cx, r = rng.uniform(0.2, 0.8) * H, rng.uniform(0.2, 0.8) * W, fill * min(H, W)
    out = []
    for t in range(T):
        m = ((yy - cy - t) ** 2 + (xx - cx + 2 * t) ** 2 < r * r) ^ (rng.random((H, W)) > 0.97)
        out.append(m.astype(np.uint8))
    return np.stack(out)


def build_spec(seed=20251017):
    rng = np.random.default_rng(seed)
    videos = {"v0": (4, 40, 70), "v1": (3, 33, 50)}
    meta = {"videos": {}}
    mask_dict, files, cases = {}, {}, []
    anno = 0
    for vid, (T, H, W) in videos.items():
        exprs = {}
        for e in range(2):
            ids = []
            for _ in range(1 + e):                                     # expression 0: one GT object, expression 1: two
                rl = RO.encode_masklet(_masklet(rng, T, H, W, 0.3))
                if anno == 1:
                    rl[1] = None                                        # a missing frame (dataloader.py:364-368)
                mask_dict[str(anno)] = rl
                ids.append(anno)
                anno += 1
            exprs[str(e)] = {"exp": f"expression {e} of {vid}", "anno_id": ids}
        meta["videos"][vid] = {"expressions": exprs, "frames": [f"{t:05d}" for t in range(T)]}
        for k in range(3):                                              # grid tracks: per video
            info = {"anno_id": k, "rle": RO.encode_masklet(_masklet(rng, T, H, W, 0.2 + 0.05 * k)), "prompt_type": "SAM2 AMG MASK"}
            files[f"grid_tracks/{DATA_NAME}/{DATA_TYPE}/sam2_masklets/{vid}/{k:05d}.json"] = json.dumps(info)
        for e in exprs:                                                 # gdino tracks: per (video, expression)
            for k in range(2):
                info = {"anno_id": k, "rle": RO.encode_masklet(_masklet(rng, T, H, W, 0.25)), "prompt_type": "GDINO BOX"}
                files[f"gdino_tracks/{DATA_NAME}/{DATA_TYPE}/sam2_masklets/{vid}/{e}/{k:05d}.json"] = json.dumps(info)
            for preds in ([0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [0, 1, 0, 0, 1], [0, 0, 0, 1, 0], [1, 0, 0, 0, 0]):
                cases.append({"video_id": vid, "expression_id": e, "preds": preds,
                              "root_types": ["grid_tracks"] * 3 + ["gdino_tracks"] * 2,
                              "prompt_types": ["SAM2 AMG MASK"] * 3 + ["GDINO BOX"] * 2, "sam2_anno_ids": [0, 1, 2, 0, 1]})
    return meta, mask_dict, files, cases


def write_tree(root, files):
    for rel, text in files.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)


def reference_dataset(root, meta, mask_dict):
    dl = R.dataloader_module()
    dl.mask_utils = types.SimpleNamespace(decode=RO.decode)            # stands in for pycocotools.mask (absent in this image)
    ds = object.__new__(dl.AlignDataset)                               # the constructor loads RoBERTa tokens etc.: not on this path
    ds.data_name, ds.data_type, ds.track_root, ds.sam2_output_dirs = DATA_NAME, DATA_TYPE, root, list(SAM2_DIRS)
    ds.meta, ds.mask_dict, ds.video_id = meta, mask_dict, None
    return ds


def run_reference(meta, mask_dict, files, cases):
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write_tree(root, files)
        ds = reference_dataset(root, meta, json.loads(json.dumps(mask_dict)))
        for k, c in enumerate(cases):
            if ds.video_id != c["video_id"]:
                ds.set_video(c["video_id"])
            gt = ds.get_gt_masklet(c["video_id"], c["expression_id"])
            pred = ds.get_sam2_masklet(c["video_id"], c["expression_id"], np.asarray(c["preds"]), c["root_types"], c["prompt_types"], c["sam2_anno_ids"])
            out[f"gt_{k}"] = pack_bits(np.asarray(gt).astype(np.uint8))
            out[f"pred_{k}"] = pack_bits(np.asarray(pred).astype(np.uint8))
            out[f"shape_{k}"] = np.asarray(np.asarray(pred).shape, dtype=np.int64)
    return out


def main():
    meta, mask_dict, files, cases = build_spec()
    arrays = run_reference(meta, mask_dict, files, cases)
    blob = lambda o: np.frombuffer(json.dumps(o).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, meta_json=blob(meta), mask_dict_json=blob(mask_dict), files_json=blob(files), cases_json=blob(cases), **arrays)
    print(f"wrote {OUT}: {len(cases)} cases, {len(files)} track files, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
