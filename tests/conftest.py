import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "maskpath_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        self.z = np.load(GOLDEN)
        self.meta = json.loads(bytes(self.z["meta_json"]).decode())
        self.greedy = json.loads(bytes(self.z["greedy_json"]).decode())

    def __getitem__(self, k):
        return self.z[k]

    def masks(self, key, W):
        """unpack a stored bit-packed mask array to uint8 (..., H, W)"""
        from oracle.maskpath_oracle import unpack_bits
        return unpack_bits(self.z[key], W)


@pytest.fixture(scope="session")
def golden():
    return Golden()


def greedy_table(golden):
    """The synthetic masklet table of the greedy golden cases: (masklets uint8 (n,T,H,W), frame_idx, stability)."""
    n, T, H, W = golden.meta["greedy_shape"]
    return golden.masks("greedy_masklets", W), golden["greedy_frame_idx"], golden["greedy_stability"]


def greedy_prompts(golden):
    m, frame_idx, stab = greedy_table(golden)
    return [{"prompt_id": k, "frame_idx": int(frame_idx[k]), "segmentation": np.ascontiguousarray(m[k, frame_idx[k]]),
             "expression_id": "0" if k % 4 else "1", "stability_score": float(stab[k])} for k in range(m.shape[0])]


GREEDY_CASES = {
    "grid_default": ("grid", dict(n_max_tracks=64, batch_size=4), 8),
    "grid_cap5": ("grid", dict(n_max_tracks=5, batch_size=4), 8),
    "grid_bs2": ("grid", dict(n_max_tracks=64, batch_size=2, miou_thresh=0.5), 8),
    "grid_long_video": ("grid", dict(n_max_tracks=64, batch_size=4), 250),
    "gdino_default": ("gdino", dict(n_max_tracks=16, batch_size=4, stability_score_thresh=0.85), 8),
    "gdino_cap6": ("gdino", dict(n_max_tracks=6, batch_size=4, stability_score_thresh=0.85), 8),
    "gdino_loose": ("gdino", dict(n_max_tracks=16, batch_size=4, stability_score_thresh=0.5, miou_thresh=0.45), 8),
    "gdino_loose_cap4": ("gdino", dict(n_max_tracks=4, batch_size=3, stability_score_thresh=0.5, miou_thresh=0.45), 8),
}
