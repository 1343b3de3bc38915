"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference functions from /root/reference.

This shim exists so that (a) `oracle/gen_golden.py` can produce golden vectors from the real
reference implementation and (b) the `-m "not gpu"` tests can re-check the restated oracle
(`oracle/maskpath_oracle.py`) against the reference whenever `/root/reference` is mounted
(builder container only; the tree does not exist on the GPU box, so everything here is gated
by `available()`).

Nothing under `sola_b200/` may import this module.

How the import works: the reference modules import third-party packages that are not installed
(`pycocotools`, `sam2`, `groundingdino`).  None of the functions on the masklet-scoring path touch
them, so they are replaced by empty stand-in modules in `sys.modules` for the duration of the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SOLA_REFERENCE_ROOT", "/root/reference")

_STUBS = (
    "pycocotools", "pycocotools.mask",
    "groundingdino", "groundingdino.util", "groundingdino.util.slconfig", "groundingdino.util.utils",
    "groundingdino.models", "groundingdino.datasets", "groundingdino.datasets.transforms",
    "sam2", "sam2.build_sam", "sam2.sam2_image_predictor",
)

_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "track_generation", "seg_utils.py"))


class _Anything(types.ModuleType):
    """Module stand-in: any attribute resolves to a dummy callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def _install_stubs():
    installed = []
    for name in _STUBS:
        if name not in sys.modules:
            mod = _Anything(name)
            mod.__path__ = []          # behave like a package so dotted imports resolve
            sys.modules[name] = mod
            installed.append(name)
            parent, _, leaf = name.rpartition(".")
            if parent and parent in sys.modules:
                setattr(sys.modules[parent], leaf, mod)
    return installed


def _load():
    if _cache:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    installed = _install_stubs()
    tg = os.path.join(REFERENCE_ROOT, "track_generation")
    added = [p for p in (tg, REFERENCE_ROOT) if p not in sys.path]
    saved_mods = {k: sys.modules.get(k) for k in ("utils", "seg_utils", "prompt_generator", "evaluator",
                                                  "dataloader", "tools", "tools.metric", "tools.loss")}
    for p in added:
        sys.path.insert(0, p)
    try:
        for k in saved_mods:
            sys.modules.pop(k, None)
        _cache["seg_utils"] = importlib.import_module("seg_utils")          # track_generation/seg_utils.py
        _cache["utils"] = importlib.import_module("utils")                  # track_generation/utils.py
        _cache["prompt_generator"] = importlib.import_module("prompt_generator")
        _cache["evaluator"] = importlib.import_module("evaluator")
        _cache["dataloader"] = importlib.import_module("dataloader")
        _cache["metric"] = importlib.import_module("tools.metric")
    finally:
        for p in added:
            if p in sys.path:
                sys.path.remove(p)
        # do not leave the reference's generic module names ("utils", "tools", ...) importable
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
        for name in installed:
            sys.modules.pop(name, None)
    return _cache


# ---- unbound reference callables (self=None where the method uses no attributes) -------------------

def compute_mask_iou(a, b):                       # track_generation/seg_utils.py:129
    return _load()["seg_utils"].compute_mask_iou(a, b)


def compute_masklet_iou(a, b, device="cpu"):      # track_generation/seg_utils.py:110
    return _load()["seg_utils"].compute_masklet_iou(a, b, device)


def reshape_masklet(m, target_shape=None):        # track_generation/seg_utils.py:145
    return _load()["seg_utils"].reshape_masklet(m, target_shape)


def compute_mask_iou_torch(a, b):                 # track_generation/utils.py:65
    return _load()["utils"].compute_mask_iou_torch(a, b)


def compute_mask_metrics(p, g, reduction="mean"):  # track_generation/utils.py:132
    return _load()["utils"].compute_mask_metrics(p, g, reduction)


def compute_P(parts, full):                       # track_generation/utils.py:178
    return _load()["utils"].compute_P(parts, full)


def get_stability_score(logit, mask_threshold=0.0, threshold_offset=1.0):   # prompt_generator.py:169
    return _load()["prompt_generator"].PromptGenerator.get_stability_score(None, logit, mask_threshold, threshold_offset)


def compute_J(pred, gt):                          # evaluator.py:227
    return _load()["evaluator"].Evaluator.compute_J(None, pred, gt)


def compute_F(pred, gt):                          # evaluator.py:239
    return _load()["evaluator"].Evaluator.compute_F(None, pred, gt)


def recall_per_track(*a):                         # tools/metric.py:2
    return _load()["metric"].recall_per_track(*a)


def recall_per_exp(*a):                           # tools/metric.py:34
    return _load()["metric"].recall_per_exp(*a)


def evaluator_module():
    return _load()["evaluator"]


def dataloader_module():
    return _load()["dataloader"]
