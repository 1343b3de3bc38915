"""sola_b200 — B200 (sm_100a) kernels for the masklet-scoring hot path of cvlab-kaist/SOLA.

Host side in Python (as the reference is Python): thin mirrors of the reference's own function signatures over a
C-ABI CUDA library (`include/sola_maskpath.h`, `sola_b200/csrc/`).  No CPU fallback: importing is cheap, but any
computation without the built library / a CUDA device raises `SolaError`."""
from ._lib import SolaError, load as load_library, launch_count  # noqa: F401
from . import packed, seg_utils, utils, prompt_generator, evaluator, dedup, dataloader_ops, metric, sharding, rle, api  # noqa: F401
from .packed import (PackedMasks, binarize_pack_stability, binarize_pack_resize, pack_masks, unpack_masks, frame_counts,  # noqa: F401
                     frame_counts_packed, pairwise_inter_matrix, gathered_inter, resize_bilinear_bin,
                     resize_nearest, or_merge, boundary_counts, jf_boundary_counts, JFSweepPlan)

__version__ = "0.2.1"
from .api import pairwise_iou_matrix, gathered_iou, greedy_filter, jf_batch  # noqa: F401,E402
