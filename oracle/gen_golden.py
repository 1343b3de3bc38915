"""TEST INFRASTRUCTURE ONLY — generates tests/golden/maskpath_golden.npz by running the UNMODIFIED reference
functions (imported through oracle/ref_shim.py) on small seeded inputs.  Run in the builder container, where
/root/reference is mounted:

    python -m oracle.gen_golden

Inputs are stored next to the outputs so the fixtures do not depend on generator code staying unchanged.
Torch CPU (the container has no GPU) — `reshape_masklet` vectors therefore pin ATen's *CPU* bilinear kernel; the
CUDA kernel is pinned on the GPU box against torch-CUDA itself (tests/test_gpu_resize.py).
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import greedy_oracle, ref_shim as R
from .maskpath_oracle import pack_bits

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maskpath_golden.npz")


def _smooth(n, H, W, seed, cell=8):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn((n, 1, H // cell + 2, W // cell + 2), generator=g)
    up = torch.nn.functional.interpolate(z, size=(H, W), mode="bicubic", align_corners=False)[:, 0]
    return (6.0 * (up - 0.3) + 0.5 * torch.randn((n, H, W), generator=g)).float()


def _adversarial(H, W):
    vals = torch.tensor([0.0, -0.0, 1.0, -1.0, float("nan"), float("inf"), -float("inf"), 1.0000001, 0.99999994,
                         -0.99999994, -1.0000001, 1e-45, -1e-45, 3.0, -3.0, 0.5])
    idx = torch.arange(H * W) % vals.numel()
    return torch.stack([vals[idx].reshape(H, W), torch.full((H, W), -5.0), torch.full((H, W), 5.0),
                        vals[(idx * 7 + 3) % vals.numel()].reshape(H, W)])


def _blob_masks(n, H, W, seed, fill=0.35):
    x = _smooth(n, H, W, seed)
    thr = torch.quantile(x.flatten(), 1 - fill)
    return (x > thr).float()


def main():
    G = {}
    meta = {"torch": torch.__version__, "numpy": np.__version__, "device": "cpu"}

    # ---- B1 + S1: binarise and stability score -------------------------------------------------------------
    for tag, (H, W) in {"a": (48, 96), "b": (37, 70)}.items():       # W % 32 == 0 and ragged W
        logits = torch.cat([_smooth(4, H, W, 11 + len(tag)), _adversarial(H, W)])
        G[f"stab_{tag}_logits"] = logits.numpy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            G[f"stab_{tag}_score"] = np.array([R.get_stability_score(l.numpy()) for l in logits], dtype=np.float64)
            G[f"stab_{tag}_score_t05_o025"] = np.array([R.get_stability_score(l.numpy(), 0.5, 0.25) for l in logits], dtype=np.float64)
        G[f"stab_{tag}_binarized_packed"] = pack_bits((logits > 0.0).float().numpy())     # generate_tokens_grid.py:219

    # ---- I1 / I2 / I3 ------------------------------------------------------------------------------------------
    H, W = 60, 94
    g = torch.Generator().manual_seed(5)
    A = torch.cat([(torch.rand((3, H, W), generator=g) > 0.5).float(), _blob_masks(3, H, W, 21),
                   torch.zeros(1, H, W), torch.zeros(1, H, W), torch.ones(1, H, W)])
    B = torch.cat([(torch.rand((3, H, W), generator=g) > 0.7).float(), _blob_masks(3, H, W, 22),
                   torch.zeros(1, H, W), torch.ones(1, H, W), torch.ones(1, H, W)])
    G["iou_A"], G["iou_B"] = pack_bits(A.numpy()), pack_bits(B.numpy())
    meta["iou_shape"] = [H, W]
    G["iou_mask_iou"] = np.array([R.compute_mask_iou(a, b) for a, b in zip(A, B)])
    vals = []
    for a, b in zip(A, B):
        try:
            vals.append(R.compute_mask_iou_torch(a, b))
        except ZeroDivisionError:
            vals.append(np.nan)                      # nan marks "raises ZeroDivisionError"
    G["iou_mask_iou_torch"] = np.array(vals)
    G["iou_masklet_iou"] = np.array([R.compute_masklet_iou(A[:6], B[:6], "cpu"), R.compute_masklet_iou(A[6:7], B[6:7], "cpu")])

    # ---- M1 ----------------------------------------------------------------------------------------------------
    pred = torch.cat([_blob_masks(4, H, W, 31), torch.zeros(2, H, W), _blob_masks(1, H, W, 33)])
    gt = torch.cat([_blob_masks(4, H, W, 32), torch.zeros(1, H, W), _blob_masks(1, H, W, 34), torch.zeros(1, H, W)])
    G["mm_pred"], G["mm_gt"] = pack_bits(pred.numpy()), pack_bits(gt.numpy())
    p, r, i = R.compute_mask_metrics(pred, gt, "none")
    G["mm_none"] = torch.stack([p, r, i]).numpy()
    p, r, i = R.compute_mask_metrics(pred, gt)
    G["mm_mean"] = torch.stack([p, r, i]).numpy()

    # ---- J1 / F1 -----------------------------------------------------------------------------------------------
    G["jf_J"] = np.array([R.compute_J(pred, gt), R.compute_J(torch.zeros_like(gt), gt), R.compute_J(gt, gt)])
    G["jf_F"] = np.array([R.compute_F(pred, gt), R.compute_F(torch.zeros_like(gt), gt), R.compute_F(gt, gt)])

    # ---- R1 / R2 -----------------------------------------------------------------------------------------------
    land = _blob_masks(3, 72, 128, 41)
    port = _blob_masks(2, 128, 72, 42)
    G["rs_land_in"], G["rs_port_in"] = pack_bits(land.numpy()), pack_bits(port.numpy())
    G["rs_land_out"] = pack_bits(R.reshape_masklet(land).numpy())                    # -> 540 x 960
    G["rs_port_out"] = pack_bits(R.reshape_masklet(port).numpy())                    # -> 960 x 540
    G["rs_land_out_45x80"] = pack_bits(R.reshape_masklet(land, (45, 80)).numpy())
    sq = _blob_masks(1, 64, 64, 43)
    G["rs_sq_in"] = pack_bits(sq.numpy())
    G["rs_sq_out"] = pack_bits(R.reshape_masklet(sq).numpy())                        # square -> portrait 960 x 540
    prompt = land[0].numpy().astype(np.uint8)
    near = torch.nn.functional.interpolate(torch.from_numpy(prompt).float()[None, None], size=(540, 960), mode="nearest")[0, 0]
    G["rs_nearest_out"] = pack_bits(near.numpy())                                    # generate_tokens_grid.py:271-272

    # ---- P1 ----------------------------------------------------------------------------------------------------
    parts, full = _blob_masks(5, 40, 64, 51, fill=0.2), _blob_masks(1, 40, 64, 52, fill=0.5)[0]
    G["P_parts"], G["P_full"] = pack_bits(parts.numpy()), pack_bits(full.numpy())
    G["P_out"] = R.compute_P(parts, full).numpy()

    # ---- X1 ----------------------------------------------------------------------------------------------------
    gt_ids = [3, 5, 9]
    preds = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0, 0.0])
    labels = torch.tensor([1, 1, 0, 1, 1, 0])
    corr = [3, 3, 5, 5, 7, 9]
    G["x1_recall_per_track"] = np.array(R.recall_per_track(gt_ids, preds, labels, corr))
    G["x1_recall_per_exp"] = np.array([R.recall_per_exp(gt_ids, preds, labels, corr)])

    # ---- G1 / G2: restated loops driven by the REFERENCE's compute_mask_iou / reshape_masklet -----------------
    class RefImpl:
        compute_mask_iou = staticmethod(R.compute_mask_iou)
        reshape_masklet = staticmethod(R.reshape_masklet)

    n, T, H, W = 14, 8, 72, 128
    g = torch.Generator().manual_seed(61)
    base = torch.nn.functional.interpolate(torch.randn((5, 1, H // 16 + 3, W // 16 + 3), generator=g), size=(H + 8, W + 8),
                                           mode="bicubic", align_corners=False)[:, 0]
    cluster = torch.randint(0, 5, (n,), generator=g)
    shift = torch.randint(2, 7, (n, 2), generator=g)
    level = 0.3 + 0.3 * torch.rand((n,), generator=g)
    masklets = torch.stack([
        torch.stack([(base[cluster[i], shift[i, 0]: shift[i, 0] + H, shift[i, 1]: shift[i, 1] + W] - level[i] - 0.02 * t
                      + 0.05 * torch.randn((H, W), generator=g)) > 0 for t in range(T)]) for i in range(n)]).float()
    frame_idx = (torch.randint(0, 2, (n,), generator=g) * 4).numpy()
    frame_idx[3] = 2                                              # one off-bin prompt -> status 3
    areas = np.array([int(masklets[i, frame_idx[i]].sum()) for i in range(n)])
    order = np.argsort(-areas, kind="stable")
    masklets = masklets[torch.as_tensor(order)]
    frame_idx = frame_idx[order]
    stab = np.linspace(0.80, 0.95, n)
    stab[5] = np.nan                                              # NaN score must be kept (gdino :162)
    stab[6] = 0.85                                                # == thresh must be kept
    G["greedy_masklets"] = pack_bits(masklets.numpy())
    G["greedy_frame_idx"] = frame_idx.astype(np.int32)
    G["greedy_stability"] = stab
    meta["greedy_shape"] = [n, T, H, W]

    def make_prompts():
        return [{"prompt_id": k, "frame_idx": int(frame_idx[k]), "segmentation": masklets[k, frame_idx[k]].numpy().astype(np.uint8),
                 "expression_id": "0" if k % 4 else "1", "stability_score": float(stab[k])} for k in range(n)]

    track_fn = lambda frame, batch: {p["prompt_id"]: masklets[p["prompt_id"]] for p in batch}
    greedy_out = {}
    for name, kw in {"grid_default": dict(n_max_tracks=64, batch_size=4), "grid_cap5": dict(n_max_tracks=5, batch_size=4),
                     "grid_bs2": dict(n_max_tracks=64, batch_size=2, miou_thresh=0.5)}.items():
        log = []
        res = greedy_oracle.grid_greedy(make_prompts(), T, track_fn, bin_size=4, impl=RefImpl, log=log, **kw)
        greedy_out[name] = {k: res[k] for k in ("tracked", "filtered", "not_used", "not_tracked", "batches", "n_tracked", "n_filtered")}
        greedy_out[name]["filtered_by"] = {str(k): v for k, v in res["filtered_by"].items()}
        greedy_out[name]["filtered_iou"] = {str(k): v for k, v in res["filtered_iou"].items()}
        greedy_out[name]["iou_log"] = log
    res = greedy_oracle.grid_greedy(make_prompts(), 250, track_fn, bin_size=4, impl=RefImpl, n_max_tracks=64, batch_size=4)
    greedy_out["grid_long_video"] = {k: res[k] for k in ("tracked", "filtered", "batches")}
    for name, kw in {"gdino_default": dict(n_max_tracks=16, batch_size=4, stability_score_thresh=0.85),
                     "gdino_cap6": dict(n_max_tracks=6, batch_size=4, stability_score_thresh=0.85),
                     "gdino_loose": dict(n_max_tracks=16, batch_size=4, stability_score_thresh=0.5, miou_thresh=0.45),
                     "gdino_loose_cap4": dict(n_max_tracks=4, batch_size=3, stability_score_thresh=0.5, miou_thresh=0.45)}.items():
        for eid in ("0", "1"):
            log = []
            res = greedy_oracle.gdino_greedy(make_prompts(), eid, T, track_fn, bin_size=4,
                                             impl=RefImpl, log=log, **kw)
            greedy_out[f"{name}_exp{eid}"] = {k: res[k] for k in ("tracked", "filtered", "batches", "n_tracked", "n_filtered", "n_not_used")}
            greedy_out[f"{name}_exp{eid}"]["filtered_by"] = {str(k): v for k, v in res["filtered_by"].items()}
            greedy_out[f"{name}_exp{eid}"]["iou_log"] = log
    G["greedy_json"] = np.frombuffer(json.dumps(greedy_out).encode(), dtype=np.uint8)

    G["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
