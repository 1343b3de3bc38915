"""Every kernel of libsola_maskpath.so once, on small awkward shapes — run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import dedup, evaluator, rle, synth, utils

g = torch.Generator().manual_seed(0)
for shape in [(3, 40, 64), (2, 37, 70), (4, 21, 100), (2, 5, 33), (2, 96, 160)]:
    x = (torch.randn(shape, generator=g) * 2).cuda()
    for dt in (torch.float32, torch.bfloat16):
        xx = x.to(dt)
        p, c = S.binarize_pack_stability(xx)
        for tgt in ((27, 48), (60, 50)):
            fused = S.binarize_pack_resize(xx, target_shape=tgt, want_area=True)
            assert torch.equal(fused[2].words, S.resize_bilinear_bin(p, tgt).words), "fused K1+R1 != K1 then R1"
    m = (x > 0)
    pk, area = S.pack_masks(m.float(), want_area=True)
    S.pack_masks(m.to(torch.uint8))
    S.unpack_masks(pk, torch.float32); S.unpack_masks(pk, torch.uint8)
    S.frame_counts(m.float(), (x > 0.5).float()); S.frame_counts(m.to(torch.uint8), (x > 0.5).to(torch.uint8))
    r = S.resize_bilinear_bin(pk, (30, 45), want_area=True)
    S.packed.resize_bilinear_bin_f32(x, (30, 45), want_f32=True)
    S.resize_nearest(m.to(torch.uint8), 30, 45); S.resize_nearest(pk, 30, 45)
    S.boundary_counts(pk, S.pack_masks((x > 0.3).float()))
    S.or_merge(pk, select=[1] + [0] * (shape[0] - 1))
    enc = rle.encode_rle_masklet_torch(pk)
    assert np.array_equal(rle.decode_rle_masklet(enc), m.cpu().numpy().astype(np.uint8))
tracks = S.pack_masks((torch.randn((70, 3, 16, 64), generator=g) > 0).cuda().float())
S.pairwise_inter_matrix(tracks)
S.packed.pairwise_inter_matrix_part(tracks, 1, 3)
full = S.pairwise_inter_matrix(tracks)
words = tracks.words[0].numel()
ptrs = torch.tensor([tracks.words[i].data_ptr() for i in range(70)], dtype=torch.int64, device="cuda")
assert torch.equal(sum(S.packed.pairwise_inter_matrix_rows(ptrs, words, k, 2) for k in range(2)), full)       # peer-load K2 entry
acc = torch.zeros_like(full)
for lo, hi in ((0, 32), (32, words)):
    chunk = torch.empty((70, hi - lo), dtype=torch.int32, device="cuda")
    S.packed.pull_rows(ptrs, lo, hi - lo, chunk)
    S.packed.pairwise_inter_accumulate(chunk, acc)                                                              # pull + TMA ring K2
assert torch.equal(acc, full)
odd = S.pack_masks((torch.randn((5, 1, 5, 70), generator=g) > 0).cuda().float())
S.pairwise_inter_matrix(odd)
prompts = S.pack_masks((torch.randn((9, 16, 64), generator=g) > 0).cuda().float())
S.gathered_inter(tracks[:6], prompts, np.arange(9) % 3)
S.frame_counts_packed(tracks[:5], tracks[5:11])
utils.suppress_part_masks((torch.randn((6, 20, 30), generator=g) > 0).float().cuda())
logits, pr = synth.dedup_candidates(8, 8, 72, 128, seed=5, device="cpu", n_clusters=2, jitter=1)
job = dedup.VideoDedupJob([{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"]} for p in pr], 8, mode="grid")
job.enqueue(logits.cuda(), torch.from_numpy(np.stack([p["segmentation"] for p in pr])).cuda())
job.finish()
evaluator.compute_JF(*synth.jf_pair(4, 50, 77, 1, device="cuda"))
pj, gj = synth.jf_pair(4, 50, 77, 1, device="cuda")
for a_, b_ in ((pj, gj), (pj.float(), gj.float()), (S.pack_masks(pj), S.pack_masks(gj))):
    S.packed.jf_accumulators(a_, b_)
# fused J&F kernel: both modes, a multi-band frame, a mixed sweep (two tile classes), merged RLE decode, labels in the job
from sola_b200 import packed as P
for (T_, H_, W_) in ((3, 48, 85), (2, 300, 70), (1, 1080, 1920)):
    a_, b_ = synth.object_pair(T_, H_, W_, 3, device="cuda", speckle=0.01)
    P.jf_boundary_counts(S.pack_masks(a_), S.pack_masks(b_))
    P.jf_boundary_counts(S.pack_masks(a_), S.pack_masks(b_), with_boundary=False)
units = synth.mevis_like_sweep(3, 2, 9, "cuda", t_range=(2, 4), shapes=[(48, 85), (120, 214), (1080, 1920)], pack=S.pack_masks)
P.JFSweepPlan([(p_, g_) for _, _, p_, g_ in units], with_boundary=True).run()
enc2 = rle.encode_rle_masklet_torch(S.pack_masks(synth.blob_masklet(3, 40, 70, 2, device="cuda")))
rle.decode_rle_masklets_merged([enc2, enc2])
job.set_gt_masklets(S.pack_masks(torch.stack([synth.blob_masklet(8, 540, 960, 4 + k, device="cuda") for k in range(2)])))
job.enqueue(logits.cuda(), torch.from_numpy(np.stack([p["segmentation"] for p in pr])).cuda())
assert "labels" in job.finish()
# region-only build of the J&F kernel on a mixed sweep (aligned and word-shifted planes), the job with its optional streams
P.JFSweepPlan([(p_, g_) for _, _, p_, g_ in units], with_boundary=False).run()
flat = torch.zeros(units[0][2].words.numel() + 1, dtype=torch.int32, device="cuda")
flat[1:] = units[0][2].words.reshape(-1)
P.jf_boundary_counts(P.PackedMasks(flat[1:].view(units[0][2].words.shape), units[0][2].H, units[0][2].W), units[0][3], with_boundary=False)
job2 = dedup.VideoDedupJob([{"prompt_id": p["prompt_id"], "frame_idx": p["frame_idx"]} for p in pr], 8, mode="grid",
                           aux_stream=torch.cuda.Stream(), tail_stream=torch.cuda.Stream(priority=-1))
job2.set_gt_masklets(job.gt_planes)
for _ in range(2):
    job2.enqueue(logits.cuda(), torch.from_numpy(np.stack([p["segmentation"] for p in pr])).cuda())
    assert "labels" in job2.finish()
S.frame_counts_packed(tracks[:7], tracks[7:12])            # odd track count: the track tile's clamped slot
torch.cuda.synchronize()
print("sanitize smoke ok")
