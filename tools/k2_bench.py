"""K2 N x N kernel alone on (a) the bench's object-like resized masklets and (b) dense random planes (no all-zero quad to skip)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
from sola_b200 import synth


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


logits, _ = synth.dedup_candidates(64, 80, 720, 1280, seed=1236, device="cuda", bin_size=4)
_, _, resized = S.binarize_pack_resize(logits)
del logits
ref = S.pairwise_inter_matrix(resized)
words = resized.words[0].numel()
pairs = 64 * 63 // 2
out = {}
ms = timed(lambda: S.pairwise_inter_matrix(resized))
out["object_like_64x80x540x960"] = {"ms": ms, "pair_words_per_s": pairs * words / ms * 1e3, "checksum": int(ref.sum().item())}
dense = S.PackedMasks(torch.randint(-2**31, 2**31 - 1, resized.words.shape, dtype=torch.int32, device="cuda"), resized.H, resized.W)
d = S.pairwise_inter_matrix(dense)
ms = timed(lambda: S.pairwise_inter_matrix(dense))
out["dense_random"] = {"ms": ms, "pair_words_per_s": pairs * words / ms * 1e3, "checksum": int(d.sum().item())}
print(json.dumps(out))
