"""R2 alone: nearest resize + pack of 64 uint8 prompt masks 720x1280 -> 540x960 (what one bench step runs)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sola_b200 as S
g = torch.Generator(device="cuda"); g.manual_seed(3)
m = (torch.rand((64, 720, 1280), generator=g, device="cuda") > 0.6).to(torch.uint8)
for _ in range(3):
    out = S.resize_nearest(m, 540, 960)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    S.resize_nearest(m, 540, 960)
b.record(); torch.cuda.synchronize()
print(json.dumps({"build": os.environ.get("SOLA_EXTRA_NVCC_FLAGS", "default"), "us": a.elapsed_time(b) / 50 * 1e3, "checksum": int(out.words.long().sum().item())}))
