"""STREAM-style host memory bandwidth of the GPU box (copy and read-sum over buffers far larger than the caches, all cores via torch's
CPU thread pool) next to the aggregate pinned H2D rate N ranks reach at once — the evidence behind bench.py's multi-GPU `e2e` numbers."""
import json, os, sys, time
import torch
n = int(os.environ.get("HB_GB", "8")) * (1 << 30) // 4
torch.set_num_threads(os.cpu_count() or 1)
a = torch.ones(n, dtype=torch.float32)
b = torch.empty_like(a)
out = {"cores": os.cpu_count(), "threads": torch.get_num_threads(), "buffer_GB": n * 4 / 1e9}
for name, fn, bytes_moved in (("copy (read + write)", lambda: b.copy_(a), 2 * n * 4), ("read-sum", lambda: a.sum(), n * 4), ("fill (write)", lambda: b.fill_(2.0), n * 4)):
    fn()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    out[name + " GB/s"] = round(bytes_moved / min(ts) / 1e9, 1)
if torch.cuda.is_available():
    G = torch.cuda.device_count()
    size = 1 << 30
    hosts = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(G)]
    devs = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
    streams = [torch.cuda.Stream(device=g) for g in range(G)]
    def h2d_all(k):
        for g in range(k):
            with torch.cuda.stream(streams[g]):
                for _ in range(4):
                    devs[g].copy_(hosts[g], non_blocking=True)
        for g in range(k):
            streams[g].synchronize()
    for k in sorted({1, 2, 4, G} & set(range(1, G + 1))):
        h2d_all(k)
        t0 = time.perf_counter(); h2d_all(k); dt = time.perf_counter() - t0
        out[f"pinned H2D, {k} GPU(s) at once, aggregate GB/s"] = round(k * 4 * size / dt / 1e9, 1)
print(json.dumps(out))
