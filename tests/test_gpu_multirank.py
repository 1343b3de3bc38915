"""GPU, 2 ranks over NCCL (skipped on boxes with fewer than 2 GPUs): BASELINE config 5's exchange paths —
`sharding.pairwise_inter_matrix_sharded` (split="words" / "tiles") and `sharding.PeerPlanes` (one-kernel TMA / pull / direct) — plus the J&F
sweep's final all-reduce, each against the single-rank result (tests/multirank_worker.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(n, extra=()):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("MULTIRANK_RESULT ")]
    assert p.returncode == 0 and line, f"rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}"
    return json.loads(line[-1][len("MULTIRANK_RESULT "):])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_exchange_paths_match_single_rank():
    res = _run(2)
    assert res["world"] == 2
    for name in ("nccl_words", "nccl_tiles", "peer_pull", "peer_direct", "peer_tma", "peer_auto", "jf_allreduce_int_totals"):
        assert res[name] is True, (name, res[name])
