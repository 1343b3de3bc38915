"""In-tree build of libsola_maskpath.so (nvcc, sm_100a only).  No JIT cache, no torch extension machinery:
the shared object lands next to the sources so it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsola_maskpath.so")
STAMP = os.path.join(LIB_DIR, "libsola_maskpath.stamp")

SOURCES = ["abi.cu", "binarize_pack.cu", "counts.cu", "pair_iou.cu", "resize.cu", "fused_pack_resize.cu", "jf_fused.cu", "rle.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "--shared", "-cudart", "shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _extra_flags():
    """Build-time experiment hook (tools only): extra nvcc flags, e.g. SOLA_EXTRA_NVCC_FLAGS="-DJF_PRE_R=3".  Part of the digest."""
    return os.environ.get("SOLA_EXTRA_NVCC_FLAGS", "").split()


def _source_digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def source_digest() -> str:
    return _source_digest()


def is_current() -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into one shared library.  Returns its path."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # several ranks may import at once (torchrun): one builds under an exclusive lock, into a temp file renamed atomically
    import fcntl
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():          # another process finished the build while we waited
                return LIB_PATH
            srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            # the digest of the sources is compiled in (sola_build_digest()), so the loader can tell a stale library from a current one
            # even when the stamp file did not travel with it
            cmd = [_nvcc(), *NVCC_FLAGS, *_extra_flags(), f'-DSOLA_SOURCE_DIGEST="{_source_digest()}"', "-I", CSRC, "-o", tmp, *srcs]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(res.stderr, file=sys.stderr)
            os.replace(tmp, LIB_PATH)
            with open(STAMP, "w") as f:
                f.write(_source_digest())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
