"""In-tree build of libmrag.so (sm_100a only) with plain nvcc.

`python -m motionrag_b200.build` compiles motionrag_b200/csrc/*.cu into
motionrag_b200/_lib/libmrag.so. nvcc cross-compiles without a GPU, so this runs in the
CPU-only build container; the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_lib"
LIB = OUT_DIR / "libmrag.so"
STAMP = OUT_DIR / "libmrag.stamp"

SOURCES = ["api.cu", "api_cama.cu", "k1_stream.cu", "k2_batch.cu", "k2_batch2.cu", "k3_merge.cu", "k4_gather.cu", "k5_cama.cu"]
HEADERS = ["common.cuh", "kernels.h", "k2_common.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xfatbin", "-compress-all",     # cubins carry -lineinfo tables; compressed they are a third of the size
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: libmrag cannot be built (and there is no fallback path)")
    return exe


def _digest() -> str:
    h = hashlib.sha256()
    root = PKG.parent
    for f in [CSRC / s for s in SOURCES + HEADERS] + [root / "include" / "mrag.h"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile libmrag.so if sources changed; returns the library path."""
    OUT_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = _nvcc()
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
