"""In-tree build of libslime_b200.so (hand-written sm_100a CUDA behind a C-ABI).

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so travels
with the repo snapshot to the B200 box.  No torch, no pybind: the library is plain C-ABI and is
loaded with ctypes (slime_b200/_lib.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
REPO = PKG_DIR.parent
BUILD_DIR = REPO / "build" / "slime_b200"
LIB_PATH = PKG_DIR / "libslime_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",  # host double math of preprocess.cu must round like Pillow's C code
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
    "-I", str(REPO / "include"),
    "-I", str(CSRC),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: slime_b200 has no non-CUDA fallback")
    return nvcc


def _needs_rebuild(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _compile_one(src: Path, headers: list[Path], verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    if _needs_rebuild(obj, [src] + headers):
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = BUILD_DIR / (src.stem + ".log")
        log.write_text(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[slime_b200.build] compiled {src.name}")
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libslime_b200.so in-tree."""
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((REPO / "include").glob("*.h"))
    if force:
        for o in BUILD_DIR.glob("*.o"):
            o.unlink()
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, headers, verbose), srcs))
    if _needs_rebuild(LIB_PATH, objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[slime_b200.build] linked {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB_PATH)
