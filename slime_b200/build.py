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


def _compile_one(src: Path, headers: list[Path], verbose: bool, variant: str) -> Path:
    obj_dir = BUILD_DIR / variant
    obj = obj_dir / (src.stem + ".o")
    if _needs_rebuild(obj, [src] + headers):
        extra = ["-DSLIME_FP16"] if variant == "fp16" else []
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = obj_dir / (src.stem + ".log")
        log.write_text(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name} ({variant}):\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[slime_b200.build] compiled {src.name} ({variant})")
    return obj


# element type of the build -> library.  Same sources; -DSLIME_FP16 switches the 16-bit element type (common.cuh).
VARIANTS = {"bf16": LIB_PATH, "fp16": PKG_DIR / "libslime_b200_fp16.so"}


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libslime_b200.so (+ the fp16 build) in-tree."""
    srcs = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((REPO / "include").glob("*.h"))
    for variant in VARIANTS:
        (BUILD_DIR / variant).mkdir(parents=True, exist_ok=True)
        if force:
            for o in (BUILD_DIR / variant).glob("*.o"):
                o.unlink()
    jobs = [(s, v) for v in VARIANTS for s in srcs]
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda sv: _compile_one(sv[0], headers, verbose, sv[1]), jobs))
    for variant, lib_path in VARIANTS.items():
        vobjs = [o for o, (_, v) in zip(objs, jobs) if v == variant]
        if _needs_rebuild(lib_path, vobjs):
            # -Bsymbolic: both libraries export the same names; each must bind its internal calls to itself
            cmd = [_nvcc(), "-shared", "-o", str(lib_path), *map(str, vobjs),
                   "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic", "-ldl"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(f"[slime_b200.build] linked {lib_path}")
    return LIB_PATH


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB_PATH)
