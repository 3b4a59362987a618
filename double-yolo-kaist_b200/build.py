"""In-tree build of libdyk_b200.so (sm_100a only) with plain nvcc — no torch headers, no JIT cache.

    python double-yolo-kaist_b200/build.py [--force] [--verbose]

The shared object lands next to this file so that it travels with the source tree; the Python host
(`_native.py`) loads it with ctypes and fails loudly when it is missing.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libdyk_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libdyk_b200.so cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    files = list(CSRC.glob("*")) + [HERE.parent / "include" / "dyk_b200.h", Path(__file__)]
    return max(f.stat().st_mtime for f in files if f.is_file())


def up_to_date() -> bool:
    return LIB.exists() and LIB.stat().st_mtime >= _deps_mtime()


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> Path:
    """profile=True builds libdyk_b200_prof.so: same sources with -DDYK_CONV_PROFILE (role-cycle counters in the
    tcgen05 conv kernel, see dyk_conv_set_profile); select it at run time with DYK_B200_LIB=<path>."""
    lib = LIB.with_name("libdyk_b200_prof.so") if profile else LIB
    if not force and lib.exists() and lib.stat().st_mtime >= _deps_mtime():
        return lib
    nvcc = _nvcc()
    objdir = OBJ / "prof" if profile else OBJ
    objdir.mkdir(exist_ok=True, parents=True)
    srcs = _sources()
    extra = ["-DDYK_CONV_PROFILE"] if profile else []

    def compile_one(src: Path) -> Path:
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = lib.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile="--profile" in sys.argv)
    print(p)
