"""Build the CUDA library in-tree: lightmotif_b200/liblightmotif_b200.so.

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but
travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "liblightmotif_b200.so"
SOURCES = ["capi.cu", "comm.cu", "score.cu", "prune.cu", "seqops.cu", "reduce.cu", "upload.cu", "hostpack.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# no --use_fast_math: f32 scores must stay IEEE (no ftz, no re-association)
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "lightmotif_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: Path | None = None) -> Path:
    """Compile the library.  `extra_flags` / `out` build experiment variants
    (tools/kbench.py) without touching the product .so."""
    lib = out or LIB
    if out is None and not force and not stale():
        return LIB
    nvcc = nvcc_path()
    objdir = PKG / "_build" / (lib.stem if out else "")
    objdir.mkdir(parents=True, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        (objdir / (src + ".log")).write_text(out)
        objs.append(str(obj))
    cmd = [nvcc, *ARCH, "-shared", "-o", str(lib), *objs]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
