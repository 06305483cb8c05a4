"""In-tree build of ``libhermes_b200.so`` (sm_100a only) with plain nvcc.

``python -m hermespy_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the resulting shared object lives next to the sources (``hermespy_b200/lib/``), is git-ignored
and travels to the GPU box with the repo snapshot.
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
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = ROOT / "build" / "obj"
LIBNAME = "libhermes_b200.so"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I",
    str(ROOT / "include"),
    "-I",
    str(CSRC),
]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; set NVCC or add /usr/local/cuda/bin to PATH")
    return cand


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _headers():
    return sorted(list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h")))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def library_path() -> Path:
    return LIBDIR / LIBNAME


def build_library(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link the C-ABI shared library."""
    nvcc = nvcc_path()
    LIBDIR.mkdir(parents=True, exist_ok=True)
    OBJDIR.mkdir(parents=True, exist_ok=True)
    hdr_digest = _digest(_headers())
    srcs = _sources()
    jobs = []
    objs = []
    for src in srcs:
        obj = OBJDIR / (src.stem + ".o")
        stamp = OBJDIR / (src.stem + ".sha")
        want = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
            continue
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if ptxas_info:
            cmd += ["-Xptxas", "-v"]
        jobs.append((cmd, stamp, want, src))

    def run(job):
        cmd, stamp, want, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose or ptxas_info:
            sys.stderr.write(r.stderr)
        stamp.write_text(want)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(run, jobs):
                if verbose:
                    print(f"[hermespy_b200.build] compiled {name}")
    lib = library_path()
    if jobs or force or not lib.exists():
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", str(lib), *map(str, objs), "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[hermespy_b200.build] linked {lib}")
    return lib


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(p)
