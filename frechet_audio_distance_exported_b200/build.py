"""Build libfadb200.so in-tree with nvcc for sm_100a (B200).  No JIT cache: the .so travels with the tree.

    python -m frechet_audio_distance_exported_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfadb200.so")
STAMP = os.path.join(HERE, "libfadb200.so.stamp")
SOURCES = ["api.cu", "gemm_tc.cu", "conv1.cu", "pack.cu", "frontend.cu", "stats.cu", "frechet.cu", "resample.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfadb200.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(root, "include", "fadb.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu to an object (in parallel, one nvcc per source) and link libfadb200.so.  Objects go to the
    git-ignored build/ directory at the repo root; the .so stays next to the package so it travels with the tree."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    objdir = os.path.join(os.path.dirname(HERE), "build")
    os.makedirs(objdir, exist_ok=True)
    print("[fadb] building", LIB, flush=True)
    jobs = [(nvcc, os.path.join(CSRC, s), os.path.join(objdir, s.replace(".cu", ".o")), verbose) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile_one, jobs))
    failed = [r for r in results if r[1] != 0]
    for src, rc, out in results:
        if rc != 0 or verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libfadb200.so: " + ", ".join(os.path.basename(r[0]) for r in failed))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"] +
                       [j[2] for j in jobs] + ["-o", LIB], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking libfadb200.so failed")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
