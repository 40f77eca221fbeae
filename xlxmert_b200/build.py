"""Builds the CUDA sources under ``csrc/`` into ``lib/libxlxmert_b200.so`` for sm_100a with nvcc.

In-tree on purpose: the built ``.so`` travels to the GPU box with the repo snapshot (it is git-ignored),
and the product path refuses to run without it (``_lib.load`` raises).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libxlxmert_b200.so")
GEMM_TEST = os.path.join(LIBDIR, "gemm_test")     # standalone tcgen05-GEMM probe (csrc/gemm_test.cu), run by tests/test_gemm_probe.py
SOURCES = ["gemm_sm100.cu", "kernels.cu", "attention.cu", "encoder.cu", "heads.cu", "generator.cu", "optim.cu", "kmeans.cu", "inputs.cu", "sampler.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libxlxmert_b200.so")
    return exe


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def source_hash() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "xlxmert_b200.h")]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p) and (n.endswith((".cu", ".cuh", ".h"))):
            h.update(n.encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into one shared library; returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    want = source_hash()
    if (not force and os.path.exists(LIB) and os.path.exists(GEMM_TEST) and os.path.exists(stamp)
            and open(stamp).read().strip() == want):
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    objs = []
    procs = []
    for s in _sources():
        o = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    probe = subprocess.Popen([_nvcc(), *[f for f in flags if f != "-fPIC" and f != "-Xcompiler"],
                              os.path.join(CSRC, "gemm_test.cu"), os.path.join(CSRC, "gemm_sm100.cu"), "-o", GEMM_TEST,
                              "-lcuda"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    for s, p in procs + [("gemm_test.cu", probe)]:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}:\n{out.decode()}")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode:
        raise RuntimeError(f"link failed:\n{r.stdout.decode()}")
    for o in objs:
        os.remove(o)
    open(stamp, "w").write(want)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
