"""Builds libasr_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libasr_b200.so")
SOURCES = ["engine.cu", "sched.cu", "gemm_tcgen05.cu", "gemm_ln.cu", "fbank.cu", "layers.cu", "beam.cu", "convmod.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "kernels.cuh", "tc_ptx.cuh", "fft_regs.cuh", "sched_hooks.h", os.path.join("..", "..", "include", "asr_b200.h")]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """out / defines: a second, diagnostic build next to the product library (e.g. -DASR_EPI_TIMING), loaded with ASR_B200_LIB=<out>."""
    if out is None and not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
           "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden", "-shared", "-cudart", "static", "-o", out or LIB] + [f"-D{d}" for d in defines] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libasr_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return out or LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None, defines=defs))
