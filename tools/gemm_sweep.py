"""GEMM microbenchmark sweep on the GPU box (diagnostic): python tools/gemm_sweep.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200 import _lib

lib = _lib.load_library()


def t(M, N, K, bn, epi, split=0, iters=10):
    ms = C.c_float()
    rc = lib.asr_debug_gemm_time(M, N, K, split, bn, epi, iters, C.byref(ms), 0)
    if rc:
        return None
    return ms.value


cases = [("qkv", 1536, 512), ("out", 512, 512), ("ffn1", 2048, 512), ("ffn2", 512, 2048), ("big", 4096, 4096)]
for M in (5120, 20480, 81920):
    for name, N, K in cases:
        row = []
        for bn in (128, 256, 512):
            for epi in (0, 1, 2):
                ms = t(M, N, K, bn, epi)
                tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12 if ms else float("nan")
                row.append(f"bn{bn}/e{epi}: {ms*1e3:7.1f}us {tf:6.0f}TF")
        print(f"M={M:6d} {name:5s} N={N:5d} K={K:5d} | " + " | ".join(row), flush=True)
