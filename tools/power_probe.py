"""SM clock / board power while one GEMM variant runs back to back for ~1.5 s: python tools/power_probe.py
   (is the K = 512 GEMM + epilogue power-capped?  Compares the mainloop alone, the full epilogue, and a long-K shape.)"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np
import pynvml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_streaming_b200 import _lib  # noqa: E402

lib = _lib.load_library()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def probe(name, M, N, K, bn, epi, iters):
    stop = threading.Event()
    clk, pw = [], []

    def sample():
        while not stop.is_set():
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.01)
    th = threading.Thread(target=sample)
    th.start()
    ms = C.c_float()
    rc = lib.asr_debug_gemm_time(M, N, K, 0, bn, epi, iters, C.byref(ms), 0)
    stop.set()
    th.join()
    k = len(clk) // 2                                   # second half: the clock has settled
    print(f"{name:34s} rc {rc} {ms.value * 1e3:8.1f} us {2.0 * M * N * K / ms.value / 1e9:6.0f} TFLOP/s | SM clock median {np.median(clk[k:]):6.0f} MHz (min {min(clk[k:])}), "
          f"power median {np.median(pw[k:]):6.0f} W max {max(pw):6.0f} W, {len(clk)} samples", flush=True)


M = 81920
probe("warm-up", M, 2048, 512, 512, 3, 2000)
probe("K=512 N=2048 mainloop only", M, 2048, 512, 512, 3, 12000)
probe("K=512 N=2048 GELU->bf16 LSU", M, 2048, 512, 512, 2, 8000)
probe("K=512 N=2048 GELU->bf16 TMA store", M, 2048, 512, 515, 2, 8000)
probe("K=512 N=2048 fp32 store", M, 2048, 512, 512, 0, 6000)
probe("K=2048 N=2048 mainloop only", M, 2048, 2048, 512, 3, 3000)
probe("K=2048 N=2048 GELU->bf16 TMA store", M, 2048, 2048, 515, 2, 3000)
time.sleep(1.0)
print("idle: clock", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "MHz, power", pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, "W")
