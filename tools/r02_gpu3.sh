#!/bin/bash
# round 2, GPU call 3: TMA-store epilogue with the bias in the kernel parameters (constant bank)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_parity_gpu.py -x -q -k "tma_store" > gpurun_out/t3_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t3_parity.log
tail -3 gpurun_out/t3_parity.log
for shape in "81920 2048 512" "81920 1536 512"; do
  echo "== $shape"; timeout 300 python tools/ares_time.py $shape
done 2>&1 | tee gpurun_out/t3_ares.log
export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_dbg.so
for shape in "81920 2048 512" "81920 1536 512"; do
  for d in 0 1 3 7 4; do
    echo "== shape $shape dbg $d"
    ASR_EPI_DBG=$d timeout 120 python - $shape <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from asr_streaming_b200 import _lib
lib = _lib.load_library()
M, N, K = (int(a) for a in sys.argv[1:4])
ms = C.c_float()
rc = lib.asr_debug_gemm_time(M, N, K, 0, 515, 2, 20, C.byref(ms), 0)
print("  bn 515 rc", rc, "%.1f us  %.0f TFLOP/s" % (ms.value * 1e3, 2.0 * M * N * K / ms.value / 1e9) if not rc else lib.asr_last_error(), flush=True)
PY
  done
done 2>&1 | tee gpurun_out/t3_epi_dbg.log
unset ASR_B200_LIB
for v in 0 1; do
  ASR_B200_NO_TMA_STORE=$v timeout 600 python bench.py --workload streams4096 --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/t3_bench4096_notma$v.json 2> gpurun_out/t3_bench4096_notma$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/t3_bench4096_notma$v.json"))
print("NO_TMA_STORE=$v", d["ms_per_step"], d["value"], d["kernel_families_ms_per_step"], d["clocks"])
PY
done
