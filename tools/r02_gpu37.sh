#!/bin/bash
# attention: K fragments by ldmatrix.x4 (new) vs LDS.32 pairs (old), same box; bit-identity tests first
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t37_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t37_pytest.log
for v in new old new old; do
  if [ $v = old ]; then export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_old.so; else unset ASR_B200_LIB; fi
  python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t37_bench_$v.json 2> gpurun_out/t37_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t37_bench_$v.json"))
f = d["kernel_families_ms_per_step"]
print("$v", round(d["ms_per_step"],3), round(d["value"]), "attention", f["attention"], "qkv", f["gemm_qkv"], "out", f["gemm_out_proj"], d["clocks"]["sm_mhz"], d["ragged"]["host_ms_per_tick"].get("of which prestage (overlaps the running tick)"))
PY
done
