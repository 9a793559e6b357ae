#!/bin/bash
# gemm_ln: L2 prefetch of the next tile's A rows (new) vs none (libasr_b200_nopf.so), stand-alone and in the default bench, same box
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in new old new old; do
  if [ $v = old ]; then export ASR_B200_LIB=$PWD/asr_streaming_b200/libasr_b200_nopf.so; else unset ASR_B200_LIB; fi
  python tools/gemm_ln_knock.py 2>&1 | tail -1
  python bench.py --steps 10 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t42_bench_$v.json 2> gpurun_out/t42_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t42_bench_$v.json"))
f = d["kernel_families_ms_per_step"]
print("$v", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "ffn2", f["gemm_ffn2"], "out", f["gemm_out_proj"], "ffn1", f["gemm_ffn1"], "qkv", f["gemm_qkv"], d["clocks"]["sm_mhz"])
PY
done
