#!/bin/bash
# device gather with the small persistent gather kernel vs host gather, one GPU (same box), + scheduler tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "scheduler or prestaged or soak or gather" > gpurun_out/t33_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t33_pytest.log
for v in 1 0 1 0; do
  ASR_B200_DEVICE_GATHER=$v python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/t33_bench_dg$v.json 2> gpurun_out/t33_bench_dg$v.err; echo "bench dg=$v rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/t33_bench_dg$v.json"))
print("device_gather=$v value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/pass", round(d["e2e"]["ms_per_step"], 3), "gpu busy", round(d["ragged"]["gpu_busy_ms_per_pass"], 3), d["ragged"]["batch_assembly"], d["ragged"]["host_ms_per_tick"])
PY
done
