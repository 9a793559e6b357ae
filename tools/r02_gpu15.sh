#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t15_all.log 2>&1; echo "rc=$?" >> gpurun_out/t15_all.log
tail -6 gpurun_out/t15_all.log
for dg in 0 1; do
ASR_B200_DEVICE_GATHER=$dg timeout 900 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/t15_bench_dg$dg.json 2> gpurun_out/t15_bench_dg$dg.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/t15_bench_dg$dg.json"))
print("device_gather=$dg", d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ragged"]["host_ms_per_tick"], d["ragged"]["gpu_busy_ms_per_pass"])
PY
done
timeout 600 python bench.py --workload longform --long-chunks 300 > gpurun_out/t15_longform300.json 2> gpurun_out/t15_longform300.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/t15_longform300.json'))
print(d['value'], d['e2e']['value'], d['e2e']['wall_s'], d['chunk_latency_ms'])"
