#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -k "peaked or exact_precision_full or scheduler_ticks or prefix_beam" > gpurun_out/t10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t10_tests.log
tail -25 gpurun_out/t10_tests.log
timeout 900 python bench.py > gpurun_out/t10_bench_default.json 2> gpurun_out/t10_bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/t10_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/t10_bench_default.json"))
print({k:(v if not isinstance(v,(dict,list)) else '...') for k,v in d.items()})
print("exact", d.get("exact"))
print("configs", json.dumps(d.get("configs"))[:3000])
print("kernel_rooflines", json.dumps(d.get("kernel_rooflines"))[:3000])
print("cpu", d.get("cpu_baseline"))
PY
timeout 600 python bench.py --workload longform --long-chunks 200 > gpurun_out/t10_bench_longform200.json 2> gpurun_out/t10_bench_longform200.err; echo "longform rc=$?"; tail -3 gpurun_out/t10_bench_longform200.err; cat gpurun_out/t10_bench_longform200.json | cut -c1-1500
