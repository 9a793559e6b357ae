#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/t46.err; echo "bench rc=$?"; tail -4 gpurun_out/t46.err
( time python bench.py --impl reference ) > gpurun_out/t46_ref.json 2> gpurun_out/t46_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/t46_ref.err; cut -c1-400 gpurun_out/t46_ref.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_final_1gpu.json"))
print(d["steps"], d["value"], d["ms_per_step"], d["e2e"], d["clocks"], d["gpu_launches"])
PY
