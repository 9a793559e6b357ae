#!/bin/bash
# re-entry sanity run: GPU tests with durations, smoke, default bench line
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/t20_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t20_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t20_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/t20_smoke.log
( time python bench.py ) > gpurun_out/t20_bench.json 2> gpurun_out/t20_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/t20_bench.err; cut -c1-1500 gpurun_out/t20_bench.json
