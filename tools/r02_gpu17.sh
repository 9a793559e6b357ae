#!/bin/bash
# round 2 evidence: ncu launch list + --set full capture of the layer kernels of the default bench command, Kaldi fbank capture
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline"
$CMD > gpurun_out/t17_plain.json 2> gpurun_out/t17_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attention|fbank|ctc_greedy|beam_kernel|ln_' -s 600 -c 400 --csv --log-file gpurun_out/r02_launches_ragged4096.csv $CMD > gpurun_out/t17_ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/t17_ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel|gemm_ln_kernel|attention_stream|fbank_kernel|beam_kernel|ctc_greedy' -s 230 -c 16 -o gpurun_out/r02_prof_ragged4096 $CMD > gpurun_out/t17_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/t17_ncu_full.log
CMD2="python bench.py --workload fbank1024 --steps 5 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/t17_fbank_plain.json 2> gpurun_out/t17_fbank_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -s 4 -c 2 -o gpurun_out/r02_prof_fbank1024 $CMD2 > gpurun_out/t17_ncu_fbank.log 2>&1
echo "fbank rc=$?"; cat gpurun_out/t17_fbank_plain.json | cut -c1-600
ls -la gpurun_out/*.ncu-rep | tail -3
