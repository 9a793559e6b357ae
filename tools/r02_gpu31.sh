#!/bin/bash
# round-2 final evidence: default bench line (all legs), ncu launch list + --set full capture of the same command, Kaldi fbank line
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_bench_final_1gpu.json 2> gpurun_out/t31_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/t31_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline"
$CMD > gpurun_out/t31_plain.json 2> gpurun_out/t31_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attention|fbank|ctc_greedy|beam_kernel|ln_|gather_blocks' -s 600 -c 400 --csv --log-file gpurun_out/r02_launches_ragged4096.csv $CMD > gpurun_out/t31_ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/t31_ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc2_kernel|gemm_ln_kernel|attention_stream|fbank_kernel|beam_kernel|ctc_greedy' -s 230 -c 16 -o gpurun_out/r02_prof_ragged4096 $CMD > gpurun_out/t31_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/t31_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:'beam_kernel|ctc_greedy' -s 4 -c 2 -o gpurun_out/r02_prof_decode $CMD > gpurun_out/t31_ncu_decode.log 2>&1
echo "decode rc=$?"; tail -2 gpurun_out/t31_ncu_decode.log
python bench.py --workload fbank1024 --steps 5 --warmup 3 > gpurun_out/r02_bench_fbank1024.json 2> gpurun_out/t31_fbank.err; echo "fbank rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
