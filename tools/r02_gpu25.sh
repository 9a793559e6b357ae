#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "torchaudio_cuda or decode_stage or prefix_beam" > gpurun_out/t25_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/t25_pytest.log
