#!/bin/bash
python tools/gemm_ln_time.py 81920 65536 2>&1 | tail -16
