#!/bin/bash
# per-stream cost of a step vs batch size at wave-friendly sizes: would slices whose intermediates stay in L2 beat one full-size step?
python tools/step_sweep.py 448 896 960 1024 1408 1856 2048 2816 4096 2>&1 | tail -12
