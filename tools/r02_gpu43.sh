#!/bin/bash
# the all-kernel-families smoke script (written for a memory checker, which is closed on this pool) run plain: decode stage on tied logits,
# pre-staged ticks with device gather + speculative fbank
timeout 600 python tools/sanitize_smoke.py 2>&1 | tail -5; echo "rc=$?"
