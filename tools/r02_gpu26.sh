#!/bin/bash
for args in "2 32 128 0.95" "2 32 804 0.95" "16 64 804 0.95" "16 64 804 1.0" "16 64 804 0.999999"; do
  python tools/probe_cuctc.py $args 2>&1 | tail -3
  echo "rc=$? ($args)"
done
