#!/bin/bash
# round 2, GPU call P: sampled radix select, leaf16 cooperative emit v2, whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 300 "partialsort tests" python -m pytest tests/test_sort_gpu.py -x -q -m gpu -k partialsort 2>&1 | tail -6
step 120 "select timing" python tools/select_probe.py 30 2>&1 | tail -5
step 200 "2^30 timings" env SORT_CHECK_ONE=uniform,f32:randf,f32:normalf,bigval SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -5
step 120 "2^27 sweep" python tools/sort_check.py 27 2>&1 | grep -v "ok=True" | tail -3
step 900 "whole GPU test suite" python -m pytest tests -x -q -m gpu 2>&1 | tail -4
