#!/bin/bash
# round 2, GPU call J: look-back prefetch, 3-CTA tile shapes, workspace layout fix, whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$PWD/cuda.jl_b200/libb200arrays_tune.so
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 120 "small-n forced hybrid regression" env B200ARRAYS_LIB=$T B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 python tools/sort_check.py 20 24 2>&1 | grep -v "ok=True" | tail -5
for cfg in 0 4 5; do
  step 90 "hybrid 2^30 cfg=$cfg" env B200ARRAYS_LIB=$T B200_MSD_CFG=$cfg SORT_CHECK_ONE=uniform,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -2
done
step 240 "product lib: full sweep" python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
step 600 "whole GPU test suite" python -m pytest tests -x -q -m gpu 2>&1 | tail -8
