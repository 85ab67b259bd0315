#!/bin/bash
# round 2, GPU call I: leaf emit v4, plan-driven sortperm partition, emulated multi-rank tests, whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$PWD/cuda.jl_b200/libb200arrays_tune.so
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 120 "small-n forced hybrid regression" env B200ARRAYS_LIB=$T B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 python tools/sort_check.py 20 24 2>&1 | grep -v "ok=True" | tail -5
step 120 "hybrid 2^30" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,sorted,dup16,narrow8 SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -4
step 90 "hybrid 2^29" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 python tools/sort_check.py 29 2>&1 | tail -1
step 90 "hybrid 2^28" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,narrow8 SORT_CHECK_REPS=3 python tools/sort_check.py 28 2>&1 | tail -2
step 240 "product lib: full sweep" python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
step 600 "whole GPU test suite" python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "== ncu full: leaf at 2^30"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 120 python tools/sort_check.py 30 > gpurun_out/i_plain30.log 2>&1 && \
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:msd_ -c 7 -o gpurun_out/i_msd30 python tools/sort_check.py 30 > gpurun_out/i_ncufull.log 2>&1
echo rc=$?
