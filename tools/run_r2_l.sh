#!/bin/bash
# round 2, GPU call L: leaf count/emit rewrite — correctness sweep + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 300 "sort tests" python -m pytest tests/test_sort_gpu.py tests/test_golden_gpu.py tests/test_mgpu_sort_gpu.py -x -q -m gpu 2>&1 | tail -4
step 240 "product lib: full sweep" python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
step 200 "sort timings" env SORT_CHECK_ONE=uniform,sorted,dup16 SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -4
