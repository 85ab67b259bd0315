#!/bin/bash
# round 2, GPU call M: skew-robust hybrid (heavy-digit ranking, 16-bit leaf for segments of any length)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 120 "2^27 sweep" python tools/sort_check.py 27 2>&1 | tail -22
step 200 "2^30 float + skew" env SORT_CHECK_ONE=uniform,f32:randf,f32:normalf,skew2,bigval,dup16,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -8
step 300 "sort tests" python -m pytest tests/test_sort_gpu.py tests/test_golden_gpu.py tests/test_mgpu_sort_gpu.py -x -q -m gpu 2>&1 | tail -4
