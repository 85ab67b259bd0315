#!/bin/bash
# round 2, GPU call K (re-entry): whole GPU suite on the restored tree, bench N=1, sort timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 900 "whole GPU test suite" python -m pytest tests -x -q -m gpu 2>&1 | tail -12
step 200 "sort timings" env SORT_CHECK_ONE=uniform,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -4
echo "== bench"
timeout 600 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo rc=$?
tail -3 gpurun_out/k_bench.err
python tools/show_bench.py gpurun_out/k_bench.json
