#!/bin/bash
# round 2, GPU call G: probe + per-row fast paths, per-lane emit, emulated multi-rank sort test
# short timeout and the script stops at the first step that times out (a hung kernel must not eat the budget).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$PWD/cuda.jl_b200/libb200arrays_tune.so
step() {   # step <seconds> <label> <command...>
  local secs=$1 label=$2; shift 2
  echo "== $label"
  timeout "$secs" "$@"
  local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi
  return $rc
}
step 120 "small-n forced hybrid regression" env B200ARRAYS_LIB=$T B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 python tools/sort_check.py 20 24 2>&1 | grep -v "ok=True" | tail -5
[ ${PIPESTATUS[0]} -ne 0 ] && [ ${PIPESTATUS[0]} -ne 1 ] && exit 1
step 120 "hybrid 2^30" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,dup16,hi16only,sorted,narrow8,normalf SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -7
step 90 "hybrid 2^29" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 29 2>&1 | tail -3
step 90 "hybrid 2^28" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,narrow8,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 28 2>&1 | tail -4
step 90 "LSD only 2^30" env B200ARRAYS_LIB=$T B200_SORT_HYBRID=0 SORT_CHECK_ONE=uniform,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -3
step 240 "product lib: full sweep" python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
step 240 "pytest sort + golden + scan" python -m pytest tests/test_sort_gpu.py tests/test_fullsize_gpu.py tests/test_golden_gpu.py tests/test_scan_gpu.py tests/test_mgpu_sort_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== launch list 2^30 hybrid"
SORT_CHECK_ONE=uniform,sorted,dup16 timeout 120 python tools/sort_check.py 30 > gpurun_out/g_plain30.log 2>&1 && \
SORT_CHECK_ONE=uniform,sorted,dup16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches30.csv python tools/sort_check.py 30 > gpurun_out/g_ncu30.log 2>&1
echo rc=$?
echo "== ncu full: msd kernels at 2^30"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:msd_ -c 7 -o gpurun_out/g_msd30 python tools/sort_check.py 30 > gpurun_out/g_ncufull.log 2>&1
echo rc=$?
