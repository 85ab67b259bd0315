#!/bin/bash
# round 2, GPU call H: leaf emission v3, P1/P2 tile-shape sweep, emulated multi-rank test, flaky cumsum investigation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$PWD/cuda.jl_b200/libb200arrays_tune.so
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 120 "small-n forced hybrid regression" env B200ARRAYS_LIB=$T B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 python tools/sort_check.py 20 24 2>&1 | grep -v "ok=True" | tail -5
for cfg in 0 1 2 3 4; do
  step 90 "hybrid 2^30 cfg=$cfg" env B200ARRAYS_LIB=$T B200_MSD_CFG=$cfg SORT_CHECK_ONE=uniform,sorted SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -2
done
step 120 "hybrid 2^30 dists" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=dup16,hi16only,narrow8 SORT_CHECK_REPS=3 python tools/sort_check.py 30 2>&1 | tail -3
step 90 "hybrid 2^29/2^28" env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 python tools/sort_check.py 29 2>&1 | tail -1
env B200ARRAYS_LIB=$T SORT_CHECK_ONE=uniform,narrow8 SORT_CHECK_REPS=3 timeout 90 python tools/sort_check.py 28 2>&1 | tail -2
step 240 "product lib: full sweep" python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
step 300 "pytest sort + golden + mgpu-emulated" python -m pytest tests/test_sort_gpu.py tests/test_golden_gpu.py tests/test_mgpu_sort_gpu.py -x -q -m gpu 2>&1 | tail -8
echo "== cumsum full size x3 (flaky?)"
for i in 1 2 3; do timeout 200 python -m pytest "tests/test_fullsize_gpu.py::test_cumsum_full_size" -x -q -m gpu 2>&1 | tail -15 > gpurun_out/h_cumsum_$i.log; tail -1 gpurun_out/h_cumsum_$i.log; done
step 200 "pytest fullsize sort" python -m pytest "tests/test_fullsize_gpu.py::test_sort_and_sortperm_full_size" -x -q -m gpu 2>&1 | tail -3
echo "== ncu full: leaf at 2^30"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 120 python tools/sort_check.py 30 > gpurun_out/h_plain30.log 2>&1 && \
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:msd_ -c 7 -o gpurun_out/h_msd30 python tools/sort_check.py 30 > gpurun_out/h_ncufull.log 2>&1
echo rc=$?
