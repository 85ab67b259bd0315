#!/bin/bash
# round 2, GPU call D: counting leaf (8/16-bit bins), uniform-row fast paths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_tune.so
echo "== small-n forced hybrid regression"
B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 timeout 300 python tools/sort_check.py 20 22 24 2>&1 | grep -v "ok=True" | tail -5
echo "== hybrid 2^30"
SORT_CHECK_ONE=uniform,dup16,hi16only,sorted,narrow8,normalf SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 30 2>&1 | tail -7
echo "== hybrid 2^29 (min_avg 4096) / 2^28 (2048)"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 29 2>&1 | tail -2
SORT_CHECK_ONE=uniform,narrow8 SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 28 2>&1 | tail -3
echo "== LSD only (hybrid off)"
B200_SORT_HYBRID=0 SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 30 2>&1 | tail -2
B200_SORT_HYBRID=0 SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 28 2>&1 | tail -2
unset B200ARRAYS_LIB
echo "== product lib: full sweep"
timeout 600 python tools/sort_check.py 27 30 2>&1 | grep -v "ok=True" | tail -5
echo "== pytest sort"
timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== launch list 2^30 hybrid"
SORT_CHECK_ONE=uniform timeout 300 python tools/sort_check.py 30 > gpurun_out/d_plain30.log 2>&1 && \
SORT_CHECK_ONE=uniform timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches30.csv python tools/sort_check.py 30 > gpurun_out/d_ncu30.log 2>&1
echo rc=$?
echo "== ncu full: msd kernels at 2^30"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:msd_ -c 7 -o gpurun_out/d_msd30 python tools/sort_check.py 30 > gpurun_out/d_ncufull.log 2>&1
echo rc=$?
