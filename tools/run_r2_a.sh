#!/bin/bash
# round 2, GPU call A: first run of the hybrid MSD sort + persistent LSD passes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$PWD/cuda.jl_b200/libb200arrays_tune.so
echo "== tune build, hybrid forced on small n"
B200ARRAYS_LIB=$T B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 timeout 300 python tools/sort_check.py 20 22 24 > gpurun_out/a_sc_small.log 2>&1; echo rc=$?
tail -50 gpurun_out/a_sc_small.log
echo "== product build"
timeout 600 python tools/sort_check.py 27 28 30 > gpurun_out/a_sc.log 2>&1; echo rc=$?
tail -60 gpurun_out/a_sc.log
echo "== LSD only (hybrid off), persistent passes"
B200ARRAYS_LIB=$T B200_SORT_HYBRID=0 SORT_CHECK_ONE=uniform timeout 300 python tools/sort_check.py 30 > gpurun_out/a_sc_lsd.log 2>&1; echo rc=$?
tail -3 gpurun_out/a_sc_lsd.log
echo "== pytest sort"
timeout 600 python -m pytest tests/test_sort_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== launch list 2^28"
SORT_CHECK_ONE=uniform timeout 300 python tools/sort_check.py 28 > gpurun_out/a_plain28.log 2>&1 && \
SORT_CHECK_ONE=uniform timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches28.csv python tools/sort_check.py 28 > gpurun_out/a_ncu28.log 2>&1
echo rc=$?
