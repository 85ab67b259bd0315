#!/bin/bash
# round 2, GPU call B: hybrid sort after the leaf load-chunking fix; LSD persistent vs per-tile CTAs; profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
export B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_tune.so
echo "== hybrid 2^30"
SORT_CHECK_ONE=uniform,dup16,hi16only,sorted,narrow8 SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 30 2>&1 | tail -6
echo "== hybrid 2^29 (min_avg 4096) / 2^28"
B200_MSD_MIN_AVG=4096 SORT_CHECK_ONE=uniform,narrow8 SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 29 2>&1 | tail -3
B200_MSD_MIN_AVG=2048 SORT_CHECK_ONE=uniform,narrow8 SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 28 2>&1 | tail -3
echo "== LSD persistent"
B200_SORT_HYBRID=0 B200_LSD_PERSIST=1 SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 30 2>&1 | tail -2
echo "== LSD one CTA per tile"
B200_SORT_HYBRID=0 B200_LSD_PERSIST=0 SORT_CHECK_ONE=uniform SORT_CHECK_REPS=3 timeout 300 python tools/sort_check.py 30 2>&1 | tail -2
echo "== small-n forced hybrid regression"
B200_MSD_MIN_LOG2N=16 B200_MSD_MIN_AVG=1 timeout 300 python tools/sort_check.py 20 24 2>&1 | tail -2
echo "== launch list 2^30 hybrid"
SORT_CHECK_ONE=uniform timeout 300 python tools/sort_check.py 30 > gpurun_out/b_plain30.log 2>&1 && \
SORT_CHECK_ONE=uniform timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches30.csv python tools/sort_check.py 30 > gpurun_out/b_ncu30.log 2>&1
echo rc=$?
echo "== ncu full: msd kernels at 2^30"
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:msd_ -c 7 -o gpurun_out/b_msd30 python tools/sort_check.py 30 > gpurun_out/b_ncufull.log 2>&1
echo rc=$?
ls -la gpurun_out/*.ncu-rep | tail -3
