#!/bin/bash
# round 2, GPU call Q: select filter v4 tests + timing, look-back width variants of the MSD partition passes, ncu of mgpu_scatter
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_sort_gpu.py -x -q -m gpu -k partialsort 2>&1 | tail -3
timeout 120 python tools/select_probe.py 30 2>&1 | head -3
for v in "" _look2 _look8 _look16; do
  echo "== lib$v"; B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays$v.so SORT_CHECK_ONE=uniform,f32:randf SORT_CHECK_REPS=3 timeout 120 python tools/sort_check.py 30 2>&1 | tail -2
done
R=/tmp/r2prof; mkdir -p $R
timeout 300 ncu --set full --clock-control none -k regex:mgpu_scatter -c 2 -f -o $R/mgpu_scatter python tools/mgpu_emul_time.py 27 8 > gpurun_out/prof_mgpu_scatter.log 2>&1; echo rc=$?
python tools/ncu_summary.py $R/mgpu_scatter.ncu-rep "Round 2: b200_mgpu_scatter on one GPU (8 emulated ranks, 2^27 keys, all receive buffers local)" > gpurun_out/r2_ncu_mgpu_scatter.md
