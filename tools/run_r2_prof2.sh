#!/bin/bash
# round 2, last GPU call: ncu launch list of the bench command (final state) + refreshed full capture of the final sort kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_bench_noextras.json 2> /dev/null || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_bench_launches.csv "Round 2: kernel launch list of python bench.py --steps 2 --warmup 3 --no-extras (final state)" > gpurun_out/r2_bench_launches_summary.md 2>/dev/null; head -30 gpurun_out/r2_bench_launches_summary.md
R=/tmp/r2prof; mkdir -p $R
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 timeout 300 ncu --set full --clock-control none -k regex:msd_ -c 9 -f -o $R/msd_final python tools/sort_check.py 30 > gpurun_out/prof_msd_final.log 2>&1; echo rc=$?
python tools/ncu_summary.py $R/msd_final.ncu-rep "Round 2 FINAL code: sort! 2^30 UInt32 uniform, hybrid MSD kernels (plain P1/P2 instantiation, 28 keys per thread)" > gpurun_out/r2_ncu_msd_uniform_final.md
