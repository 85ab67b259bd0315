#!/bin/bash
# round 2, GPU call PROF: ncu --set full captures of the final kernels, summarised ON THE BOX (the reports themselves are
# too large for gpurun_out: only the .md summaries come back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none"
R=/tmp/r2prof; mkdir -p $R
echo "== plain runs first"
SORT_CHECK_ONE=uniform,f32:randf SORT_CHECK_REPS=2 timeout 200 python tools/sort_check.py 30 2>&1 | tail -2 || exit 1
timeout 200 python tools/prof_misc.py || exit 1
cap() { local name=$1 title=$2 regex=$3; shift 3; echo "== ncu $name"; timeout 500 $NCU -k "regex:$regex" -f -o $R/$name "$@" > gpurun_out/prof_$name.log 2>&1; echo rc=$?
  python tools/ncu_summary.py $R/$name.ncu-rep "$title" > gpurun_out/r2_ncu_$name.md 2>> gpurun_out/prof_$name.log; rm -f $R/$name.ncu-rep; }
SORT_CHECK_ONE=uniform SORT_CHECK_REPS=1 cap msd_uniform "Round 2 final: sort! 2^30 UInt32 uniform, hybrid MSD kernels" msd_ -c 6 python tools/sort_check.py 30
SORT_CHECK_ONE=f32:randf SORT_CHECK_REPS=1 cap msd_randf "Round 2 final: sort! 2^30 Float32 uniform [0,1) (heavy digits, 16-bit leaf), hybrid MSD kernels" msd_ -c 6 python tools/sort_check.py 30
cap select "Round 2: partialsort(c, k) 2^30 Float32, sampled radix select" "select_(filter|sample|hist)" -c 6 python tools/select_probe.py 30
cap mgpu "Round 2: multi-GPU sort kernels on one GPU (8 emulated ranks, 2^27 keys): plan + fused partition/exchange" "mgpu_(scatter|plan)" -c 8 python tools/mgpu_emul_time.py 27 8
cap misc "Round 2: strided look-back scan, argreduce, fused findall, pairs LSD pass" "(scan_strided_lb|argreduce|select_flagged|radix_onesweep)" -c 8 python tools/prof_misc.py
ls -la gpurun_out/ | tail -12
