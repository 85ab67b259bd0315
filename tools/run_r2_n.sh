#!/bin/bash
# round 2, GPU call N: per-kernel times of the hybrid path on skewed inputs (ncu launch list; times are cold and serialised)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for spec in "${@:-30:f32:randf 30:f32:normalf}"; do
  lg=${spec%%:*}; d=${spec#*:}
  echo "== 2^$lg $d"
  SORT_CHECK_ONE=$d SORT_CHECK_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msd_ --csv --log-file gpurun_out/n_launch_${lg}_${d#*:}.csv python tools/sort_check.py $lg > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/n_launch_${lg}_${d#*:}.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    print("  %-48s %9.3f ms" % (r[ki][:48], float(r[vi].replace(",",""))/1e6))
PY
done
