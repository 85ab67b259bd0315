cd /root/repo
for v in "" _v5; do
  B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays$v.so timeout 300 python tools/sort_probe.py 30 2>&1 | sed "s/^/[lib$v] /"
done
timeout 300 python tools/microbench.py scan_dims 2>&1 | head -4
