cd /root/repo
for v in "" _v1; do
  B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays$v.so timeout 300 python tools/sort_probe.py 30 2>&1 | sed "s/^/[lib$v] /"
done
