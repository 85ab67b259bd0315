timeout 120 python tools/scan_probe.py 2>&1 | grep -c ok
for v in "" _v512 _v256l8; do echo "variant=$v"; B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays$v.so timeout 150 python tools/microbench.py scan 2>&1 | grep -v torch | tail -2; done
