timeout 300 python -m pytest tests/test_sort_gpu.py -m gpu -x -q --timeout 90 --timeout-method=thread 2>&1 | tail -2
for v in "" _a; do echo "variant=$v"; B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays$v.so timeout 150 python tools/microbench.py sort 2>&1 | grep -v torch | tail -4; done
