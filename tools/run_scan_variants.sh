cd /root/repo
timeout 900 python -m pytest tests/test_scan_gpu.py -x -q -m gpu > gpurun_out/scan_tests.log 2>&1; echo "rc tests $?"; tail -3 gpurun_out/scan_tests.log
run() { name=$1; shift
  env "$@" timeout 300 python tools/scan_chain_probe.py $MODES > gpurun_out/scan_v_$name.log 2>&1; echo "rc $name $?"
  sed -i "s/^\[/[$name /" gpurun_out/scan_v_$name.log; }
MODES="time" run default X=1
MODES="check time" run k7 B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_k7.so
MODES="time" run k7_wide B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_k7.so B200_SCAN_CHAIN=2
grep -h "CHECK\|cumsum\|MISMATCH\|BAD\|Error" gpurun_out/scan_v_default.log gpurun_out/scan_v_k7*.log
