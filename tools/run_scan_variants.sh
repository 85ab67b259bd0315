cd /root/repo
run() { name=$1; shift
  env "$@" timeout 300 python tools/scan_chain_probe.py $MODES > gpurun_out/scan_v_$name.log 2>&1; echo "rc $name $?"
  sed -i "s/^\[/[$name /" gpurun_out/scan_v_$name.log; }
MODES="time" run default X=1
MODES="check time" run lm B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_lm.so
grep -h "CHECK\|cumsum\|MISMATCH\|BAD\|Error" gpurun_out/scan_v_default.log gpurun_out/scan_v_lm.log
