cd /root/repo
run() { name=$1; shift
  env "$@" timeout 300 python tools/scan_chain_probe.py $MODES > gpurun_out/scan_v_$name.log 2>&1; echo "rc $name $?"
  sed -i "s/^\[/[$name /" gpurun_out/scan_v_$name.log; }
MODES="check time" run at2 B200_SCAN_CHAIN_AT=2
MODES="time" run at1 B200_SCAN_CHAIN_AT=1
grep -h "CHECK\|cumsum\|MISMATCH\|BAD\|Error" gpurun_out/scan_v_at2.log gpurun_out/scan_v_at1.log
