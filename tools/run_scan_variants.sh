cd /root/repo
timeout 900 python -m pytest tests/test_scan_gpu.py -x -q -m gpu > gpurun_out/scan_tests.log 2>&1; echo "rc tests $?"; tail -2 gpurun_out/scan_tests.log
timeout 300 python tools/scan_chain_probe.py check time 2>&1 | grep -h "CHECK\|cumsum\|MISMATCH\|BAD\|Error"
