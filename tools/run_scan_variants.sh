cd /root/repo
timeout 900 python -m pytest tests/test_scan_gpu.py -x -q -m gpu > gpurun_out/scan_tests.log 2>&1; echo "rc tests $?"; tail -2 gpurun_out/scan_tests.log
timeout 300 python tools/microbench.py scan_dims 2>&1
