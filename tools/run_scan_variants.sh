cd /root/repo
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/scan_tests.log 2>&1; echo "rc tests $?"; tail -2 gpurun_out/scan_tests.log
timeout 300 python tools/microbench.py scan_dims scan > gpurun_out/microbench_scan_c.log 2>&1; echo "rc micro $?"; cat gpurun_out/microbench_scan_c.log
B200_SCAN_CHAIN=2 timeout 300 python tools/microbench.py scan_dims > gpurun_out/microbench_scan_c2.log 2>&1; echo "forced chain2:"; cat gpurun_out/microbench_scan_c2.log
