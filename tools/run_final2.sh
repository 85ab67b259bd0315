cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "rc tests $?"; tail -2 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "rc smoke $?"; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "rc bench $?"
