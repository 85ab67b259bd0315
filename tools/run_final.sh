# Round-end validation on one B200: GPU tests, smoke, bench (both arms), ncu launch list, ncu --set full captures.
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "rc tests $?"; tail -3 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "rc smoke $?"; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "rc bench $?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; echo "rc bench ref $?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain_bench_c.log 2>&1; echo "rc plain bench $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_c.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_bench_c.log 2>&1; echo "rc ncu list $?"
timeout 300 python tools/prof_one.py cumsum_f32 28 > gpurun_out/plain_scan_c.log 2>&1; echo "rc plain scan $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 2 -o gpurun_out/scan_chain2_f32 -f python tools/prof_one.py cumsum_f32 28 > gpurun_out/ncu_scan_c.log 2>&1; echo "rc ncu scan $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 2 -o gpurun_out/scan_chain2_i64 -f python tools/prof_one.py cumsum_i64 28 > gpurun_out/ncu_scan_i64_c.log 2>&1; echo "rc ncu scan i64 $?"
timeout 300 python tools/prof_one.py sort_u32 26 > gpurun_out/plain_sort_c.log 2>&1; echo "rc plain sort $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix -c 6 -o gpurun_out/sort_r1c -f python tools/prof_one.py sort_u32 26 > gpurun_out/ncu_sort_c.log 2>&1; echo "rc ncu sort $?"
cat gpurun_out/bench_r1c.json | head -c 600
