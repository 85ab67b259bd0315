cd /root/repo
timeout 300 python tools/prof_one.py cumsum_f32 28 > gpurun_out/plain_scan_d.log 2>&1; echo "rc plain scan $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 2 -o gpurun_out/scan_chain2_f32_d -f python tools/prof_one.py cumsum_f32 28 > gpurun_out/ncu_scan_d.log 2>&1; echo "rc ncu scan $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 2 -o gpurun_out/scan_chain2_i64_d -f python tools/prof_one.py cumsum_i64 28 > gpurun_out/ncu_scan_i64_d.log 2>&1; echo "rc ncu scan i64 $?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | wc -l
