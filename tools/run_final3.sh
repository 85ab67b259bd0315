cd /root/repo
timeout 200 python tools/prof_one.py cumsum_i64 28 > gpurun_out/plain_scan_e.log 2>&1; echo "rc plain scan $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 1 -o gpurun_out/scan_chain2_i64_e -f python tools/prof_one.py cumsum_i64 28 > gpurun_out/ncu_scan_i64_e.log 2>&1; echo "rc ncu scan i64 $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_chain2 -c 1 -o gpurun_out/scan_chain2_f32_e -f python tools/prof_one.py cumsum_f32 28 > gpurun_out/ncu_scan_f32_e.log 2>&1; echo "rc ncu scan f32 $?"
