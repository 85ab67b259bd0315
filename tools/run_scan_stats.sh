cd /root/repo
export B200ARRAYS_LIB=$PWD/cuda.jl_b200/libb200arrays_stats.so PROBE_LG=28
for c in 1 0; do B200_SCAN_CHAIN=$c timeout 200 python tools/scan_chain_probe.py time > gpurun_out/scan_stats_$c.log 2>&1; echo "rc $c $?"; done
B200_SCAN_CHAIN=1 B200_SCAN_CHAIN_AT=8 timeout 200 python tools/scan_chain_probe.py time > gpurun_out/scan_stats_at8.log 2>&1
for f in gpurun_out/scan_stats_1.log gpurun_out/scan_stats_0.log gpurun_out/scan_stats_at8.log; do echo == $f; grep "scan stats" $f | sed -n '3p;5p;12p;14p'; grep cumsum $f; done
