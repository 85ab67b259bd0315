cd /root/repo
B200_BENCH_FORCE_SHARDED=1 B200_BENCH_SHARDED_LOG2N=24 timeout 40 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_forced_sharded.json 2> gpurun_out/bench_forced_sharded.err; echo "rc $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_forced_sharded.json'))
print(d['value'], json.dumps(d['extras'])[:1500])"
tail -3 gpurun_out/bench_forced_sharded.err
