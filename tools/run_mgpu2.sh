cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 24 > gpurun_out/mgpu2_check.log 2>&1; echo "rc mgpu check $?"; tail -15 gpurun_out/mgpu2_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc bench2 $?"; head -c 400 gpurun_out/bench_2gpu.json
