#!/bin/bash
# round 2, multi-GPU call #3 (N = visible GPUs): emulated-rank tests, multi-GPU parity, stage timing, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 300 "emulated ranks" python -m pytest tests/test_mgpu_sort_gpu.py -x -q -m gpu 2>&1 | tail -5
step 400 "multi-GPU parity (pytest wrapper)" python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -30
step 300 "stage timing" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/mgpu_profile.py 30 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/mg_profile_$N.log | tail -30
echo "== bench.py --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_bench_$N.json 2> gpurun_out/mg_bench_$N.err
echo rc=$?
tail -5 gpurun_out/mg_bench_$N.err
python tools/show_bench.py gpurun_out/mg_bench_$N.json
