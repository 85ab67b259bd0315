#!/bin/bash
# round 2, multi-GPU call (N = number of visible GPUs): parity of the sharded paths against the oracle, then bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
step() { local secs=$1 label=$2; shift 2; echo "== $label"; timeout "$secs" "$@"; local rc=$?
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "TIMEOUT in '$label' — stopping"; exit 1; fi; return $rc; }
step 400 "multi-GPU parity (pytest wrapper)" python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -30
step 300 "whole GPU suite on one GPU" python -m pytest tests -x -q -m gpu --deselect tests/test_multigpu_gpu.py 2>&1 | tail -5
echo "== bench.py --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_bench_$N.json 2> gpurun_out/mg_bench_$N.err
echo rc=$?
tail -5 gpurun_out/mg_bench_$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/mg_bench_$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k, v in d.get("extras", {}).get("sharded", {}).items() if isinstance(d.get("extras", {}).get("sharded"), dict) else []:
        print(k, v)
    if not isinstance(d.get("extras", {}).get("sharded"), dict): print(d.get("extras"))
except Exception as e:
    print("no json:", e)
PY
