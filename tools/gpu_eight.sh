#!/bin/bash
# 8-GPU validation (gpurun --gpus 8): slab tests at world 2/4/8, weak scaling 256^3 per GPU, configs[4] liquid_box 1024^3 cut into 8 slabs
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 240 > gpurun_out/pytest_multi8.log 2>&1; echo "pytest multi rc=$?"; tail -8 gpurun_out/pytest_multi8.log
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; tail -2 gpurun_out/$name.err | cut -c1-300; tail -c 1500 gpurun_out/$name.json
}
run scale8_smoke256 8 --steps 5 --warmup 3
run strong8_box1024 8 --workload liquid_box --grid 1024 --scaling strong --steps 2 --warmup 3
