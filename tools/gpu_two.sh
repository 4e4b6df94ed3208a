#!/bin/bash
# 2-GPU check (gpurun --gpus 2): slab tests, then the weak-scaling bench at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -12 gpurun_out/pytest_multi2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale2.json 2> gpurun_out/scale2.err
echo "scale2 rc=$?"; tail -2 gpurun_out/scale2.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale2.json") if l.startswith("{")][-1])
    print("N=%d ms/step %.2f value %.0f | %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["solve"])); print(d["roofline"]["by_kernel_ms"])
except Exception as e: print("ERR", e)
PY
