#!/bin/bash
# weak-scaling series on one 8-GPU box: N = 8, 4, 2 (256^3 per GPU), then the slab tests once more
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  echo "scale $n rc=$?"; tail -1 gpurun_out/scale_$n.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$n.json") if l.startswith("{")][-1])
    print("  N=%d ms/step %.3f value %.0f launches %d | solve %.2f ms it %d" % (d["n_gpus"], d["ms_per_step"], d["value"], d["gpu_launches"], d["solve"]["ms_solve"], d["solve"]["iterations"])); print("  ", d["roofline"]["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
done
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 200 2>&1 | tail -3
