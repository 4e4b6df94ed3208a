#!/bin/bash
# configs[3]: flip_splash 512^3 cut into 4 and 2 z-slabs (gpurun --gpus 4)
mkdir -p gpurun_out
for n in 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --workload flip_splash --grid 512 --scaling strong --steps 3 --warmup 3 > gpurun_out/strong${n}_flip512.json 2> gpurun_out/strong${n}_flip512.err
  echo "strong $n rc=$?"; tail -1 gpurun_out/strong${n}_flip512.err | cut -c1-200
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/strong${n}_flip512.json") if l.startswith("{")][-1])
    print("  N=%d ms/step %.3f value %.0f launches %d | solve %.2f ms it %d rows %d" % (d["n_gpus"], d["ms_per_step"], d["value"], d["gpu_launches"], d["solve"]["ms_solve"], d["solve"]["iterations"], d["solve"]["n_rows"])); print("  ", d["roofline"]["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
done
