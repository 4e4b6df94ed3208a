#!/bin/bash
# weak-scaling bench under torchrun for N in "$@" (gpurun --gpus max(N))
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "scale $n rc=$?"; tail -4 gpurun_out/scale_$n.err | cut -c1-400
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$n.json") if l.startswith("{")][-1])
    print("  N=%d ms/step %.2f value %.0f Mcells/s | solve %s" % (d["n_gpus"], d["ms_per_step"], d["value"], {k:(float("%.3g" % v) if isinstance(v,float) else v) for k,v in d["solve"].items()}))
    if d.get("roofline"): print("  ", d["roofline"]["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
done
