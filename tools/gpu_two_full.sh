#!/bin/bash
# 2-GPU box: whole GPU suite (single + slab tests), 1-GPU benches, weak-scaling bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("  N=%d ms/step %.3f value %.0f | %s" % (d["n_gpus"], d["ms_per_step"], d["value"], {k:(float("%.4g" % v) if isinstance(v,float) else v) for k,v in d["solve"].items()}))
    r=d["roofline"]; print("  dom", r["kernel"], "frac %.3f" % r["frac"], "whole %.3f" % r["solve_whole"]["frac"]); print("  ", r["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
}
for w in "smoke_plume 256" "smoke_plume 512"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_$1_$2.json 2> gpurun_out/q_$1_$2.err; echo "bench $w rc=$?"; tail -3 gpurun_out/q_$1_$2.err
  show gpurun_out/q_$1_$2.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale2.json 2> gpurun_out/scale2.err
echo "scale2 rc=$?"; tail -2 gpurun_out/scale2.err | cut -c1-300; show gpurun_out/scale2.json
