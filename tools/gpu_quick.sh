#!/bin/bash
# quick 1-GPU check: GPU tests + the 256^3 and 512^3 smoke benches (kernel table only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in "smoke_plume 256" "smoke_plume 512" "dambreak_solid 512"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_$1_$2.json 2> gpurun_out/q_$1_$2.err; echo "bench $w rc=$?"; tail -3 gpurun_out/q_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/q_$1_$2.json") if l.startswith("{")][-1])
    print("  ms/step %.3f value %.0f e2e %.0f | %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"] or 0, {k:(float("%.4g" % v) if isinstance(v,float) else v) for k,v in d["solve"].items()}))
    r=d["roofline"]; print("  dom", r["kernel"], "frac %.3f" % r["frac"], "whole %.3f" % r["solve_whole"]["frac"]); print("  ", r["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
done
