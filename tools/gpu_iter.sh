#!/bin/bash
# one development iteration on the GPU: the bitwise V-cycle test, the full GPU suite, 512^3 benches, optional ncu (args: kernel regexes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for w in "smoke_plume 512" "dambreak_solid 512" "smoke_plume 256"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $w rc=$?"; tail -3 gpurun_out/bench_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$1_$2.json"))
    print("  ms/step %.2f value %.0f e2e_ms %.1f (h2d %.1f d2h %.1f) | solve %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_h2d"], d["e2e"]["ms_d2h"], {k:(float("%.3g" % v) if isinstance(v,float) else v) for k,v in d["solve"].items()}))
    r=d["roofline"]; print("  roof", r["kernel"], "frac %.3f" % r["frac"], "avg_ms %.4f" % r["avg_launch_ms"], "launches", r["launches"], "whole", {k:round(v,3) for k,v in r["solve_whole"].items()})
    print("  ", r["by_kernel_ms"])
except Exception as e: print("  ERR", e)
PY
done
