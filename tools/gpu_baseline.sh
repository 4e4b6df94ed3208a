#!/bin/bash
# baseline of the current state: smoke, GPU tests, 512^3 probes (per-kernel table) for the two 512^3 configs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in "smoke_plume 512" "dambreak_solid 512" "smoke_plume 256"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $w rc=$?"; tail -c 2500 gpurun_out/bench_$1_$2.json; tail -3 gpurun_out/bench_$1_$2.err
done
