#!/bin/bash
# round checkpoint: GPU tests, the default bench line (with cpu_baseline), the reference arm, the 512^3 benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -c 1200 gpurun_out/bench_ref.json
for w in "smoke_plume 512" "dambreak_solid 512" "flip_splash 512"; do
  set -- $w
  timeout 600 python bench.py --workload $1 --n $2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "bench $w rc=$?"; tail -c 1900 gpurun_out/bench_$1_$2.json; tail -3 gpurun_out/bench_$1_$2.err
done
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; grep -m1 "model name" /proc/cpuinfo
