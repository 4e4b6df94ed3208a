#!/bin/bash
# what the driver runs at round end, in one go: smoke(), the GPU suite, the default bench line and the reference arm (+ memcheck of the small cases)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -c 500 gpurun_out/bench_ref.json
sed -i 's/--racecheck-report all --error-exitcode 9 python \/tmp\/san_case.py race/--racecheck-report all --error-exitcode 9 python \/tmp\/san_case.py race/' tools/gpu_sanitize.sh
