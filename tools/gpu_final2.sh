#!/bin/bash
# last pass of round 2 on one B200 after the advection work: smoke(), the whole GPU suite, sanitizers (now with the advection kernels), default bench line, reference arm
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
bash tools/gpu_sanitize.sh 2>&1 | tail -30
timeout 600 python bench.py > $O/r02_bench_dambreak512.json 2> $O/r02_bench_dambreak512.err; echo "bench rc=$?"; tail -2 $O/r02_bench_dambreak512.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "bench ref rc=$?"; tail -c 400 $O/r02_bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_dambreak512.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", r["frac"], "traffic", r["traffic"], "solve_whole", r["solve_whole"]["frac"])
print(json.dumps(d["sub_records"].get("advect_vector_512"))[:1500])
PY
