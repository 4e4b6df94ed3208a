#!/bin/bash
# final pass of round 2 on one B200 after the last source change: GPU suite, sanitizers, ncu captures for traffic.json, default bench line
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -s 13 -c 13 -o $O/r02_ncu_solve_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -1
ncu -i $O/r02_ncu_solve_dam512.ncu-rep --page raw --csv > $O/r02_ncu_solve_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_solve_dam512.ncu-rep
ncu --set full --clock-control none -k regex:"k_build_system|k_update_velocity|k_store_pressure" -c 6 -o $O/r02_ncu_asm_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -1
ncu -i $O/r02_ncu_asm_dam512.ncu-rep --page raw --csv > $O/r02_ncu_asm_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_asm_dam512.ncu-rep
python tools/ncu_traffic.py $O/r02_ncu_solve_dam512_raw.csv 17995468 r02 "dambreak_solid 512^3 (18.0 M unknowns in 1384 tiles of 64x16x16)" $O/r02_ncu_traffic_summary.csv | tail -8
bash tools/gpu_sanitize.sh 2>&1 | tail -24
timeout 600 python bench.py > $O/r02_bench_dambreak512.json 2> $O/r02_bench_dambreak512.err; echo "bench rc=$?"; tail -2 $O/r02_bench_dambreak512.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_dambreak512.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", r["frac"], "traffic", r["traffic"], r["traffic_source"][:60], "solve_whole", r["solve_whole"]["frac"])
PY
du -sh $O
