#!/bin/bash
# closing pass of round 2 on the final sources: fresh ncu --set full capture of the level-0 solve kernels (profiles/traffic.json is keyed to the source hash),
# smoke(), default bench line, reference arm
O=gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -s 13 -c 13 -o $O/r02_ncu_solve_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -1
ncu -i $O/r02_ncu_solve_dam512.ncu-rep --page raw --csv > $O/r02_ncu_solve_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_solve_dam512.ncu-rep
python tools/ncu_traffic.py $O/r02_ncu_solve_dam512_raw.csv 17995468 r02 "dambreak_solid 512^3 (18.0 M unknowns in 1384 tiles of 64x16x16)" $O/r02_ncu_traffic_summary.csv | tail -3
timeout 600 python bench.py > $O/r02_bench_dambreak512.json 2> $O/r02_bench_dambreak512.err; echo "bench rc=$?"; tail -2 $O/r02_bench_dambreak512.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_dambreak512.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", r["frac"], "traffic", r["traffic"], "solve_whole", r["solve_whole"]["frac"])
a = d["sub_records"]["advect_vector_512"]
print("advect", round(a["ms_per_step"], 3), round(a["e2e_ms_per_step"], 2), a["roofline"]["frac"], "smoke256", round(d["sub_records"]["smoke_plume_256"]["ms_per_step"], 3))
r = json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["unit"], r["ms_per_step"])
PY
