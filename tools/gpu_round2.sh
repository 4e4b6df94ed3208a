#!/bin/bash
# Round-2 evidence run on ONE B200: what the driver runs at round end (smoke, GPU suite, default bench, reference arm) + the other bench lines,
# the ncu captures (reduced to csv on the box: a whole-solve capture at 512^3 is > 64 MB) and the sanitizer runs. Everything lands in gpurun_out/r02_*.
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
timeout 600 python bench.py > $O/r02_bench_dambreak512.json 2> $O/r02_bench_dambreak512.err; echo "bench rc=$?"; tail -2 $O/r02_bench_dambreak512.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "bench ref rc=$?"
SHKZ_B200_HOST_COPIES=dense timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02_bench_dambreak512_dense_copies.json 2> /dev/null
timeout 600 python bench.py --workload flip_splash --n 512 --no-cpu-baseline --no-sub-records > $O/r02_bench_flipsplash512.json 2> /dev/null
timeout 600 python bench.py --workload smoke_plume --n 512 --no-cpu-baseline --no-sub-records > $O/r02_bench_smoke512.json 2> /dev/null
timeout 600 python bench.py --workload dambreak --n 64 --no-cpu-baseline --no-sub-records > $O/r02_bench_dambreak64.json 2> /dev/null
timeout 600 python bench.py --workload smoke_plume --n 256 --no-cpu-baseline --no-sub-records > $O/r02_bench_smoke256.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("/")[-1], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"].get("ms_per_step", 0), 2), d["e2e"].get("host_copies"), "iters", (d.get("solve") or {}).get("iterations"),
              "dom", r.get("kernel"), r.get("frac"), "traffic", r.get("traffic"), "solve_whole", (r.get("solve_whole") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -s 13 -c 13 -o $O/r02_ncu_solve_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -1
ncu -i $O/r02_ncu_solve_dam512.ncu-rep --page raw --csv > $O/r02_ncu_solve_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_solve_dam512.ncu-rep
ncu --set full --clock-control none -k regex:"k_build_system|k_update_velocity|k_vcycle_mid|k_store_pressure|k_coarsen_operator" -s 5 -c 6 -o $O/r02_ncu_other_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -1
ncu -i $O/r02_ncu_other_dam512.ncu-rep --page raw --csv > $O/r02_ncu_other_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_other_dam512.ncu-rep
ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -s 13 -c 9 -o $O/r02_ncu_solve_smoke512 -f python tools/profile_step.py smoke_plume 512 2>&1 | tail -1
ncu -i $O/r02_ncu_solve_smoke512.ncu-rep --page raw --csv > $O/r02_ncu_solve_smoke512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_solve_smoke512.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench_dam512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02_launches_bench.log 2>&1
sed -i 's/timeout 600 compute-sanitizer/timeout 300 compute-sanitizer/; s/timeout 900 compute-sanitizer/timeout 420 compute-sanitizer/' tools/gpu_sanitize.sh
bash tools/gpu_sanitize.sh 2>&1 | tail -30
du -sh $O
