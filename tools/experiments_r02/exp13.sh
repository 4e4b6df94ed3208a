# r02t: correctness of the sparse host copies + the cross-tile load stream, e2e lines, mid-kernel experiment, ncu captures (reports reduced to csv on the box)
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02t_pytest.txt; cat $O/r02t_pytest.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02t_bench_e2e.json 2> $O/r02t_bench_e2e.err
SHKZ_B200_HOST_COPIES=dense python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02t_bench_e2e_dense.json 2> $O/r02t_bench_e2e_dense.err
python bench.py --workload flip_splash --n 512 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02t_bench_e2e_flip.json 2> $O/r02t_bench_e2e_flip.err
python - <<'PY'
import json
for f in ("r02t_bench_e2e", "r02t_bench_e2e_dense", "r02t_bench_e2e_flip"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 3), "e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "what"}, "solve_whole", d["roofline"]["solve_whole"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
for w in "dambreak_solid 512" "flip_splash 512" "dambreak 64"; do
python tools/gpu_profile_table.py $w > $O/r02t_table_$(echo $w | tr ' ' '_')_base.txt 2>&1
SHKZ_B200_MID_CELLS=16777216 python tools/gpu_profile_table.py $w > $O/r02t_table_$(echo $w | tr ' ' '_')_mid16m.txt 2>&1
head -1 $O/r02t_table_$(echo $w | tr ' ' '_')_base.txt; head -1 $O/r02t_table_$(echo $w | tr ' ' '_')_mid16m.txt
done
ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -s 13 -c 13 -o $O/r02_ncu_solve_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -2
ncu -i $O/r02_ncu_solve_dam512.ncu-rep --page raw --csv > $O/r02_ncu_solve_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_solve_dam512.ncu-rep
ncu --set full --clock-control none -k regex:"k_build_system|k_update_velocity|k_vcycle_mid|k_store_pressure|k_coarsen_operator" -s 5 -c 6 -o $O/r02_ncu_other_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -2
ncu -i $O/r02_ncu_other_dam512.ncu-rep --page raw --csv > $O/r02_ncu_other_dam512_raw.csv 2>/dev/null; rm -f $O/r02_ncu_other_dam512.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench_dam512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02_launches_bench.log 2>&1
du -sh $O
