python -m pytest tests/test_gpu_host_sparse.py -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r02s_bench_e2e.json 2> gpurun_out/r02s_bench_e2e.err
SHKZ_B200_HOST_COPIES=dense python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r02s_bench_e2e_dense.json 2> gpurun_out/r02s_bench_e2e_dense.err
python bench.py --workload flip_splash --n 512 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r02s_bench_e2e_flip.json 2> gpurun_out/r02s_bench_e2e_flip.err
ncu --set full --clock-control none -k regex:"k_sweep_tma|k_xpay_spmv_tma|k_axpy2_norm|k_residual_restrict" -c 26 -o gpurun_out/r02_ncu_solve_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -2
ncu --set full --clock-control none -k regex:"k_build_system|k_update_velocity|k_vcycle_mid|k_store_pressure|k_coarsen_operator" -s 5 -c 6 -o gpurun_out/r02_ncu_other_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_dam512.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r02_launches_bench.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/ | tail -8
for w in "dambreak_solid 512" "flip_splash 512" "dambreak 64"; do
python tools/gpu_profile_table.py $w > gpurun_out/r02s_table_$(echo $w | tr ' ' '_')_base.txt 2>&1
SHKZ_B200_MID_CELLS=16777216 python tools/gpu_profile_table.py $w > gpurun_out/r02s_table_$(echo $w | tr ' ' '_')_mid16m.txt 2>&1
done
