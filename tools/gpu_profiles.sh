#!/bin/bash
# evidence for profiles/: launch list of the bench command + full captures of the top kernels (1 GPU)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_smoke256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"; tail -c 300 gpurun_out/ncu_launch_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tma -s 0 -c 4 -o gpurun_out/prof_sweep_tma_smoke512 -f python tools/profile_step.py smoke_plume 512 > gpurun_out/ncu_full_sweep_tma.log 2>&1; echo "ncu full sweep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_dot4|k_axpy2_norm|k_residual_restrict|k_xpay" -s 0 -c 4 -o gpurun_out/prof_cg_smoke512 -f python tools/profile_step.py smoke_plume 512 > gpurun_out/ncu_full_cg.log 2>&1; echo "ncu full cg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_system|k_update_velocity|k_label_rows|k_face_fractions" -s 4 -c 4 -o gpurun_out/prof_asm_dambreak512 -f python tools/profile_step.py dambreak_solid 512 > gpurun_out/ncu_full_asm.log 2>&1; echo "ncu full asm rc=$?"
ls -la gpurun_out/*.ncu-rep
