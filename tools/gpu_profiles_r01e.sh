#!/bin/bash
# evidence for profiles/ (round 1, set e): the default bench line, the launch list of the same command, full captures of the top kernels (1 GPU)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.json
timeout 600 python bench.py --workload smoke_plume --n 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_smoke512.json 2> gpurun_out/bench_smoke512.err; echo "bench512 rc=$?"
timeout 600 python bench.py --workload dambreak_solid --n 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dambreak512.json 2> gpurun_out/bench_dambreak512.err; echo "benchdam rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_smoke256.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
N="--set full --clock-control none --import-source on --kernel-name-base demangled"
T=".int.%s, .bool.%s, .bool.%s, .bool.%s, .bool.0>"
cap() { # name regex skip count
  timeout 400 ncu $N -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 -f python tools/profile_step.py smoke_plume 512 > gpurun_out/$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap prof_sweep_z   "k_sweep_tma<$(printf "$T" 0 1 0 0)" 0 1
cap prof_sweep     "k_sweep_tma<$(printf "$T" 0 0 0 0)" 0 1
cap prof_sweep_p   "k_sweep_tma<$(printf "$T" 1 0 1 0)" 4 1
cap prof_sweep_d   "k_sweep_tma<$(printf "$T" 1 0 0 1)" 0 1
cap prof_cg        "k_spmv_dot4|k_axpy2_norm|k_residual_restrict|k_xpay" 0 4
ls -la gpurun_out/*.ncu-rep
