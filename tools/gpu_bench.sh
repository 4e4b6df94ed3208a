#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the top kernel
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; tail -3 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rbgs -s 1 -c 2 -o gpurun_out/prof_rbgs -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_dot -s 0 -c 1 -o gpurun_out/prof_spmv -f python tools/profile_step.py > gpurun_out/ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
ls -la gpurun_out
