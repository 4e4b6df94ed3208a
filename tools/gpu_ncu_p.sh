#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_sweep_tma<.int.1, .bool.0, .bool.1, .bool.0, .bool.0>" -s 4 -c 1 -o gpurun_out/prof_sweep_p -f python tools/profile_step.py smoke_plume 512 > gpurun_out/ncu_sweep_p.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sweep_p.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_sweep_tma<.int.0, .bool.0, .bool.0, .bool.0, .bool.0>" -s 0 -c 1 -o gpurun_out/prof_sweep_plain -f python tools/profile_step.py smoke_plume 512 > gpurun_out/ncu_sweep_plain.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sweep_plain.log
ls -la gpurun_out/*.ncu-rep
