#!/bin/bash
mkdir -p gpurun_out
timeout 420 python tools/gpu_tune2.py dambreak_solid:512 flip_splash:512 liquid_box:256 dambreak:256 > gpurun_out/tune3.log 2>&1; echo "tune rc=$?"; cat gpurun_out/tune3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_tma<1, 0, 1, 0, 0>|k_sweep_tma<1, 0, 0, 0, 0>|k_sweep_tma<1, 0, 0, 1, 0>" -s 3 -c 3 -o gpurun_out/prof_sweep_p -f python tools/profile_step.py smoke_plume 512 > gpurun_out/ncu_sweep_p.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sweep_p.log
ls -la gpurun_out/*.ncu-rep
