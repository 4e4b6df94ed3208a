#!/bin/bash
# ncu: launch list of one step + full captures of the named kernels.  usage: gpu_ncu.sh <workload> <n> <regex1> [regex2...]
mkdir -p gpurun_out
w=$1; n=$2; shift 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${w}_${n}.csv python tools/profile_step.py $w $n > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_launch.log
for k in "$@"; do
  name=$(echo $k | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/prof_${name}_${w}_${n} -f python tools/profile_step.py $w $n > gpurun_out/ncu_full_${name}.log 2>&1; echo "ncu full $k rc=$?"; tail -2 gpurun_out/ncu_full_${name}.log
done
ls -la gpurun_out | tail -20
