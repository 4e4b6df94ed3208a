#!/bin/bash
# full ncu capture of given kernels (second launch of each) on a workload. usage: gpu_ncu2.sh <workload> <n> <regex>...
mkdir -p gpurun_out
w=$1; n=$2; shift 2
for k in "$@"; do
  name=$(echo $k | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_${name}_${w}_${n} -f python tools/profile_step.py $w $n > gpurun_out/ncu_full_${name}.log 2>&1; echo "ncu full $k rc=$?"; tail -1 gpurun_out/ncu_full_${name}.log
done
