#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02_pytest_gpu.log
python tools/module_timing.py dambreak_solid 256 2>&1 | tee $O/r02_module_timing_dambreak256_gpus1.txt | cut -c1-200 | tail -22
