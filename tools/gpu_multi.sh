#!/bin/bash
# multi-GPU validation (gpurun --gpus N): slab tests, then the single-GPU suite as a regression check
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; cat gpurun_out/gpus.txt | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 240 > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -30 gpurun_out/pytest_multi.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest single rc=$?"; tail -5 gpurun_out/pytest_gpu.log
