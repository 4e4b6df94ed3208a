# r02w (2 GPUs): the slab + dense-array-core combination, with its error text
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -40 > $O/r02w_pytest_multi.txt; cat $O/r02w_pytest_multi.txt
python tools/module_timing.py dambreak_solid 256 2 > $O/r02w_module_dam256_gpus2.txt 2>&1; tail -30 $O/r02w_module_dam256_gpus2.txt
