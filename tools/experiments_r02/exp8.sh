python -m pytest tests/test_gpu_plugin.py -x -q 2>&1 | tail -5
python tools/module_timing.py smoke_plume 192
python tools/module_timing.py dambreak_solid 256
