#!/bin/bash
python tools/module_timing_advect.py dambreak_solid 256 2>&1 | tee gpurun_out/r02_module_timing_advect_dambreak256.txt | cut -c1-260
