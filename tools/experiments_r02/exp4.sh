python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -k "edge or device_entry or tight or fp32_build" 2>&1 | tail -3
for w in "dambreak_solid 512" "flip_splash 512" "smoke_plume 256"; do
python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|build_system|update_velocity"
done
which ncu
ncu --set full --clock-control none --import-source on -k regex:"k_build_system|k_update_velocity" -s 4 -c 2 -o gpurun_out/r02f_asm_dam512 -f python tools/profile_step.py dambreak_solid 512 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:"k_build_system|k_update_velocity" -s 4 -c 2 -o gpurun_out/r02f_asm_smoke512 -f python tools/profile_step.py smoke_plume 512 2>&1 | tail -5
ls -la gpurun_out/
