python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -k "assembly or edge or device_entry or tight or warm" 2>&1 | tail -5
for w in "dambreak_solid 512" "flip_splash 512" "smoke_plume 256" "smoke_plume 512"; do
python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|build_system|update_velocity|face_fractions|store_pressure|coarsen_operator@0|surface"
done
