python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in "dambreak_solid 512" "flip_splash 512" "smoke_plume 256" "smoke_plume 512" "dambreak 64"; do
python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|build_system|update_velocity|store_pressure|coarsen_operator@0|compact_tiles@0"
done
