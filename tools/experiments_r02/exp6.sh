python -m pytest tests/test_gpu_parity.py -x -q -k "fused or tight or fp32_build" 2>&1 | tail -3
for w in "dambreak_solid 512" "flip_splash 512" "dambreak 64" "smoke_plume 128"; do
python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|vcycle"
done
SHKZ_B200_BZ=16 python tools/gpu_profile_table.py dambreak_solid 512 2>&1 | grep -E "iters|vcycle|sweep@1"
SHKZ_B200_BZ=8 python tools/gpu_profile_table.py dambreak_solid 512 2>&1 | grep -E "iters|vcycle|sweep@1"
SHKZ_B200_BZ=16 python tools/gpu_profile_table.py flip_splash 512 2>&1 | grep -E "iters|vcycle|sweep@1"
SHKZ_B200_BZ=8 python tools/gpu_profile_table.py flip_splash 512 2>&1 | grep -E "iters|vcycle|sweep@1"
SHKZ_B200_BZ=16 python tools/gpu_profile_table.py smoke_plume 512 2>&1 | grep -E "iters"
python tools/gpu_profile_table.py smoke_plume 512 2>&1 | grep -E "iters"
SHKZ_B200_MID_CELLS=20000000 python tools/gpu_profile_table.py dambreak_solid 512 2>&1 | grep -E "iters|vcycle|sweep@1"
