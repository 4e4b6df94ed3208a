for w in "dambreak_solid 512" "flip_splash 512"; do
echo "=== base $w"; python tools/gpu_profile_table.py $w | head -14
echo "=== balanced $w"; SHKZ_B200_STENCIL_BALANCED=1 python tools/gpu_profile_table.py $w | head -14
echo "=== bz16 $w"; SHKZ_B200_BZ=16 python tools/gpu_profile_table.py $w | head -14
echo "=== bz16+balanced $w"; SHKZ_B200_BZ=16 SHKZ_B200_STENCIL_BALANCED=1 python tools/gpu_profile_table.py $w | head -14
echo "=== bz8+balanced $w"; SHKZ_B200_BZ=8 SHKZ_B200_STENCIL_BALANCED=1 python tools/gpu_profile_table.py $w | head -14
echo "=== coarse1 $w"; SHKZ_B200_COARSE_SWEEPS=1 python tools/gpu_profile_table.py $w | head -3
done
echo "=== smoke256 balanced"; SHKZ_B200_STENCIL_BALANCED=1 python tools/gpu_profile_table.py smoke_plume 256 | head -8
echo "=== smoke256 base"; python tools/gpu_profile_table.py smoke_plume 256 | head -8
