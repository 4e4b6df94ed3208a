python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in "dambreak_solid 512" "flip_splash 512" "smoke_plume 256" "smoke_plume 512"; do
python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|xpay|spmv|axpy2"
SHKZ_B200_NO_SPMV_TMA=1 python tools/gpu_profile_table.py $w 2>&1 | grep -E "iters|xpay|spmv"
done
