# module GPUs=8 on thin slabs: what fails
python - <<'PY'
import os, sys
sys.path.insert(0, ".")
from oracle import refio
from shiokaze_b200 import scenes
for n, env in ((48, {}), (48, {"SHKZ_B200_NO_SLAB_PROLONG": "1"}), (64, {}), (48, {"CUDA_MODULE_LOADING": "LAZY"})):
    sc = scenes.smoke_plume(n)
    os.environ.update(env)
    try:
        r = refio.run_reference(sc, "f32", flags={"Residual": 1e-10, "Precision": "fp64", "GPUs": 8}, projection="b200pressure3", timeout=120)
        print("n", n, env, "ok iterations", r.iterations, flush=True)
    except Exception as e:
        print("n", n, env, "FAILED", str(e)[-900:], flush=True)
    for k in env: os.environ.pop(k, None)
PY
