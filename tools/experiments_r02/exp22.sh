# (2 GPUs) the module's GPUs=2 path, repeated: does the rare first-projection stall survive shkz_b200_prepare + eager module loading?
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python - <<'PY'
import os, sys, time
sys.path.insert(0, ".")
from oracle import refio
from shiokaze_b200 import scenes
ok = bad = 0
for sc, flags in ((scenes.smoke_plume(48), {"Residual": 1e-10, "Precision": "fp64"}), (scenes.dambreak(64, True), {}), (scenes.smoke_plume(96), {"Array": "b200array3"})):
    for rep in range(8):
        t = time.time()
        try:
            r = refio.run_reference(sc, "f32", flags={**flags, "GPUs": 2}, projection="b200pressure3", timeout=120, repeat=2)
            ok += 1
        except Exception as e:
            bad += 1
            print(sc.name, rep, "FAILED after", round(time.time() - t, 1), "s:", str(e)[-600:], flush=True)
print("module GPUs=2 runs ok", ok, "failed", bad)
PY
