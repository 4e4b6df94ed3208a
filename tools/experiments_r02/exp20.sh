# module GPUs=N on smoke 48^3: which world fails, and how
python - <<'PY'
import os, sys
sys.path.insert(0, ".")
from oracle import refio
from shiokaze_b200 import scenes
import torch
n = torch.cuda.device_count()
sc = scenes.smoke_plume(48)
for world in (2, 4, 8):
    if world > n: continue
    for rep in range(3):
        try:
            r = refio.run_reference(sc, "f32", flags={"Residual": 1e-10, "Precision": "fp64", "GPUs": world}, projection="b200pressure3", timeout=120)
            print("world", world, "rep", rep, "ok iterations", r.iterations, flush=True)
        except Exception as e:
            print("world", world, "rep", rep, "FAILED", str(e)[-1500:], flush=True)
PY
