"""Full per-kernel CUDA-event table of one project() (development aid). usage: python tools/gpu_profile_table.py workload n [Flag=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from shiokaze_b200 import MacPressureSolver3, scenes
w, n = sys.argv[1], int(sys.argv[2])
flags = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    flags[k] = float(v) if "." in v or "e" in v else (int(v) if v.lstrip("-").isdigit() else v)
sc = scenes.BENCH_SCENES[w](n)
dev = torch.device("cuda", 0)
S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, **flags)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
vel0, act0 = [t(v) for v in sc.vel], [t(a) for a in sc.vel_active]
fluid, solid = t(sc.fluid), (t(sc.solid) if sc.solid is not None else None)
pres = torch.zeros(sc.fluid.shape, dtype=torch.float32, device=dev)
pact = torch.zeros(sc.fluid.shape, dtype=torch.uint8, device=dev)
def step():
    vel, act = [v.clone() for v in vel0], [a.clone() for a in act0]
    return S.project_device(sc.dt, vel, act, solid, fluid, sc.fluid_levelset, pres, pact)
for _ in range(3): res = step()
print(w, n, flags, "iters", res.iterations, {k: round(v, 3) for k, v in res.stats.items() if k.startswith("ms_")}, "tiles", res.stats["active_tiles"], "/", res.stats["total_tiles"], "rows", res.n_rows, "launches", res.stats["kernel_launches"])
S.profile(True); res = step(); tab = S.profile_table(); S.profile(False)
tot = sum(v[1] for v in tab.values())
for k, (c, ms) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} {c:4d} x {ms / c * 1e3:8.1f} us = {ms:7.3f} ms  {100 * ms / tot:5.1f} %")
print("  total profiled", round(tot, 3))
S.close()
