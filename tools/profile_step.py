"""One warm-up project() and one measured project() of the bench workload — the command ncu wraps.
Usage: python tools/profile_step.py [workload] [n] [precision] [pre] [post] [precond]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from shiokaze_b200 import MacPressureSolver3, scenes

workload = sys.argv[1] if len(sys.argv) > 1 else "smoke_plume"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
precision = sys.argv[3] if len(sys.argv) > 3 else "mixed"
pre = int(sys.argv[4]) if len(sys.argv) > 4 else 2
post = int(sys.argv[5]) if len(sys.argv) > 5 else 2
precond = sys.argv[6] if len(sys.argv) > 6 else "mg"
sc = scenes.BENCH_SCENES[workload](n)
dev = torch.device("cuda", 0)
S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision=precision, Precond=precond, MGPreSweeps=pre, MGPostSweeps=post)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
vel0, act0 = [t(v) for v in sc.vel], [t(a) for a in sc.vel_active]
fluid, solid = t(sc.fluid), (t(sc.solid) if sc.solid is not None else None)
pres = torch.zeros(sc.fluid.shape, dtype=torch.float32, device=dev)
for rep in range(2):
    vel, act = [v.clone() for v in vel0], [a.clone() for a in act0]
    torch.cuda.synchronize()
    res = S.project_device(sc.dt, vel, act, solid, fluid, sc.fluid_levelset, pres, None)
    print("step", rep, res.iterations, res.reresid, {k: round(v, 3) for k, v in res.stats.items() if k.startswith("ms_")}, "launches", res.stats["kernel_launches"], flush=True)
S.close()
