"""Preconditioner tuning probe (development aid): iterations and best resolve time per configuration.
usage: python tools/gpu_tune.py scene n [residual]"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shiokaze_b200 import MacPressureSolver3, scenes

scene, n = sys.argv[1], int(sys.argv[2])
residual = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
sc = scenes.BENCH_SCENES[scene](n)
print("scene", scene, n, "residual", residual, flush=True)
S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision="mixed", Precond="mg", Residual=residual)
out = S.project_scene(sc)
print("  rows", out["result"].n_rows, flush=True)
for gamma, (pre, post), scale in itertools.product((1, 2), ((1, 1), (2, 2), (3, 3), (2, 1)), (0.5, 0.65)):
    S.configure(MGGamma=gamma, MGPreSweeps=pre, MGPostSweeps=post, MGCoarseScale=scale)
    best, it = None, None
    for _ in range(3):
        r = S.resolve()
        best = r.stats["ms_solve"] if best is None else min(best, r.stats["ms_solve"])
        it = r.iterations
    print(f"  gamma {gamma} sweeps ({pre},{post}) scale {scale}: iters {it} conv {r.converged} reresid {r.reresid:.2e} solve {best:.2f} ms ({best/max(it+1,1):.2f} ms/cycle)", flush=True)
S.close()
