"""Smoother tuning probe (development aid): relaxation factor x sweep counts -> iterations and best resolve time.
usage: python tools/gpu_tune2.py scene:n [scene:n ...]"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shiokaze_b200 import MacPressureSolver3, scenes

for spec in sys.argv[1:]:
    scene, n = spec.split(":"); n = int(n)
    sc = scenes.BENCH_SCENES[scene](n)
    S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision="mixed", Precond="mg", Residual=1e-4, MaxIterations=60)
    out = S.project_scene(sc)
    print("scene", scene, n, "rows", out["result"].n_rows, flush=True)
    for (pre, post), omega in itertools.product(((2, 2), (3, 3)), (1.0, 1.1, 1.15, 1.2)):
        S.configure(MGPreSweeps=pre, MGPostSweeps=post, MGOmega=omega)
        best, it = None, None
        for _ in range(2):
            r = S.resolve()
            best = r.stats["ms_solve"] if best is None else min(best, r.stats["ms_solve"])
            it = r.iterations
        print(f"  sweeps ({pre},{post}) omega {omega}: iters {it} conv {r.converged} reresid {r.reresid:.2e} solve {best:.2f} ms", flush=True)
    S.close()
