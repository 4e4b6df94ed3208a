"""Quick GPU probe (development aid): solve times per configuration. Usage: python tools/gpu_probe.py [scene] [n]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shiokaze_b200 import MacPressureSolver3, scenes

def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "smoke_plume"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    sc = scenes.BENCH_SCENES[scene](n)
    print("scene", scene, n, flush=True)
    for prec in ("mixed", "fp32", "fp64"):
        for precond, extra in (("mg", dict(MGPreSweeps=1, MGPostSweeps=1)), ("mg", dict(MGPreSweeps=2, MGPostSweeps=2)), ("none", dict(CheckEvery=50))):
            if precond == "none" and n > 256: continue
            S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision=prec, Precond=precond, **extra)
            t = time.time()
            out = S.project_scene(sc)
            t1 = time.time() - t
            r = out["result"]; st = r.stats
            best = None
            for _ in range(3):
                rr = S.resolve()
                best = rr.stats["ms_solve"] if best is None else min(best, rr.stats["ms_solve"])
            it = max(r.iterations, 1)
            print(f"  {prec:5s} {precond:4s} {extra}: rows {r.n_rows} iters {r.iterations} conv {r.converged} reresid {r.reresid:.2e} "
                  f"| project: asm {st['ms_assemble']:.2f} setup {st['ms_setup']:.2f} solve {st['ms_solve']:.2f} upd {st['ms_update']:.2f} h2d {st['ms_h2d']:.1f} d2h {st['ms_d2h']:.1f} (wall {t1*1e3:.0f} ms) "
                  f"| resolve best {best:.2f} ms = {best/it:.3f} ms/iter, {r.n_rows*it/best/1e6:.1f} Mrow-it/ms launches {rr.stats['kernel_launches']}", flush=True)
            S.close()

if __name__ == "__main__":
    main()
