import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shiokaze_b200 import MacPressureSolver3, scenes
for kind, n, pre, post in (("dambreak_solid", 40, 2, 2), ("smoke", 40, 1, 1), ("smoke", 64, 2, 2)):
    sc = scenes.dambreak(n, True) if kind == "dambreak_solid" else scenes.smoke_plume(n)
    for omega in (1.0, 1.15):
        S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision="mixed", Precond="mg", MGPreSweeps=pre, MGPostSweeps=post, MaxIterations=1, MGOmega=omega)
        S.project_scene(sc)
        fused = S.debug_vcycle(legacy=0); scalar = S.debug_vcycle(legacy=2); quad = S.debug_vcycle(legacy=3); fused2 = S.debug_vcycle(legacy=0)
        for name, a in (("fused", fused), ("quad", quad), ("fused-again", fused2)):
            diff = a != scalar
            print(kind, n, "omega", omega, name, "differs in", int(diff.sum()), "cells, max abs", float(np.abs(a - scalar).max()), "scale", float(np.abs(scalar).max()))
            if diff.any():
                k, j, i = np.nonzero(diff)
                print("    k range", k.min(), k.max(), "j range", j.min(), j.max(), "i range", i.min(), i.max(), "first", list(zip(k[:6], j[:6], i[:6])))
        S.close()
