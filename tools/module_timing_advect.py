"""advect_vector through the Shiokaze module (the reference's own host, oracle/ref_driver RefAdvect=vector): the reference's macadvection3 on the host cores
against Advection=b200advection3 on the stock tiledarray3 grids and on Array=b200array3 (dense page-locked grids handed over in place).
usage: python tools/module_timing_advect.py [workload] [n]"""
import dataclasses, importlib.util, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refio
from shiokaze_b200 import scenes
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
w = sys.argv[1] if len(sys.argv) > 1 else "dambreak_solid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = scenes.BENCH_SCENES[w](n)
sc = dataclasses.replace(sc, dt=bench.advect_dt(sc))
base = None
for name, adv, flags in (("macadvection3 (reference, tiledarray3)", None, {}), ("b200advection3, tiledarray3", "b200advection3", {}),
                         ("macadvection3 (reference), Array=b200array3", None, {"Array": "b200array3"}), ("b200advection3, Array=b200array3", "b200advection3", {"Array": "b200array3"})):
    r = refio.run_reference(sc, "f32", flags=flags, advect="vector", advection=adv, repeat=3, threads=os.cpu_count())
    m = re.search(r"project_ms_last=([0-9.]+) project_ms_mean=([0-9.]+)", r.stdout)
    if base is None:
        base = r
    same = all((r.vel[d] == base.vel[d]).all() and (r.vel_active[d] == base.vel_active[d]).all() for d in range(3))
    print(f"{w} {n}^3 {name:46s}: advect_vector() last {m.group(1)} ms, mean of 3 {m.group(2)} ms; equals the reference bit for bit: {same}", flush=True)
    for line in r.stdout.splitlines():
        if "on the GPU" in line:
            print("      ", line.strip()); break
