"""project() through the Shiokaze module (the reference's own host, oracle/ref_driver) with the default tiledarray3 grids and with Array=b200array3.
usage: python tools/module_timing.py [workload] [n] [GPUs]"""
import os, sys, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refio
from shiokaze_b200 import scenes
w = sys.argv[1] if len(sys.argv) > 1 else "smoke_plume"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = scenes.BENCH_SCENES[w](n)
for name, flags in (("tiledarray3 (default)", {}), ("lineararray3", {"Array": "lineararray3"}), ("b200array3", {"Array": "b200array3"})):
    f = dict(flags)
    if gpus > 1:
        f["GPUs"] = gpus
    r = refio.run_reference(sc, "f32", projection="b200pressure3", flags=f, repeat=4, threads=os.cpu_count())
    m = re.search(r"project_ms_last=([0-9.]+) project_ms_mean=([0-9.]+)", r.stdout)
    ph = {k: v for k, v in re.findall(r"(Gathering dense buffers|Solving on the GPU|Scattering results)\.\.\.(?:Done\. Took [^\n]*?)?([0-9.]+ (?:msec|sec))", r.stdout)[-3:]}
    print(f"{w} {n}^3 GPUs={gpus} {name:24s}: project() last {m.group(1)} ms, mean of 4 {m.group(2)} ms, iterations {r.iterations}", flush=True)
    for line in r.stdout.splitlines()[-14:]:
        if any(k in line for k in ("Gathering", "Solving on", "Scattering", "Projection done")):
            print("      ", line.strip())
