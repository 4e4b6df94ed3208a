"""gridutility3::get_volume of the UNMODIFIED reference build (src/utility/gridutility3.cpp:318-346) on one fixed scene, three calls per thread count:
the evidence behind DESIGN.md section 9 (the value differs from call to call and with the thread count, and is ~8x the liquid volume). CPU only.
usage: python tools/probe_get_volume.py [n]"""
import os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refio
from shiokaze_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sc = scenes.dambreak(n, True)
d = refio.ref_dir("f32")
print(f"dam-break + obstacle {n}^3; fraction of cells with fluid < 0: {float((sc.fluid < 0).mean()):.6f} (the liquid volume of the unit box, to first order)")
with tempfile.TemporaryDirectory() as tmp:
    fin = os.path.join(tmp, "s.bin")
    refio.write_scene(fin, sc)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = d
    for thr in (1, 2, 8, 16):
        out = subprocess.run([os.path.join(d, "ref_driver"), f"in={fin}", f"out={tmp}/o.bin", "RefVolume=3", "RefSkipProject=1", f"Threads={thr}"], env=env, cwd=tmp, capture_output=True, text=True)
        print(f"Threads={thr:2d}  get_volume x3:", "  ".join(re.findall(r"volume=([-0-9.e+]+)", out.stdout)))
