"""Bridge cost of the Shiokaze module: the reference host (oracle/ref_driver) runs one project() with Projection=b200pressure3 and prints the module's own timers."""
import sys, os, re, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refio
from shiokaze_b200 import scenes
for name, n in (("smoke", 128), ("dambreak_solid", 128), ("smoke", 192)):
    sc = scenes.smoke_plume(n) if name == "smoke" else scenes.dambreak(n, True)
    t = time.time()
    r = refio.run_reference(sc, "f32", projection="b200pressure3", repeat=2)
    wall = time.time() - t
    lines = [l for l in r.stdout.splitlines() if re.search(r"Gathering|Solving on the GPU|Scattering|Projection done|Took .* iterations", l)]
    print(name, n, "wall %.1f s (incl. scene file I/O)" % wall)
    for l in lines[-8:]:
        print("   ", l.strip()[:160])
