"""advect_vector sub-record of bench.py alone (device-resident timing, end to end, roofline): python tools/advect_time.py <workload> <n> [nocpu]"""
import importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
import torch
workload, n = sys.argv[1], int(sys.argv[2])
sc = bench.build_scene(workload, n)
torch.cuda.set_device(0)
rec = bench.advect_sub_record(torch, torch.device("cuda", 0), 0, sc, workload, n, 5, len(sys.argv) < 4)
print(json.dumps(rec))
