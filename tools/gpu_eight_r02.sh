#!/bin/bash
# 8 GPUs: slab suite (worlds 2, 4, 8), weak-scaling lines at N = 8 and 4 (they carry parity_vs_1gpu and the strong sub-records: liquid box 1024^3 on 8, FLIP 512^3 on 4)
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6 > $O/r02_pytest_multi_8gpu.log; cat $O/r02_pytest_multi_8gpu.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > $O/r02_scale_n${n}_dambreak512.json 2> $O/r02_scale_n${n}.err
tail -2 $O/r02_scale_n${n}.err
python - $n <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02_scale_n%s_dambreak512.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("N", sys.argv[1], "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "iters", d["solve"]["iterations"], "parity ok", d.get("parity_vs_1gpu", {}).get("ok"))
print("strong", {k: d.get("strong", {}).get(k) for k in ("grid", "ms_per_step", "value")}, (d.get("strong", {}).get("solve") or {}).get("iterations"))
print("by_kernel", d["roofline"]["by_kernel_ms"])
PY
done
