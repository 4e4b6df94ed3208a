# r02v (2 GPUs): whole GPU suite incl. slab + module GPUs=2 tests, weak-scaling line at N=2 (carries parity_vs_1gpu and the strong FLIP 512^3 sub-record), module timing
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02v_pytest.txt; cat $O/r02v_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02v_bench_n2.json 2> $O/r02v_bench_n2.err
tail -2 $O/r02v_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02v_bench_n2.json").read().strip().splitlines()[-1])
print("n2 ms", round(d["ms_per_step"], 3), "e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "what"})
print("parity", d.get("parity_vs_1gpu")); print("strong", d.get("strong")); print("by_kernel", d["roofline"]["by_kernel_ms"])
PY
python tools/module_timing.py dambreak_solid 256 2>&1 | tail -24 > $O/r02v_module_dam256.txt; cat $O/r02v_module_dam256.txt
python tools/module_timing.py dambreak_solid 256 2 2>&1 | grep "project()" > $O/r02v_module_dam256_gpus2.txt; cat $O/r02v_module_dam256_gpus2.txt
