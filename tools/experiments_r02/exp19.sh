# r02z (2 GPUs): fused slab prolongation with the colour rule
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > $O/r02z_pytest_multi.txt; tail -5 $O/r02z_pytest_multi.txt
for rep in 1 2; do
python tools/module_timing.py dambreak_solid 256 2 > $O/r02z_module_dam256_gpus2_$rep.txt 2>&1; grep -E "project\(\)|failed|Error|error" $O/r02z_module_dam256_gpus2_$rep.txt | head -5
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02z_bench_n2.json 2> $O/r02z_bench_n2.err
tail -2 $O/r02z_bench_n2.err
python - <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02z_bench_n2.json").read().strip().splitlines()[-1])
print("n2 ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "iters", d["solve"]["iterations"], "parity", d.get("parity_vs_1gpu"))
print("by_kernel", d["roofline"]["by_kernel_ms"]); print("strong", d.get("strong", {}).get("ms_per_step"))
PY
