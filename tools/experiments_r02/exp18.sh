# r02y (2 GPUs): fused slab prolongation + fused slab product, eager module loading in the Shiokaze modules; module GPUs=2 repeated to catch the rare stall
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > $O/r02y_pytest_multi.txt; cat $O/r02y_pytest_multi.txt
for rep in 1 2 3; do
python tools/module_timing.py dambreak_solid 256 2 > $O/r02y_module_dam256_gpus2_$rep.txt 2>&1; grep -E "project\(\)|failed|Error|error" $O/r02y_module_dam256_gpus2_$rep.txt | head -5
done
python tools/module_timing.py smoke_plume 192 2 2>&1 | grep -E "project\(\)|failed|rror" | head -5
for tag in fused noprolong; do
if [ $tag = noprolong ]; then export SHKZ_B200_NO_SLAB_PROLONG=1; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-sub-records > $O/r02y_bench_n2_$tag.json 2> $O/r02y_bench_n2_$tag.err
tail -2 $O/r02y_bench_n2_$tag.err
python - $tag <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02y_bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "n2 ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "iters", d["solve"]["iterations"], "parity ok", d.get("parity_vs_1gpu", {}).get("ok"))
print("by_kernel", d["roofline"]["by_kernel_ms"])
PY
done
