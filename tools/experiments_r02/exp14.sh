# r02u: sparse host copies (fixed), new tail, hybrid tile walk
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02u_pytest.txt; cat $O/r02u_pytest.txt
for w in "dambreak_solid 512" "flip_splash 512" "dambreak 64" "smoke_plume 256" "smoke_plume 512"; do
f=$O/r02u_table_$(echo $w | tr ' ' '_')
python tools/gpu_profile_table.py $w > ${f}_hybrid.txt 2>&1
SHKZ_B200_NO_HYBRID=1 python tools/gpu_profile_table.py $w > ${f}_strided.txt 2>&1
head -1 ${f}_hybrid.txt; head -1 ${f}_strided.txt
done
for fl in "MGMinSize=8" "MGMinSize=16" "MGCoarseSweeps=4"; do
python tools/gpu_profile_table.py dambreak_solid 512 $fl 2>&1 | head -1
python tools/gpu_profile_table.py dambreak 64 $fl 2>&1 | head -1
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r02u_bench.json 2> $O/r02u_bench.err
python bench.py --workload flip_splash --n 512 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records > $O/r02u_bench_flip.json 2> $O/r02u_bench_flip.err
python - <<'PY'
import json
for f in ("r02u_bench", "r02u_bench_flip"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 3), "e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "what"}, "solve_whole", d["roofline"]["solve_whole"]["frac"], "sub", {k: (v.get("ms_per_step"), v.get("e2e_ms_per_step")) for k, v in d.get("sub_records", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
du -sh $O
