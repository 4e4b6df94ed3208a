#!/bin/bash
# advection (first part) on one B200: parity suite, timing of the bench sub-record, one ncu --set full capture of its two kernels
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_advect.py -q -x --timeout 300 > $O/r02_pytest_gpu_advect.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02_pytest_gpu_advect.log
for w in "dambreak_solid 256" "smoke_plume 256" "dambreak_solid 512"; do
  timeout 300 python tools/advect_time.py $w > $O/r02_advect_${w// /_}.json 2> $O/r02_advect.err || tail -5 $O/r02_advect.err
  python - "$O/r02_advect_${w// /_}.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "kernels", round(r["kernels_ms"], 3), "Mfaces/s", round(d["value"]), "e2e ms", round(d["e2e_ms_per_step"], 2), "frac", round(r["frac"], 3), d.get("cpu_baseline", {}).get("value"))
PY
done
timeout 300 ncu --set full --clock-control none -k regex:"k_advect" -c 4 -o $O/r02_ncu_advect -f python tools/advect_time.py dambreak_solid 256 nocpu 2>&1 | tail -1
ncu -i $O/r02_ncu_advect.ncu-rep --page raw --csv > $O/r02_ncu_advect_raw.csv 2>/dev/null; rm -f $O/r02_ncu_advect.ncu-rep
