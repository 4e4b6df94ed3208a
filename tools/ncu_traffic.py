"""profiles/traffic.json from an `ncu --set full` capture of one project() (runs here, no GPU): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per
unknown and launch of the level-0 solve kernels, keyed by the profiler tags bench.py uses, stamped with the hash of the kernel sources they were captured from
(bench.py quotes the table only while that hash matches).
usage: python tools/ncu_traffic.py <capture.ncu-rep | capture_raw.csv> <n_rows of the captured workload> <set name> <workload text> [summary.csv]"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)

rep, rows_n, set_name, workload = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4]
summary = sys.argv[5] if len(sys.argv) > 5 else None
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
# (a capture of a whole solve at 512^3 is > 64 MB, more than gpurun brings back: the box reduces it with `ncu -i X.ncu-rep --page raw --csv > X_raw.csv`)
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]


def val(r, key):
    return float(r[idx[key]].replace(",", "")) * UNIT.get(units[idx[key]], 1.0)


def tag_of(name):
    m = re.search(r"k_sweep_tma<(\d+), (\d+), (\d+), (\d+), (\d+)>", name)
    if m:
        return "sweep@0" + ("z" if m.group(2) == "1" else "") + ("p" if m.group(3) == "1" else "") + ("d" if m.group(4) == "1" else "")
    for key, tag in (("k_xpay_spmv_tma", "xpay_spmv_dot"), ("k_spmv_dot4", "spmv_dot"), ("k_axpy2_norm", "axpy2_norm"), ("k_residual_restrict", "residual_restrict@0"), ("k_xpay<", "xpay")):
        if key in name:
            return tag
    return None


best = {}
lines = []
for r in rows[2:]:
    name = re.sub(r"\bshkz::", "", r[idx["Kernel Name"]])
    tag = tag_of(name)
    if not tag:
        continue
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    lines.append([tag, name[:90], r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]], "%.1f" % (b / 1e6), "%.2f" % (b / rows_n)])
    if tag not in best or b > best[tag][0]:   # the level-0 launch of a kernel is the one that moves the most bytes
        best[tag] = (b, name, t)
out = {"_meta": {"set": set_name, "workload": workload, "rows": rows_n, "kernel_source_hash": bench.kernel_source_hash(),
                 "what": "dram__bytes_read.sum + dram__bytes_write.sum of the level-0 launch of each kernel / unknown rows of the captured workload (ncu --set full --clock-control none)"}}
for tag, (b, name, t) in sorted(best.items()):
    out[tag] = {"bytes_per_row": b / rows_n, "kernel": name[:120]}
    print("%-22s %8.2f B/row   %s" % (tag, b / rows_n, name[:80]))
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
if summary:
    with open(summary, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["tag", "kernel", "duration", "unit", "dram_MB", "dram_B_per_row"])
        w.writerows(lines)
