"""Turn .ncu-rep captures (ncu --set full) into the per-kernel summary CSV kept under profiles/ and, optionally, the
per-row DRAM traffic table bench.py reads (profiles/traffic.json). Runs here (no GPU needed): ncu -i ... --page raw --csv.
usage: python tools/ncu_summary.py out.csv rows a.ncu-rep [b.ncu-rep ...]      (rows = unknowns of the captured run, 0: no traffic table)"""
import csv, json, os, re, subprocess, sys

COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

out, rows_n, reps = sys.argv[1], float(sys.argv[2]), sys.argv[3:]
header, units, lines, traffic = None, None, [], {}
for rep in reps:
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    header = ["Kernel Name"] + COLS
    units = [""] + [rows[1][idx[c]] if c in idx else "" for c in COLS]
    for r in rows[2:]:
        name = re.sub(r"\bshkz::", "", r[idx["Kernel Name"]])
        lines.append([name] + [r[idx[c]] if c in idx else "" for c in COLS])
        if rows_n > 0:
            rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[rows[1][idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[rows[1][idx["dram__bytes_write.sum"]]]
            traffic.setdefault(name, []).append((rd + wr) / rows_n)
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(header); w.writerow(units); w.writerows(lines)
print("wrote", out, len(lines), "kernels")
for k, v in traffic.items():
    print("  %-110s %.2f B/row" % (k[:110], sum(v) / len(v)))
if rows_n > 0:
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out.replace(".csv", "_traffic.json"), "w"), indent=1)
