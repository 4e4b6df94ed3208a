"""Per-kernel sums of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>`).
usage: python tools/launch_summary.py X.csv out_summary.csv"""
import csv, re, sys
src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = {}
for r in rows[1:]:
    name = re.sub(r"\(.*$", "", r[ik])
    t = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(r[iu], 1.0)
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in acc.values())
with open(out, "w", newline="") as f:
    w = csv.writer(f); w.writerow(["kernel", "launches", "total_ns", "share"])
    for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, round(t, 1), round(t / tot, 4)])
print(len(rows) - 1, "launches,", len(acc), "kernels,", round(tot / 1e6, 3), "ms under ncu")
