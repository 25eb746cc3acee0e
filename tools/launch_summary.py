"""Per-kernel totals of ONE time step from an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches_*.csv).
usage: python tools/launch_summary.py profiles/r1g_launches_C4.csv   (takes the last complete step: k_reset .. next k_reset)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
names = []
for r in rows[hi + 1:]:
    if len(r) > vi:
        try:
            names.append((r[ki].split("(")[0].replace("symb200::", ""), float(r[vi].replace(",", ""))))
        except ValueError:
            pass
first = "k_reset" if any(n == "k_reset" for n, _ in names) else "k_prepare"
idx = [i for i, (n, _) in enumerate(names) if n == first]
step = names[idx[-2]:idx[-1]]
agg = collections.OrderedDict()
for n, v in step:
    e = agg.setdefault(n, [0, 0.0])
    e[0] += 1
    e[1] += v
tot = sum(v for _, v in step)
print("| kernel | launches | µs | share |\n|---|---|---|---|")
for n, (c, v) in agg.items():
    print(f"| {n} | {c} | {v / 1000:.1f} | {100 * v / tot:.1f} % |")
print(f"| total | {len(step)} | {tot / 1000:.1f} | |")
