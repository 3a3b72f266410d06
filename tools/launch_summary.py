"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share, launches, average).
usage: python tools/launch_summary.py launches.csv [--step]   (--step: only the launches of the first whole step,
i.e. from the first normalize_image_kernel up to the second one)"""
import csv, re, sys
rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.reader(lines):
    if r[0] == "ID":
        h = r
        continue
    rows.append((r[h.index("Kernel Name")], r[h.index("Grid Size")], float(r[h.index("Metric Value")]) / 1e3))
if "--step" in sys.argv:
    marks = [i for i, r in enumerate(rows) if "normalize_image_kernel" in r[0]]
    rows = rows[marks[0]: marks[1] if len(marks) > 1 else len(rows)]


def short(n):
    n = re.sub(r"\(.*$", "", n).replace("<unnamed>::", "").replace("void ", "")
    return n[:80]


agg = {}
for n, g, us in rows:
    a = agg.setdefault(short(n), [0.0, 0])
    a[0] += us
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"# total {tot / 1e3:.2f} ms over {len(rows)} launches")
print("# share%  total_us  launches  avg_us  kernel")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{100 * us / tot:6.2f} {us:10.1f} {n:6d} {us / n:8.1f}  {k}")
