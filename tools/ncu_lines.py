"""Per-source-line stall samples of an .ncu-rep captured with --import-source on (-lineinfo builds).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur, h, out = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        h = r
    elif len(r) > 6 and h and r[0].isdigit():
        def iv(x):
            try:
                return int(x)
            except ValueError:
                return 0
        n = iv(r[h.index("# Samples")])
        st = sorted(((iv(r[i]), h[i]) for i in range(len(h)) if h[i].startswith("stall_") and "Not" not in h[i]), reverse=True)[:2]
        out.append((n, cur, int(r[0]), r[1].strip()[:90], st))
tot = sum(o[0] for o in out) or 1
for n, f, ln, src, st in sorted(out, reverse=True)[:topn]:
    print(f"{100 * n / tot:5.1f}%  {f}:{ln:<4d} {src:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
