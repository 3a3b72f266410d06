"""Summarise an .ncu-rep: headline metrics + the hottest SASS instructions with their dominant stall reason.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
for row in r[2:]:
    print("== kernel:", row[r[0].index("Kernel Name")][:90] if "Kernel Name" in r[0] else "")
    for k in want:
        if k in r[0]:
            i = r[0].index(k)
            print(f"  {k} = {row[i]} {r[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rr = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = rr[0]
iS, iN, iI = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
rows = []
for x in rr[1:]:
    if len(x) != len(h) or x[0] == "Address":
        continue
    try:
        n = int(x[iN] or 0)
    except ValueError:
        continue
    st = sorted(((int(x[i] or 0), h[i]) for i in stalls), reverse=True)[:2]
    rows.append((n, int(x[iI] or 0), x[iS].strip(), st))
tot = sum(a for a, *_ in rows) or 1
toti = sum(b for _, b, *_ in rows) or 1
print(f"total samples {tot}, warp instructions {toti}")
agg = {}
for n, _, _, st in rows:
    for c, k in st[:1]:
        agg[k] = agg.get(k, 0) + n
print("samples by dominant stall of the instruction:", sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
for n, ins, s, st in sorted(rows, key=lambda t: -t[0])[:topn]:
    print(f"{100 * n / tot:5.1f}%  {100 * ins / toti:4.1f}%i  {s[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
