"""Summarise tools/ubench/attn_trace output: tile period and per-event offsets relative to the S tile becoming ready."""
import sys
import numpy as np
path = sys.argv[1]
names = open(path).readline().split()
a = np.array([[int(x) for x in l.split()] for l in open(path) if l.strip() and l.split()[0].isdigit()])
t = a[:, 1:]
print("tile period:", np.diff(t[:, 3])[4:24].mean())
rel = t[4:28] - t[4:28, 5:6]
for i in range(16):
    print(f"{names[i + 1]:>14s} {rel[:, i].mean():9.0f}")
