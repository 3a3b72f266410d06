"""Timing of ms2_attention on the memory-attention / Hiera-global shapes (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
from medsam2_b200 import ops

shapes = [("mem.cross.100k", 1, 1, 4096, 100 * 1024 + 64, 256), ("mem.cross.209k", 1, 1, 4096, 209120, 256),
          ("mem.self", 1, 1, 4096, 4096, 256), ("hiera.global.b8", 8, 4, 4096, 4096, 96), ("hiera.global.b1", 1, 4, 4096, 4096, 96)]
shapes.append(("mem.cross.dv.209k", 1, 1, 4096, 209120, 256))
only = os.environ.get("ONLY")
if only:
    shapes = [s for s in shapes if s[0] in only.split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"{'name':18s} {'B':>2s} {'H':>2s} {'Lq':>6s} {'Lk':>7s} {'D':>4s} {'us':>9s} {'TF/s':>7s}")
for name, B, H, Lq, Lk, D in shapes:
    q = torch.randn(B, Lq, H * D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Lk, H * D, device="cuda").to(torch.bfloat16)
    dv = ".dv." in name
    v = torch.randn(B, Lk, 64 if dv else H * D, device="cuda").to(torch.bfloat16)
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.attention_dv(q, k, v) if dv else ops.attention(q, k, v, H)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[1:])[len(ts[1:]) // 2]
    print(f"{name:18s} {B:2d} {H:2d} {Lq:6d} {Lk:7d} {D:4d} {us:9.1f} {(2.0 * B * Lq * Lk * (D + 64) if dv else 4.0 * B * H * Lq * Lk * D) / us / 1e6:7.1f}")
