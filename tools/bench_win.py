"""Timing of ms2_window_attention on the Hiera block geometries (batch = ENC_BATCH slices)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
from medsam2_b200 import ops
B = int(os.environ.get("ENC_BATCH", "8"))
cfgs = [("blk0  s1 ws8", 256, 1, 8, 0, 1), ("blk1  s1 ws8 pool", 256, 2, 8, 1, 1), ("blk2  s2 ws4", 128, 2, 4, 0, 1),
        ("blk3  s2 ws4 pool", 128, 4, 4, 1, 1), ("blk4+ s3 ws14", 64, 4, 14, 0, 7), ("blk14 s3 ws14 pool", 64, 8, 14, 1, 1),
        ("blk15 s4 ws7", 32, 8, 7, 0, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
print(f"{'block':20s} {'us':>8s} {'x':>2s} {'GB/s':>7s} {'TF/s':>6s}")
for name, H, heads, ws, qp, mult in cfgs:
    D = 96
    qkv = torch.randn(B, H, H, 3 * heads * D, device="cuda").to(torch.bfloat16)
    bias = torch.randn(3 * heads * D, device="cuda")
    ts = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ops.window_attention(qkv, bias, B, H, H, heads, D, ws, qp); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[1:])[2]
    nw = ((H + ws - 1) // ws) ** 2 * B
    lq = (ws // 2) ** 2 if qp else ws * ws
    fl = 4.0 * nw * heads * lq * ws * ws * D
    byt = qkv.numel() * 2 + o.numel() * 2
    tot += us * mult
    print(f"{name:20s} {us:8.1f} {mult:2d} {byt / us / 1e3:7.0f} {fl / us / 1e6:6.1f}")
print(f"per {B}-slice batch: {tot / 1e3:.2f} ms")
