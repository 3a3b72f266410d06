"""Per-shape timing of ms2_gemm on the Hiera / memory-attention shapes (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
from medsam2_b200 import ops

B = int(os.environ.get("ENC_BATCH", "8"))
T1, T2, T3, T4 = 65536 * B, 16384 * B, 4096 * B, 1024 * B
shapes = [  # (name, M, N, K, out_dtype, act, residual)
    ("s1.qkv", T1, 288, 96, "bf16", 0, 0), ("s1.proj", T1, 96, 96, "f32", 0, 1), ("s1.fc1", T1, 384, 96, "bf16", 1, 0),
    ("s1.fc2", T1, 96, 384, "f32", 0, 1), ("b1.qkv", T1, 576, 96, "bf16", 0, 0), ("b1.proj", T1, 192, 96, "f32", 0, 0),
    ("s2.qkv", T2, 576, 192, "bf16", 0, 0), ("s2.proj", T2, 192, 192, "f32", 0, 1), ("s2.fc1", T2, 768, 192, "bf16", 1, 0),
    ("s2.fc2", T2, 192, 768, "f32", 0, 1), ("s3.qkv", T3, 1152, 384, "bf16", 0, 0), ("s3.proj", T3, 384, 384, "f32", 0, 1),
    ("s3.fc1", T3, 1536, 384, "bf16", 1, 0), ("s3.fc2", T3, 384, 1536, "f32", 0, 1), ("s4.qkv", T4, 2304, 768, "bf16", 0, 0),
    ("s4.fc1", T4, 3072, 768, "bf16", 1, 0), ("s4.fc2", T4, 768, 3072, "f32", 0, 1),
    ("neck0", T1, 256, 96, "f32", 0, 0), ("conv_s0", T1, 32, 256, "f32", 0, 0), ("neck2", T3, 256, 384, "f32", 0, 0),
    ("mem.kproj", 209120, 256, 64, "bf16", 0, 0), ("mem.qkv", 4096, 768, 256, "bf16", 0, 0), ("mem.ffn1", 4096, 2048, 256, "bf16", 2, 0),
    ("mem.ffn2", 4096, 256, 2048, "f32", 0, 1), ("dec.convT1", 4096, 256, 256, "f32", 0, 0), ("dec.convT2", 16384, 128, 64, "f32", 0, 0),
    ("big", 8192, 8192, 8192, "bf16", 0, 0),
]
only = os.environ.get("ONLY")
if only:
    shapes = [s for s in shapes if s[0] in only.split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"{'name':10s} {'M':>8s} {'N':>5s} {'K':>5s}  {'us':>8s} {'TF/s':>7s} {'GB/s':>7s}")
for name, M, N, K, od, act, res in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    odt = torch.bfloat16 if od == "bf16" else torch.float32
    out = torch.empty(M, N, device="cuda", dtype=odt)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, bias, out_dtype=odt, act=act, residual=r, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[1:])[len(ts[1:]) // 2]
    byt = M * K * 2 + N * K * 2 + M * N * (2 if od == "bf16" else 4) + (M * N * 4 if res else 0)
    print(f"{name:10s} {M:8d} {N:5d} {K:5d}  {us:8.1f} {2.0 * M * N * K / us / 1e6:7.1f} {byt / us / 1e3:7.0f}")
    del a, w, r, out
