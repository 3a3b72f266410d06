"""Device time of the latency-bound token-side kernels of the mask decoder (CUPTI durations, L2 flushed between
launches so that the weights come from HBM as they do once per slice in the real step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
from medsam2_b200 import ops

bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = []
for name, M, N, K in (("tok.proj 9x256x256", 9, 256, 256), ("tok.proj 9x128x256", 9, 128, 256), ("tok.mlp1 9x2048x256", 9, 2048, 256),
                      ("tok.mlp2 9x256x2048", 9, 256, 2048), ("ptr.proj 1x256x256", 1, 256, 256)):
    a = torch.randn(M, K, device="cuda").to(bf); w = torch.randn(N, K, device="cuda").to(bf); b = torch.randn(N, device="cuda")
    cases.append((name, lambda a=a, w=w, b=b: ops.gemm(a, w, b, out_dtype=bf)))
x = torch.randn(9, 256, device="cuda"); g = torch.ones(256, device="cuda"); z = torch.zeros(256, device="cuda")
cases.append(("layernorm 9x256", lambda: ops.layernorm(x, g, z, 1e-5)))
q = torch.randn(1, 9, 128, device="cuda").to(bf); k = torch.randn(1, 4096, 128, device="cuda").to(bf)
cases.append(("attn tokens->image 9q x 4096k, 8 heads x 16", lambda: ops.attention(q, k, k, 8)))
cases.append(("attn image->tokens 4096q x 9k, 8 heads x 16", lambda: ops.attention(k, q, q, 8)))
q2 = torch.randn(1, 9, 256, device="cuda").to(bf)
cases.append(("attn tokens self 9x9, 8 heads x 32", lambda: ops.attention(q2, q2, q2, 8)))
for name, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for cold in (True, False):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(8):
                if cold:
                    flush.zero_()
                fn()
            torch.cuda.synchronize()
        rows = [e for e in prof.key_averages() if "FillFunctor" not in e.key and "Memset" not in e.key]
        tot = sum(e.device_time_total for e in rows) / 8
        print(f"{name:48s} {'cold L2' if cold else 'warm L2'}  {tot:7.1f} us  ({', '.join(f'{e.key[:28]} {e.device_time_total / e.count:.1f}' for e in rows)})")
