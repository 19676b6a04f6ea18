"""Cross-attention (short-KV kernel: S_kv = 77 text tokens) at the UNet's four resolutions, 256 rows, 8 heads.
Usage: python tools/xattn_bench.py   (DFB200_LIB=<other build> for an A/B on the same box)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

B, H, SKV = int(os.environ.get("ROWS", "256")), 8, int(os.environ.get("SKV", "77"))
QT = int(os.environ.get("QT", "0"))            # query tiles per CTA (1..15; 0 = the library's choice)
g = torch.Generator().manual_seed(0)
for sq, d in ((4096, 40), (1024, 80), (256, 160), (64, 160)):
    dp = ops.pad16(d)
    q = torch.randn(B, sq, H * dp, generator=g).bfloat16().cuda()
    kv = torch.randn(B, SKV, 2 * H * dp, generator=g).bfloat16().cuda()
    o = torch.empty(B, sq, H * dp, dtype=torch.bfloat16, device="cuda")
    f = lambda: ops.attention(q, kv[..., :H * dp], kv[..., H * dp:], o, heads=H, dp=dp, scale=d ** -0.5, dbg_flags=QT << 8)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (2 * q.numel() * 2 + kv.numel() * 2) / 1e9
    print(f"Sq {sq:5d} d {d:4d}: {ms:7.3f} ms   {gb / ms:7.2f} TB/s of tensor bytes   checksum {float(o.float().abs().sum()):.6e}", flush=True)
