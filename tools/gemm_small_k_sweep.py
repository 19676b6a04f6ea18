"""Tile-width / CTA-pair sweep of the transformer blocks' small-K GEMMs at the 256-row step's shapes (the launches DESIGN §6 lists as
shared-memory-port- or HBM-bound): default choice of dfb_gemm against forced block_n x cta_group.
Usage: python tools/gemm_small_k_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g)
dev = "cuda"


def timed(f, reps=4):
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [("GEGLU proj 64x64", 1048576, 2560, 320, "geglu"), ("q|k|v 64x64", 1048576, 1152, 320, "bf16"),
         ("attn out-proj 64x64 (+res)", 1048576, 320, 384, "res"), ("to_out / proj 64x64 (+res)", 1048576, 320, 320, "res"),
         ("ff.out 64x64 (+res)", 1048576, 320, 1280, "res"), ("N=640 K=640 32x32 (+res)", 262144, 640, 640, "res"),
         ("to_q 32x32", 262144, 640, 640, "bf16"), ("GEGLU proj 32x32", 262144, 5120, 640, "geglu"), ("q|k|v 32x32", 262144, 1920, 640, "bf16")]
for name, M, N, K, kind in cases:
    a = rn(M // 16, K).bfloat16().to(dev).repeat(16, 1)
    if kind == "geglu":
        w, b = ops.pack_geglu(rn(N, K) * K ** -0.5, rn(N) * 0.1)
        w, b = w.to(dev), b.to(dev)
        out = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
        kw = dict(bias=b, geglu=True)
    else:
        w = ops.pack_linear(rn(N, K) * K ** -0.5).to(dev)
        b = rn(N).to(dev)
        if kind == "res":
            out = torch.zeros(M, N, dtype=torch.float32, device=dev)
            kw = dict(bias=b, residual=out)
        else:
            out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            kw = dict(bias=b)
    flop = 2.0 * M * N * K
    base = timed(lambda: ops.gemm([a], w, N, out=out, **kw))
    print(f"--- {name}: M={M} N={N} K={K}: default {base:.3f} ms  {flop / base / 1e9:7.1f} TFLOP/s", flush=True)
    for cg in (2, 1):
        row = []
        for bn in (96, 128, 160, 192, 224, 256):
            if kind == "geglu" and bn % 32:
                continue
            try:
                ms = timed(lambda: ops.gemm([a], w, N, out=out, block_n=bn, cta_group=cg, **kw))
                row.append(f"{bn}: {ms:.3f}{'*' if ms < 0.98 * base else ''}")
            except Exception as e:  # noqa: BLE001
                row.append(f"{bn}: -")
        print(f"    cta_group {cg}:  " + "   ".join(row), flush=True)
