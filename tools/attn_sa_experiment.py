"""Round-2 experiment: attn_fwd_sa_kernel variants against round 1's double-buffered kernel at the UNet's S = 4096, 8 heads,
d = 40 -> 48 shape (ROWS batch rows; 64 by default), plus the in-kernel stamps of the default variant.
Usage: python tools/attn_sa_experiment.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

B = int(os.environ.get("ROWS", "64"))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, 4096, 3 * 384, generator=g).bfloat16().cuda()
qkv.view(B, 4096, 3, 8, 48)[..., 40:] = 0
qkv.view(B, 4096, 3, 8, 48)[:, :, 2, :, 40] = 1.0
o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device="cuda")
tiles_per_sm = B * 8 * 32 * 32 / 148


def run(flags, ones, reps=5):
    f = lambda: ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags, ones_col=ones)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


wsp = torch.empty(ops.attention_ws_elems(B, 8, 4096), dtype=torch.int32, device="cuda")


def run(flags, ones, w=None, reps=5):
    f = lambda: ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags, ones_col=ones, workspace=w)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


variants = [("round-1 double-buffered (dbg bit12)", 4096, None, None, "0"), ("sa + ones", 1 << 13, 40, None, "0"),
            ("sa8 column split", 1 << 13, 40, wsp, "0"), ("sa8 column split + poly 2/16", 2 << 13, 40, wsp, "0"),
            ("sa8 TILE split", 1 << 13, 40, wsp, "1"), ("sa8 TILE split + poly 2/16", 2 << 13, 40, wsp, "1")]
for rnd in range(3):
    for name, flags, ones, w, tiles in variants:
        os.environ["DFB_ATTN_SA8_TILES"] = tiles[0]
        ms = run(flags, ones, w)
        print(f"{name:40s} {ms:8.3f} ms  {4.0 * B * 8 * 4096 * 4096 * 40 / ms / 1e9:8.1f} TFLOP/s (d = 40)  {ms * 1e-3 / tiles_per_sm * 1e9:7.1f} ns per 128x128 tile per SM", flush=True)
print("flags raised:", int(wsp.sum()))
