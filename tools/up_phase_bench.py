"""Times the four phase GEMMs of a fused Upsample2D (ops.gemm(..., up_phase=(a, b))) at the UNet's three upsample shapes
(256 rows: 8x8x1280, 16x16x1280, 32x32x640 low-resolution inputs) against the literal 3x3 conv on the same low-resolution grid,
with the knobs that could explain a slow shape (CTA pair on / off, GroupNorm partials on / off, tile width).
Usage: python tools/up_phase_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

R = int(os.environ.get("ROWS", "256"))
g = torch.Generator().manual_seed(0)


def timed(f, reps=5):
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for H, C in ((8, 1280), (16, 1280), (32, 640)):
    lo = torch.randn(R, H, H, C, generator=g).bfloat16().cuda()
    w = torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5
    phases = [p.cuda() for p in ops.pack_upsample_phases(w)]
    w3 = ops.pack_conv3x3(w).cuda()
    bias = torch.randn(C, generator=g).cuda()
    out = torch.empty(R, 2 * H, 2 * H, C, dtype=torch.float32, device="cuda")
    out_lo = torch.empty(R, H, H, C, dtype=torch.float32, device="cuda")
    part = torch.empty((4 * R * H * H // 32) * C, dtype=torch.float32, device="cuda")
    flop = 2.0 * R * H * H * C * 4 * C
    print(f"--- low-resolution grid {H}x{H}, {C} channels, {R} rows: one phase = {flop / 1e12:.3f} TFLOP", flush=True)
    for name, kw in (("default", {}), ("no gn partials", {"gn_partial": None}), ("1-CTA kernel", {"cta_group": 1}),
                     ("CTA pair forced", {"cta_group": 2}), ("block_n 128", {"block_n": 128}), ("block_n 160", {"block_n": 160}),
                     ("block_n 192", {"block_n": 192}), ("block_n 256", {"block_n": 256})):
        for a in (0, 1):
            for b in (0, 1):
                args = dict(taps=[ops.upsample_phase_taps(a, b)], conv_geom=(R, H, H), bias=bias, gn_partial=part, up_phase=(a, b))
                args.update(kw)
                try:
                    ms = timed(lambda: ops.gemm([lo], phases[2 * a + b], C, out=out, **args))
                    print(f"phase ({a},{b}) {name:18s} {ms:7.3f} ms  {flop / ms / 1e9:7.1f} TFLOP/s", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"phase ({a},{b}) {name:18s} failed: {e}", flush=True)
                if name != "default":
                    break
            if name != "default":
                break
    # the same taps as a plain conv on the low-resolution grid (no scatter), and the literal 3x3 conv there
    ms = timed(lambda: ops.gemm([lo], phases[0], C, out=out_lo, taps=[ops.upsample_phase_taps(0, 0)], conv_geom=(R, H, H), bias=bias))
    print(f"4-tap conv, no scatter      {ms:7.3f} ms  {flop / ms / 1e9:7.1f} TFLOP/s", flush=True)
    ms = timed(lambda: ops.gemm([lo], w3, C, out=out_lo, taps=[ops.TAPS_3X3], conv_geom=(R, H, H), bias=bias))
    print(f"literal 3x3 conv (9 taps)   {ms:7.3f} ms  {flop * 9 / 4 / ms / 1e9:7.1f} TFLOP/s", flush=True)
