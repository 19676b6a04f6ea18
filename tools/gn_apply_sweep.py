"""GroupNorm-apply launch shape sweep (DFB_GN_CTAS_PER_SM must be set per PROCESS: read once).  Times the UNet's GN shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops
shapes = [(256, 64, 64, 320, 10), (256, 32, 32, 640, 11), (256, 16, 16, 1280, 9), (256, 64, 64, 640, 2), (256, 32, 32, 1280, 4), (256, 8, 8, 1280, 6)]
tot = 0.0
for B, H, W, C, cnt in shapes:
    x = torch.randn(B, H, W, C, device="cuda")
    part = torch.randn(*ops.gn_partial_shape(B * H * W, C), device="cuda").abs()
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    ws = torch.empty(ops.groupnorm_ws_floats(B, 32), device="cuda")
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    f = lambda: ops.groupnorm(x, None, g, b, groups=32, eps=1e-5, silu=True, stats_ws=ws, out=out, partials=(part, None))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = x.numel() * 6 / 1e9
    tot += ms * cnt
    print(f"  ({B},{H},{W},{C}) x{cnt}: {ms:.4f} ms  {gb / ms:.2f} TB/s")
print(f"CTAS_PER_SM={os.environ.get('DFB_GN_CTAS_PER_SM', '16')}: weighted total {tot:.3f} ms")
