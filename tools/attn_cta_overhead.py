"""Per-CTA fixed cost of the long self-attention kernel: time Sq = 4096 against Skv = 1024 / 2048 / 4096 / 8192 keys (same CTA count,
16 / 32 / 64 / 128 K/V tiles per CTA) and fit t = CTAs / 296 * (T_fix + tiles * T_tile)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops
B = int(os.environ.get("ROWS", "32"))
g = torch.Generator().manual_seed(0)
res = {}
for skv in (1024, 2048, 4096, 8192):
    q = torch.randn(B, 4096, 384, generator=g).bfloat16().cuda()
    kv = torch.randn(B, skv, 768, generator=g).bfloat16().cuda()
    for t in (q.view(B, 4096, 8, 48), kv.view(B, skv, 16, 48)):
        t[..., 40:] = 0
    kv.view(B, skv, 2, 8, 48)[:, :, 1, :, 40] = 1.0
    o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device="cuda")
    wsp = torch.empty(ops.attention_ws_elems(B, 8, 4096), dtype=torch.int32, device="cuda")
    for name, w in (("sa8", wsp), ("sa", None)):
        f = lambda: ops.attention(q, kv[..., :384], kv[..., 384:], o, heads=8, dp=48, scale=40 ** -0.5, ones_col=40, workspace=w)
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record(); torch.cuda.synchronize()
        res[(name, skv)] = e0.elapsed_time(e1) / 5
        print(f"{name:4s} Skv {skv:5d}: {res[(name, skv)]:8.3f} ms", flush=True)
ctas = B * 8 * 32
for name in ("sa8", "sa"):
    for a, b in ((1024, 2048), (2048, 4096), (4096, 8192)):
        ta, tb = res[(name, a)], res[(name, b)]
        per_tile = (tb - ta) / ((b - a) / 64) / (ctas / 296) * 1e6        # ns per 64-wide tile per CTA slot
        fix = (ta / (ctas / 296) * 1e6) - (a / 64) * per_tile
        print(f"{name}: from Skv {a} -> {b}: T_tile = {per_tile:7.1f} ns, T_fix = {fix:8.1f} ns per CTA ({fix / (fix + 64 * per_tile) * 100:4.1f} % of a 64-tile CTA)")
