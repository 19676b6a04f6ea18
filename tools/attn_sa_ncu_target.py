"""One launch of each attention variant at S = 4096, d = 40 -> 48 for `ncu --set full --import-source on` (source-level stalls).
ROWS batch rows (default 8: 3 waves of CTAs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops
B = int(os.environ.get("ROWS", "8"))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, 4096, 3 * 384, generator=g).bfloat16().cuda()
qkv.view(B, 4096, 3, 8, 48)[..., 40:] = 0
qkv.view(B, 4096, 3, 8, 48)[:, :, 2, :, 40] = 1.0
o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device="cuda")
wsp = torch.empty(ops.attention_ws_elems(B, 8, 4096), dtype=torch.int32, device="cuda")
for flags, ones, w in ((1 << 13, 40, None), (1 << 13, 40, None), (0, 40, wsp)):
    ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags, ones_col=ones, workspace=w)
torch.cuda.synchronize()
