"""Diagnostic: is forward_nhwc invariant to the batch it runs a row in?  Rows [0, b) of a B = 48 batch vs the same rows run
alone (b = 16, 20, 32), per block tap; tiny UNet, shared prefix off."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_unet_gpu import _mk
from difashion_b200 import ops

oracle, unet = _mk("tiny")
cfg = oracle.cfg
g = torch.Generator().manual_seed(7)
B = 48
x = torch.randn(B, cfg.sample_size, cfg.sample_size, cfg.in_channels, generator=g).bfloat16().cuda()
ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g).cuda()
t = torch.full((B,), 981.0, device="cuda")


def run(b):
    ws = unet.workspace(("diag", b), torch.device("cuda"))
    c, kv = unet.set_context(ctx[:b].contiguous())
    taps = {}
    out = unet.forward_nhwc(x[:b].contiguous(), t[:b], c, kv, ws, taps=taps).clone()
    torch.cuda.synchronize()
    taps["eps"] = out
    return {k: v.clone() for k, v in taps.items()}


full = run(B)
for b in (16, 20, 32, 12, 8):
    part = run(b)
    msg = []
    for k in full:
        d = (full[k][:b].float() - part[k].float()).abs()
        rows = (d.reshape(b, -1).amax(1) > 0).nonzero().flatten().tolist()
        msg.append(f"{k}:{float(d.max()):.1e}" + (f"@rows{rows[:6]}" if rows else ""))
    print(f"b={b}: " + " ".join(msg), flush=True)
