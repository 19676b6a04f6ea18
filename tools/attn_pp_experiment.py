"""Next-round experiment: the ping-pong attention kernel (dbg bit12) vs the shipped double-buffered kernel at the UNet's
S = 4096, 8 heads, d = 40 -> 48 shape.  Run under `timeout` (a faulty kernel traps after its 4 s mbarrier watchdog):
  gpurun --timeout 300 -- 'DFB_TEST_PP=1 timeout 120 python -m pytest tests/test_attn_gpu.py -q -k ping_pong; timeout 120 python tools/attn_pp_experiment.py'
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

B = int(os.environ.get("ROWS", "64"))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, 4096, 3 * 384, generator=g).bfloat16().cuda()
o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device="cuda")
for name, flags in (("double-buffered (shipped)", 0), ("ping-pong", 4096), ("double-buffered (shipped)", 0), ("ping-pong", 4096)):
    for _ in range(2):
        ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:28s} {ms:8.3f} ms  {4.0 * B * 8 * 4096 * 4096 * 48 / ms / 1e9:8.1f} TFLOP/s (padded d)", flush=True)
