"""Timing experiments for the S=4096 self-attention kernel (tuning hooks of dfb_attn_params.dbg_flags)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops
B = int(os.environ.get("B", "32"))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, 4096, 1152, generator=g).bfloat16().cuda()
o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device="cuda")
def run(flags, bkv=0, reps=5):
    f = lambda: ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, block_kv=bkv, dbg_flags=flags)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tiles = B * 8 * 32 * 32 / 148
for name, flags, bkv in [("split-KV, 8 softmax warps (experiment)", 32, 64), ("double-buffer KV=64 (smem P)", 16, 64), ("double-buffer KV=64, P in TMEM (default)", 0, 64),
                         ("double-buffer KV=128 (smem P, 1 CTA/SM)", 16, 128), ("double-buffer KV=128, P in TMEM (1 CTA/SM)", 0, 128)]:
    ms = run(flags, bkv)
    print(f"{name:44s} {ms:8.3f} ms -> {ms * 1e-3 * 1.9e9 / tiles:7.0f} cycles per 128x128 tile per SM (@1.9GHz)")
tl = torch.zeros(4096, dtype=torch.int64, device="cuda")
ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5, block_kv=64, dbg_timeline=tl)
torch.cuda.synchronize()
t = tl.cpu().tolist()
base = t[8 * 8]
print("double-buffer KV=64 P-in-TMEM, 2 CTA/SM: softmax thread of CTA(0,0,0): tile | begin  S ready  math done  arrived")
for j in range(8, 20):
    r = [t[j * 8 + k] - base for k in range(4)]
    print(f"   {j:2d} | " + "  ".join(f"{v:7d}" for v in r) + f"   (wait {r[1]-r[0]}, math {r[2]-r[1]}, fence+arrive {r[3]-r[2]})")
# ---- cross-attention (S_kv = 77): generic single-tile kernel vs the short-KV kernel that streams query tiles ----
Bx = int(os.environ.get("BX", "128"))
q = torch.randn(Bx, 4096, 384, generator=g).bfloat16().cuda()
kv = torch.randn(Bx, 77, 768, generator=g).bfloat16().cuda()
ox = torch.empty(Bx, 4096, 384, dtype=torch.bfloat16, device="cuda")
def runx(flags, reps=5):
    f = lambda: ops.attention(q, kv[..., :384], kv[..., 384:], ox, heads=8, dp=48, scale=40 ** -0.5, dbg_flags=flags)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
gb = (2 * q.numel() * 2 + kv.numel() * 2) / 1e9
for name, flags in [("generic single-tile kernel (one CTA per 128 queries)", 64), ("short-KV, 1 tile/CTA", 1 << 8), ("short-KV, 2 tiles/CTA", 2 << 8),
                    ("short-KV, 4 tiles/CTA", 4 << 8), ("short-KV, 8 tiles/CTA (default)", 0), ("short-KV, 15 tiles/CTA", 15 << 8)]:
    ms = runx(flags)
    print(f"cross-attn B={Bx} {name:56s} {ms:8.3f} ms  ({gb / ms * 1e3 / 1e3:6.2f} TB/s of Q+O+KV bytes)")
