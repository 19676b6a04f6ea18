"""Launch each hot kernel once at a representative shape (after a warm-up launch) so that
`ncu --set full -k regex:...` captures exactly these.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|attn_fwd|groupnorm|layernorm|cfg_step' \
      -o gpurun_out/prof python tools/ncu_targets.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difashion_b200 import ops  # noqa: E402

dev = "cuda"
R = int(os.environ.get("ROWS", "64"))          # UNet rows (batch)
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g)

which = set(os.environ.get("WHICH", "geglu,outproj,qkv,conv,attn,xattn,gn,ln,cfg").split(","))
for rep in range(2):                           # launch 0 = warm-up, launch 1 = the one to look at
    if "geglu" in which:                       # GEGLU proj at 64x64: M = R*4096, N = 2560, K = 320
        M, C = R * 4096, 320
        a = rn(M, C).bfloat16().to(dev)
        wp, bp = ops.pack_geglu(rn(8 * C, C) * C ** -0.5, rn(8 * C) * 0.1)
        out = torch.empty(M, 4 * C, dtype=torch.bfloat16, device=dev)
        ops.gemm([a], wp.to(dev), 8 * C, out=out, bias=bp.to(dev), geglu=True)
    if "outproj" in which:                     # attention out-proj with fp32 residual: N = 320, K = 384
        M = R * 4096
        a = rn(M, 384).bfloat16().to(dev)
        w = ops.pack_linear(rn(320, 384) * 384 ** -0.5).to(dev)
        res = rn(M, 320).to(dev)
        ops.gemm([a], w, 320, out=res, bias=rn(320).to(dev), residual=res)
    if "qkv" in which:                         # fused q/k/v projection: bf16 out, no bias/residual, N = 1152, K = 320
        M = R * 4096
        a = rn(M, 320).bfloat16().to(dev)
        w = ops.pack_linear(rn(1152, 320) * 320 ** -0.5).to(dev)
        out = torch.empty(M, 1152, dtype=torch.bfloat16, device=dev)
        ops.gemm([a], w, 1152, out=out)
    if "conv" in which:                        # ResNet conv 640->640 at 32x32
        x = rn(R, 32, 32, 640).bfloat16().to(dev)
        w = ops.pack_conv3x3(rn(640, 640, 3, 3) * (9 * 640) ** -0.5).to(dev)
        out = torch.empty(R, 32, 32, 640, dtype=torch.float32, device=dev)
        ops.gemm([x], w, 640, out=out, taps=[ops.TAPS_3X3], conv_geom=(R, 32, 32), bias=rn(640).to(dev))
    if "conv320" in which:                     # ResNet conv 320->320 at 64x64 (N = 320: two 160-column tiles), 1-CTA and CTA pair
        x = rn(R, 64, 64, 320).bfloat16().to(dev)
        w = ops.pack_conv3x3(rn(320, 320, 3, 3) * (9 * 320) ** -0.5).to(dev)
        out = torch.empty(R, 64, 64, 320, dtype=torch.float32, device=dev)
        for cg in (1, 2):
            ops.gemm([x], w, 320, out=out, taps=[ops.TAPS_3X3], conv_geom=(R, 64, 64), bias=rn(320).to(dev), cta_group=cg)
    if "conv320wide" in which:                 # the same conv as two 160-column tiles vs ONE 320-column tile (two MMA sub-tiles sharing A)
        x = rn(R, 64, 64, 320).bfloat16().to(dev)
        w = ops.pack_conv3x3(rn(320, 320, 3, 3) * (9 * 320) ** -0.5).to(dev)
        out = torch.empty(R, 64, 64, 320, dtype=torch.float32, device=dev)
        for bn in (160, 0):                    # 0 = automatic: the wide tile (K = 2880 >= 2048)
            ops.gemm([x], w, 320, out=out, taps=[ops.TAPS_3X3], conv_geom=(R, 64, 64), bias=rn(320).to(dev), cta_group=2, block_n=bn)
    if "conv640" in which:                     # ResNet conv 640->640 at 32x32 (N = 640: three 224-column tiles), 1-CTA and CTA pair
        x = rn(R, 32, 32, 640).bfloat16().to(dev)
        w = ops.pack_conv3x3(rn(640, 640, 3, 3) * (9 * 640) ** -0.5).to(dev)
        out = torch.empty(R, 32, 32, 640, dtype=torch.float32, device=dev)
        for cg in (1, 2):
            ops.gemm([x], w, 640, out=out, taps=[ops.TAPS_3X3], conv_geom=(R, 32, 32), bias=rn(640).to(dev), cta_group=cg)
    if "attn" in which:                        # self-attention at 64x64: S = 4096, 8 heads, d = 40 -> 48
        B = max(1, R // 4)
        qkv = rn(B, 4096, 3 * 384).bfloat16().to(dev)
        o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device=dev)
        ops.attention(qkv[..., :384], qkv[..., 384:768], qkv[..., 768:], o, heads=8, dp=48, scale=40 ** -0.5)
    if "xattn" in which:                       # cross-attention: S_kv = 77
        B = max(1, R // 4)
        q = rn(B, 4096, 384).bfloat16().to(dev)
        kv = rn(B, 77, 768).bfloat16().to(dev)
        o = torch.empty(B, 4096, 384, dtype=torch.bfloat16, device=dev)
        ops.attention(q, kv[..., :384], kv[..., 384:], o, heads=8, dp=48, scale=40 ** -0.5)
    if "gn" in which:                          # GroupNorm+SiLU over a skip concat 640+320 at 64x64
        x0, x1 = rn(R, 64, 64, 640).to(dev), rn(R, 64, 64, 320).to(dev)
        out = torch.empty(R, 64, 64, 960, dtype=torch.bfloat16, device=dev)
        raw = torch.empty_like(out)
        ws = torch.empty(ops.groupnorm_ws_floats(R, 32), dtype=torch.float32, device=dev)
        ops.groupnorm(x0, x1, rn(960).to(dev), rn(960).to(dev), groups=32, eps=1e-5, silu=True, stats_ws=ws, out=out, raw_out=raw)
    if "ln" in which:
        x = rn(R * 4096, 320).to(dev)
        out = torch.empty(R * 4096, 320, dtype=torch.bfloat16, device=dev)
        ops.layernorm(x, rn(320).to(dev), rn(320).to(dev), out)
    if "cfg" in which:                         # CFG + DDIM for 4096 items (bigger than L2)
        N = 1024
        eps = rn(4 * N, 64, 64, 4).to(dev)
        x = rn(N, 4, 64, 64).to(dev)
        ops.cfg_step(eps, [4.0, 1.0, 7.0, -11.0], x, 0.98, [-0.05], x_out=x)
    torch.cuda.synchronize()
print("done")
