"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv kernel against plain PyTorch (fp64 math on
the same bf16-rounded operands).  Tolerance: fp32 accumulation of bf16 products -> rel-L2 <= 2e-3
for bf16 outputs (output rounding 2^-9), <= 1e-5 for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import err_report, rel_l2

pytestmark = pytest.mark.gpu


def _ops():
    from difashion_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 32, 64, 32), (128, 64, 64, 0), (128, 128, 256, 0), (256, 256, 512, 0), (384, 320, 320, 0),
    (1000, 640, 1280, 0), (4096, 1280, 320, 0), (77, 100, 768, 0), (16, 1280, 320, 0),
    (4096, 2560, 320, 0), (300, 4, 320, 0), (2048, 1152, 320, 0),
])
def test_plain_gemm_fp32_out(M, N, K, bn):
    ops = _ops()
    a = _rand((M, K), 1).bfloat16().cuda()
    w = _rand((N, K), 2, K ** -0.5).cuda()
    wp = ops.pack_linear(w)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm([a], wp, N, out=out, block_n=bn)
    torch.cuda.synchronize()
    ref = a.double() @ wp[:, :K].double().t()
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, f"gemm {M}x{N}x{K}")


def test_gemm_epilogue_bias_residual_bf16():
    ops = _ops()
    M, N, K = 1024, 640, 640
    a = _rand((M, K), 3).bfloat16().cuda()
    w = _rand((N, K), 4, K ** -0.5).cuda()
    bias = _rand((N,), 5).cuda()
    res = _rand((M, N), 6).cuda()
    wp = ops.pack_linear(w)
    ref = a.double() @ wp.double().t() + bias.double() + res.double()
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm([a], wp, N, out=out, bias=bias, residual=res)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, "bias+res fp32")
    outb = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    resb = res.bfloat16()
    ops.gemm([a], wp, N, out=outb, bias=bias, residual=resb)
    torch.cuda.synchronize()
    refb = a.double() @ wp.double().t() + bias.double() + resb.double()
    assert rel_l2(outb, refb) < 3e-3, err_report(outb, refb, "bias+res bf16")


def test_gemm_rowbias_and_strided_out():
    ops = _ops()
    B, S, N, K = 4, 256, 320, 320
    a = _rand((B * S, K), 7).bfloat16().cuda()
    w = _rand((N, K), 8, K ** -0.5).cuda()
    rb = _rand((B, 1000), 9).cuda()[:, 100:100 + N]          # strided row-bias slice
    wp = ops.pack_linear(w)
    big = torch.zeros(B * S, 2 * N, dtype=torch.float32, device="cuda")
    out = big[:, N:]                                           # write into a column slice
    ops.gemm([a], wp, N, out=out, rowbias=rb, rows_per_batch=S)
    torch.cuda.synchronize()
    ref = a.double() @ wp.double().t() + rb.double().repeat_interleave(S, 0)
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, "rowbias")
    assert float(big[:, :N].abs().max()) == 0.0


def test_gemm_geglu():
    ops = _ops()
    M, C = 512, 320
    a = _rand((M, C), 10).bfloat16().cuda()
    w = _rand((8 * C, C), 11, C ** -0.5).cuda()
    b = _rand((8 * C,), 12, 0.1).cuda()
    wp, bp = ops.pack_geglu(w, b)
    out = torch.empty(M, 4 * C, dtype=torch.bfloat16, device="cuda")
    ops.gemm([a], wp, 8 * C, out=out, bias=bp, geglu=True)
    torch.cuda.synchronize()
    y = a.double() @ w.bfloat16().double().t() + b.double()
    ref = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
    assert rel_l2(out, ref) < 3e-3, err_report(out, ref, "geglu")


def test_gemm_two_segments_concat_k():
    ops = _ops()
    M, K1, K2, N = 640, 320, 640, 320
    a1 = _rand((M, K1), 13).bfloat16().cuda()
    a2 = _rand((M, K2), 14).bfloat16().cuda()
    w = _rand((N, K1 + K2), 15, (K1 + K2) ** -0.5).cuda()
    wp = ops.pack_linear(w)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm([a1, a2], wp, N, out=out)
    torch.cuda.synchronize()
    ref = torch.cat([a1, a2], 1).double() @ wp.double().t()
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, "2seg")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (2, 64, 64, 320, 320), (3, 32, 32, 640, 640), (4, 16, 16, 1280, 1280), (5, 8, 8, 1280, 1280),
    (2, 64, 64, 8, 320), (2, 64, 64, 320, 4), (3, 4, 4, 128, 128), (2, 2, 2, 64, 128),
])
def test_conv3x3(B, H, W, Cin, Cout):
    ops = _ops()
    x = _rand((B, H, W, Cin), 20).bfloat16().cuda()
    w = _rand((Cout, Cin, 3, 3), 21, (9 * Cin) ** -0.5).cuda()
    bias = _rand((Cout,), 22).cuda()
    wp = ops.pack_conv3x3(w)
    out = torch.empty(B, H, W, Cout, dtype=torch.float32, device="cuda")
    ops.gemm([x], wp, Cout, out=out, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=bias)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.bfloat16().double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert rel_l2(out, ref) < 1e-5, err_report(out.reshape(-1, Cout), ref.reshape(-1, Cout), f"conv {B}x{H}x{W} {Cin}->{Cout}")


def test_conv3x3_plus_shortcut_segment_and_rowbias():
    """conv2 of a ResnetBlock2D with the 1x1 shortcut folded in as an extra K segment."""
    ops = _ops()
    B, H, W, Cin, Cout = 2, 32, 32, 960, 640
    hn = _rand((B, H, W, Cout), 30).bfloat16().cuda()
    xr = _rand((B, H, W, Cin), 31).bfloat16().cuda()
    w2 = _rand((Cout, Cout, 3, 3), 32, (9 * Cout) ** -0.5).cuda()
    ws = _rand((Cout, Cin, 1, 1), 33, Cin ** -0.5).cuda()
    temb = _rand((B, Cout), 34).cuda()
    wp = torch.cat([ops.pack_conv3x3(w2), ops.pack_linear(ws)], dim=1).contiguous()
    out = torch.empty(B, H, W, Cout, dtype=torch.float32, device="cuda")
    ops.gemm([hn, xr], wp, Cout, out=out, taps=[ops.TAPS_3X3, ops.TAP_CENTER], conv_geom=(B, H, W),
             rowbias=temb, rows_per_batch=H * W)
    torch.cuda.synchronize()
    ref = (F.conv2d(hn.double().permute(0, 3, 1, 2), w2.bfloat16().double(), padding=1)
           + F.conv2d(xr.double().permute(0, 3, 1, 2), ws.bfloat16().double())
           + temb.double()[:, :, None, None]).permute(0, 2, 3, 1)
    assert rel_l2(out, ref) < 1e-5, err_report(out.reshape(-1, Cout), ref.reshape(-1, Cout), "conv+shortcut")


@pytest.mark.parametrize("M,N,K,out_dtype", [
    (25600, 320, 384, torch.float32),      # 400 tiles of 128x160 on 148 CTAs: chunk rotation + prefetch across tiles
    (25600, 320, 384, torch.bfloat16),
    (19000, 328, 320, torch.float32),      # ragged M (partial row block), N % 32 != 0 (half-valid chunk)
    (33000, 640, 640, torch.float32),
    (70000, 96, 64, torch.bfloat16),       # one k-block, 3 chunks over 2 warps
])
def test_gemm_pipelined_residual_epilogue_many_tiles(M, N, K, out_dtype):
    """Specialised epilogue (residual prefetched one chunk ahead, also across tile boundaries; in-place
    ``out is residual`` as the transformer blocks use it)."""
    ops = _ops()
    a = _rand((M, K), 21).bfloat16().cuda()
    w = _rand((N, K), 22, K ** -0.5).cuda()
    bias = _rand((N,), 23).cuda()
    res = _rand((M, N), 24).cuda()
    wp = ops.pack_linear(w)
    ref = a.double() @ wp[:, :K].double().t() + bias.double() + res.double()
    if out_dtype == torch.float32:
        out = res.clone()
        ops.gemm([a], wp, N, out=out, bias=bias, residual=out)      # in place
        tol = 1e-5
    else:
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        ops.gemm([a], wp, N, out=out, bias=bias, residual=res)
        tol = 3e-3
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < tol, err_report(out, ref, f"res epilogue {M}x{N}x{K}")
    # every row block individually (a wrong tile/chunk mapping hides in a global norm)
    blk = (out.double() - ref).reshape(-1)[: (M // 128) * 128 * N].reshape(M // 128, -1).norm(dim=1) / \
        ref.reshape(-1)[: (M // 128) * 128 * N].reshape(M // 128, -1).norm(dim=1)
    assert float(blk.max()) < 10 * tol


def test_gemm_geglu_many_tiles_matches_exact_erf_gelu():
    ops = _ops()
    M, C = 20000, 320
    a = _rand((M, C), 30).bfloat16().cuda()
    w = _rand((8 * C, C), 31, 2.0 * C ** -0.5).cuda()          # gates spread over ~[-6, 6]
    b = _rand((8 * C,), 32, 0.1).cuda()
    wp, bp = ops.pack_geglu(w, b)
    out = torch.empty(M, 4 * C, dtype=torch.bfloat16, device="cuda")
    ops.gemm([a], wp, 8 * C, out=out, bias=bp, geglu=True)
    torch.cuda.synchronize()
    y = a.double() @ w.bfloat16().double().t() + b.double()
    ref = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
    assert rel_l2(out, ref) < 3e-3, err_report(out, ref, "geglu many tiles")
    # the erf-GELU approximation error (<= 2.7e-5 abs) must stay far below the bf16 output rounding
    err = (out.double() - ref).abs()
    assert float((err - 2.0 ** -8 * ref.abs()).max()) < 2e-4


# ------------------------------------------------------------------------------------------------
# CTA-pair kernel (cluster of 2, tcgen05.mma.cta_group::2): forced with cta_group=2 on shapes of every size, incl.
# an odd number of 128-row tiles (the second CTA of the last pair owns a tile that is entirely out of range)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,bn", [(256, 64, 64, 0), (128, 32, 64, 32), (384, 320, 320, 0), (1000, 640, 1280, 0), (77, 100, 768, 0),
                                     (4096, 1280, 320, 0), (40000, 320, 384, 0), (3 * 128 * 74 + 5, 256, 128, 0)])
def test_pair_plain_gemm_fp32_out(M, N, K, bn):
    ops = _ops()
    a = _rand((M, K), 1).bfloat16().cuda()
    w = _rand((N, K), 2, K ** -0.5).cuda()
    wp = ops.pack_linear(w)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ref1 = torch.empty_like(out)
    ops.gemm([a], wp, N, out=out, block_n=bn, cta_group=2)
    ops.gemm([a], wp, N, out=ref1, block_n=bn, cta_group=1)
    torch.cuda.synchronize()
    ref = a.double() @ wp[:, :K].double().t()
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, f"pair gemm {M}x{N}x{K}")
    assert torch.equal(out, ref1)                      # same MMA order per output row: bitwise the 1-CTA result


def test_pair_epilogues_residual_rowbias_geglu_gn_partials():
    ops = _ops()
    # fp32 residual (pipelined prefetch across work items) + bias, in place, many tiles
    M, N, K = 148 * 128 * 3 + 64, 320, 384
    a = _rand((M, K), 3).bfloat16().cuda()
    w = ops.pack_linear(_rand((N, K), 4, K ** -0.5).cuda())
    bias, res = _rand((N,), 5).cuda(), _rand((M, N), 6).cuda()
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    out = res.clone()
    ops.gemm([a], w, N, out=out, bias=bias, residual=out, cta_group=2)
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, "pair residual")
    outb = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm([a], w, N, out=outb, bias=bias, cta_group=2)
    assert rel_l2(outb, a.double() @ w.double().t() + bias.double()) < 3e-3
    # row bias + GroupNorm partial statistics
    B, S = 6, 1024
    a2 = _rand((B * S, 320), 7).bfloat16().cuda()
    w2 = ops.pack_linear(_rand((320, 320), 8, 320 ** -0.5).cuda())
    rb = _rand((B, 320), 9).cuda()
    o1, o2 = torch.empty(B * S, 320, dtype=torch.float32, device="cuda"), torch.empty(B * S, 320, dtype=torch.float32, device="cuda")
    p1 = torch.zeros(ops.gn_partial_shape(B * S, 320), dtype=torch.float32, device="cuda")
    p2 = torch.zeros_like(p1)
    ops.gemm([a2], w2, 320, out=o1, rowbias=rb, rows_per_batch=S, gn_partial=p1, cta_group=1)
    ops.gemm([a2], w2, 320, out=o2, rowbias=rb, rows_per_batch=S, gn_partial=p2, cta_group=2)
    assert torch.equal(o1, o2) and torch.equal(p1, p2)
    assert rel_l2(o2, a2.double() @ w2.double().t() + rb.double().repeat_interleave(S, 0)) < 1e-5
    # GEGLU (16 epilogue warps) and an activation (generic epilogue)
    C = 320
    a3 = _rand((3000, C), 10).bfloat16().cuda()
    wg, bg = ops.pack_geglu(_rand((8 * C, C), 11, C ** -0.5), _rand((8 * C,), 12, 0.1))
    g1, g2 = torch.empty(3000, 4 * C, dtype=torch.bfloat16, device="cuda"), torch.empty(3000, 4 * C, dtype=torch.bfloat16, device="cuda")
    ops.gemm([a3], wg.cuda(), 8 * C, out=g1, bias=bg.cuda(), geglu=True, cta_group=1)
    ops.gemm([a3], wg.cuda(), 8 * C, out=g2, bias=bg.cuda(), geglu=True, cta_group=2)
    assert torch.equal(g1, g2)
    s1, s2 = torch.empty(3000, 320, dtype=torch.bfloat16, device="cuda"), torch.empty(3000, 320, dtype=torch.bfloat16, device="cuda")
    ops.gemm([a3], w2, 320, out=s1, bias=bias, act=ops.ACT_SILU, cta_group=1)
    ops.gemm([a3], w2, 320, out=s2, bias=bias, act=ops.ACT_SILU, cta_group=2)
    torch.cuda.synchronize()
    assert torch.equal(s1, s2)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(3, 64, 64, 64, 320), (5, 32, 32, 128, 640), (7, 16, 16, 64, 96), (37, 8, 8, 64, 128), (1, 4, 256, 64, 64)])
def test_pair_conv3x3_with_shortcut_segment(B, H, W, Cin, Cout):
    """Implicit-GEMM conv (4-D TMA windows, zero fill = padding) + 1x1 shortcut segment on the CTA-pair kernel;
    odd tile counts and images packed several per tile."""
    ops = _ops()
    x = _rand((B, Cin, H, W), 13).bfloat16().cuda()
    xs = _rand((B, 64, H, W), 14).bfloat16().cuda()
    w = _rand((Cout, Cin, 3, 3), 15, (9 * Cin) ** -0.5).cuda()
    wsc = _rand((Cout, 64, 1, 1), 16, 0.125).cuda()
    bias = _rand((Cout,), 17).cuda()
    wp = torch.cat([ops.pack_conv3x3(w), ops.pack_linear(wsc)], 1).contiguous()
    xn, xsn = x.permute(0, 2, 3, 1).contiguous(), xs.permute(0, 2, 3, 1).contiguous()
    o1 = torch.empty(B, H, W, Cout, dtype=torch.float32, device="cuda")
    o2 = torch.empty_like(o1)
    ops.gemm([xn, xsn], wp, Cout, out=o1, taps=[ops.TAPS_3X3, ops.TAP_CENTER], conv_geom=(B, H, W), bias=bias, cta_group=1)
    ops.gemm([xn, xsn], wp, Cout, out=o2, taps=[ops.TAPS_3X3, ops.TAP_CENTER], conv_geom=(B, H, W), bias=bias, cta_group=2)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.bfloat16().double(), padding=1) + F.conv2d(xs.double(), wsc.bfloat16().double()) + bias.double()[None, :, None, None]
    assert rel_l2(o2.permute(0, 3, 1, 2), ref) < 1e-5, err_report(o2.reshape(-1, Cout), ref.permute(0, 2, 3, 1).reshape(-1, Cout), "pair conv")
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("M,N,K,bn,cg", [(1000, 320, 640, 320, 2), (1000, 320, 640, 320, 1), (4096, 512, 256, 512, 2), (300, 384, 128, 384, 1),
                                        (128 * 150 * 2 + 17, 320, 384, 320, 2), (5000, 640, 320, 320, 2)])
def test_wide_tiles_two_mma_subtiles(M, N, K, bn, cg):
    """block_n in (256, 512]: one work item = two MMA column sub-tiles sharing the A tile, accumulator ring one deep."""
    ops = _ops()
    a = _rand((M, K), 21).bfloat16().cuda()
    w = ops.pack_linear(_rand((N, K), 22, K ** -0.5).cuda())
    bias, res = _rand((N,), 23).cuda(), _rand((M, N), 24).cuda()
    out, ref1 = torch.empty(M, N, dtype=torch.float32, device="cuda"), torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm([a], w, N, out=out, bias=bias, residual=res, block_n=bn, cta_group=cg)
    ops.gemm([a], w, N, out=ref1, bias=bias, residual=res, block_n=160 if N % 160 == 0 else 128, cta_group=1)
    torch.cuda.synchronize()
    ref = a.double() @ w[:, :K].double().t() + bias.double() + res.double()
    assert rel_l2(out, ref) < 1e-5, err_report(out, ref, f"wide tile {M}x{N}x{K} bn={bn} cg={cg}")
    assert torch.equal(out, ref1)


def test_wide_tile_is_the_automatic_choice_for_deep_n320_convs():
    """The UNet's N = 320 3x3 convs (K >= 2880) at full size take the 256 x 320 CTA-pair tile; result bitwise that of
    the forced 2 x 160 tiling; GroupNorm partials included."""
    ops = _ops()
    B, H, W, C = 40, 64, 64, 320
    x = _rand((B, H, W, C), 25).bfloat16().cuda()
    w = ops.pack_conv3x3(_rand((C, C, 3, 3), 26, (9 * C) ** -0.5).cuda())
    bias = _rand((C,), 27).cuda()
    o1, o2 = torch.empty(B, H, W, C, dtype=torch.float32, device="cuda"), torch.empty(B, H, W, C, dtype=torch.float32, device="cuda")
    p1 = torch.zeros(ops.gn_partial_shape(B * H * W, C), dtype=torch.float32, device="cuda")
    p2 = torch.zeros_like(p1)
    ops.gemm([x], w, C, out=o1, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=bias, gn_partial=p1)                    # automatic
    ops.gemm([x], w, C, out=o2, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=bias, gn_partial=p2, block_n=160, cta_group=1)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and torch.equal(p1, p2)
    ref = F.conv2d(x[:2].cpu().permute(0, 3, 1, 2).double(), _rand((C, C, 3, 3), 26, (9 * C) ** -0.5).bfloat16().double(), bias.cpu().double(), padding=1)
    assert rel_l2(o1[:2].permute(0, 3, 1, 2).cpu(), ref) < 1e-5


@pytest.mark.parametrize("B,H,W,C,N", [(3, 8, 8, 1280, 1280), (2, 16, 16, 1280, 1280), (2, 32, 32, 640, 640), (2, 8, 8, 64, 128),
                                       (150, 32, 32, 64, 64)])
def test_upsample_as_four_phase_convs(B, H, W, C, N):
    """Upsample2D (nearest-2x + conv3x3 pad 1) as four 4-tap phase convolutions on the low-resolution input
    (dfb_gemm up2x: rows scattered to pixels (2i + a, 2j + b); GroupNorm partials of the WHOLE output, image-major).
    Reference: fp64 upsample + conv of the same bf16 input with the bf16-rounded ORIGINAL weights — the phases round the
    summed weights instead, so the bar is operand-rounding level (1e-2 of it), and exact vs the phases' own packed weights."""
    ops = _ops()
    x = _rand((B, H, W, C), 11)
    w = _rand((N, C, 3, 3), 12, (9 * C) ** -0.5)
    bias = _rand((N,), 13).cuda()
    lo = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    ops.cast_f32(x.cuda(), lo)
    assert torch.equal(lo.cpu(), x.bfloat16())
    packs = [p.cuda() for p in ops.pack_upsample_phases(w)]
    out = torch.full((B, 2 * H, 2 * W, N), float("nan"), dtype=torch.float32, device="cuda")
    hw_ok = (4 * H * W) % 32 == 0 and (H * W) % 32 == 0
    part = torch.zeros(ops.gn_partial_shape(B * 4 * H * W, N), dtype=torch.float32, device="cuda") if hw_ok else None
    for a in (0, 1):
        for b in (0, 1):
            ops.gemm([lo], packs[2 * a + b], N, out=out, taps=[ops.upsample_phase_taps(a, b)], conv_geom=(B, H, W), bias=bias,
                     gn_partial=part, up_phase=(a, b))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()                       # every output pixel was written by exactly one phase
    nb = min(B, 3)                                          # fp64 reference on a few images (first, and the last ones)
    sel = [0] + list(range(B - nb + 1, B))
    up = F.interpolate(lo[sel].cpu().double().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    # exact reference for the phases' own arithmetic: per phase a 2x2 conv with the packed (summed, bf16) weights
    ref_exact = torch.zeros(len(sel), 2 * H, 2 * W, N, dtype=torch.float64)
    cp = (C + 63) // 64 * 64
    xl = lo[sel].cpu().double()
    for a in (0, 1):
        for b in (0, 1):
            acc = torch.zeros(len(sel), H, W, N, dtype=torch.float64)
            for t, (dh, dw, _) in enumerate(ops.upsample_phase_taps(a, b)):
                sh = torch.zeros_like(xl)
                hs, he = max(0, -dh), min(H, H - dh)
                ws_, we = max(0, -dw), min(W, W - dw)
                sh[:, hs:he, ws_:we] = xl[:, hs + dh:he + dh, ws_ + dw:we + dw]
                acc += sh @ packs[2 * a + b][:, t * cp:t * cp + C].cpu().double().t()
            ref_exact[:, a::2, b::2] = acc + bias.cpu().double()
    got = out[sel].cpu().double()
    assert rel_l2(got, ref_exact) < 1e-5, err_report(got.reshape(-1, N), ref_exact.reshape(-1, N), "phase conv (exact)")
    ref = F.conv2d(up, w.bfloat16().double(), bias.cpu().double(), padding=1).permute(0, 2, 3, 1)
    assert rel_l2(got, ref) < 5e-3
    if part is not None:
        # per image the partial blocks sum to the image's channel-pair sums, whatever their order
        blk = 4 * H * W // 32
        s = part.view(B, blk, N // 2, 2).double().sum(1).cpu()
        o = out.view(B, 4 * H * W, N // 2, 2).double().cpu()
        assert rel_l2(s[..., 0], o.sum(dim=(1, 3))) < 1e-5 and rel_l2(s[..., 1], (o ** 2).sum(dim=(1, 3))) < 1e-5


def test_unet_upsample_phases_match_upsample_then_conv():
    """The UNet with the fused upsample phases vs the literal upsample-then-convolve sequence: same network, the Upsample2D
    convs round summed instead of individual weights, i.e. two bf16-rounding realisations of one computation: each must sit
    within the 1e-2 parity bar of the fp32 oracle, and they differ from each other by less than that bar (measured on
    B200: 5.7e-3, the rounding change of an early up block carried through the rest of the network)."""
    from oracle.unet_oracle import make_oracle_unet, tiny_config
    from difashion_b200.unet import B200UNet2DConditionModel
    cfg = tiny_config()
    o = make_oracle_unet(cfg, seed=0)
    unet = B200UNet2DConditionModel(sample_size=cfg.sample_size, in_channels=cfg.in_channels, block_out_channels=tuple(cfg.block_out_channels),
                                    cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=cfg.attention_head_dim)
    unet.load_state_dict(o.state_dict())
    unet.cuda()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    ctx = torch.randn(4, 77, cfg.cross_attention_dim, generator=g)
    ref = o(x, torch.tensor(500), ctx)
    outs = {}
    try:
        for flag in (True, False):
            B200UNet2DConditionModel.upsample_phases = flag
            outs[flag] = unet(x.cuda(), 500, ctx.cuda()).sample.cpu()
    finally:
        B200UNet2DConditionModel.upsample_phases = True
    e_ph, e_lit, d = rel_l2(outs[True], ref), rel_l2(outs[False], ref), rel_l2(outs[True], outs[False])
    print(f"\n[upsample phases] vs oracle: phases {e_ph:.3e}, literal {e_lit:.3e}; phases vs literal {d:.3e}")
    assert e_ph <= 1e-2 and e_lit <= 1e-2
    assert 0 < d < 1e-2
