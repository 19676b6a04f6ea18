"""GPU parity of the HBM-streaming and norm kernels vs plain PyTorch fp32/fp64 references."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import err_report, rel_l2

pytestmark = pytest.mark.gpu


def _r(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("B,H,W,c0,c1,silu,raw", [
    (2, 64, 64, 320, 0, True, False), (3, 32, 32, 1280, 640, True, True), (2, 16, 16, 640, 320, True, True),
    (2, 8, 8, 1280, 1280, True, True), (2, 32, 32, 640, 0, False, False), (3, 4, 4, 64, 128, True, True),
])
def test_groupnorm(B, H, W, c0, c1, silu, raw):
    from difashion_b200 import ops
    C = c0 + c1
    x0 = (_r((B, H, W, c0), 1) * 2 + 0.5).cuda()
    x1 = (_r((B, H, W, c1), 2) * 0.7 - 0.3).cuda() if c1 else None
    gamma, beta = _r((C,), 3).cuda(), _r((C,), 4).cuda()
    eps = 1e-5 if silu else 1e-6
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    rawo = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda") if raw else None
    ws = torch.empty(ops.groupnorm_ws_floats(B, 32), dtype=torch.float32, device="cuda")
    ops.groupnorm(x0, x1, gamma, beta, groups=32, eps=eps, silu=silu, stats_ws=ws, out=out, raw_out=rawo)
    torch.cuda.synchronize()
    x = x0 if x1 is None else torch.cat([x0, x1], -1)
    ref = F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert rel_l2(out, ref) < 4e-3, err_report(out.reshape(-1, C), ref.reshape(-1, C), "groupnorm")
    if raw:
        assert torch.equal(rawo, x.bfloat16())


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (77, 1280), (64, 64), (5, 128)])
def test_layernorm(rows, C):
    from difashion_b200 import ops
    x = (_r((rows, C), 5) * 3 + 1).cuda()
    gamma, beta = _r((C,), 6).cuda(), _r((C,), 7).cuda()
    out = torch.empty(rows, C, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(x, gamma, beta, out)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5)
    assert rel_l2(out, ref) < 4e-3, err_report(out, ref, "layernorm")


def test_cfg_step_ddim_and_plms_forms():
    from difashion_b200 import ops
    N, H, W = 8, 64, 64
    eps = _r((4 * N, H, W, 4), 8).cuda()
    x = _r((N, 4, H, W), 9).cuda()
    h1, h2, h3 = (_r((N, 4, H, W), 10 + i).cuda() for i in range(3))
    noise = _r((N, 4, H, W), 20).cuda()
    s_h, s_m, s_c = 4.0, 5.0, 12.0
    w = [s_h, s_m - s_h, s_c - s_m, 1.0 - s_c]
    e = eps.reshape(4, N, H, W, 4).permute(0, 1, 4, 2, 3).double()
    e_cfg = e[3] + s_h * (e[0] - e[1]) + s_m * (e[1] - e[2]) + s_c * (e[2] - e[3])
    # DDIM form
    out = ops.cfg_step(eps, w, x, 0.98, [-0.05])
    torch.cuda.synchronize()
    ref = 0.98 * x.double() + (-0.05) * e_cfg
    assert rel_l2(out, ref) < 1e-5, err_report(out.reshape(N * 4, -1), ref.reshape(N * 4, -1), "cfg ddim")
    # PLMS 4th-order form with history + eps_out + noise
    eo = torch.empty_like(x)
    out = ops.cfg_step(eps, w, x, 1.01, [0.55, -0.59, 0.37, -0.09], hist=[h1, h2, h3], noise=noise, cn=0.3, eps_out=eo)
    torch.cuda.synchronize()
    ref = 1.01 * x.double() + 0.55 * e_cfg - 0.59 * h1.double() + 0.37 * h2.double() - 0.09 * h3.double() + 0.3 * noise.double()
    assert rel_l2(out, ref) < 1e-5
    assert rel_l2(eo, e_cfg) < 1e-5


def test_mutual_gather_and_blend():
    from difashion_b200 import ops
    bsz, olen, H, W = 3, 4, 64, 64
    olists = torch.tensor([[0, 0, 0, 0], [5, 0, 7, 9], [0, 3, 0, 4]])
    from difashion_b200.pipeline import mutual_index_table
    idx = mutual_index_table(olists)
    N = int((olists == 0).sum())
    all_lat = _r((bsz * olen, 4, H, W), 30).cuda()
    prev = _r((N, 4, H, W), 31).cuda()
    out = torch.empty(N, 4 * H * W, dtype=torch.bfloat16, device="cuda")
    ops.mutual_gather_sum(all_lat, prev, idx.cuda(), out)
    torch.cuda.synchronize()
    # independent reference following difashion.py:475-488
    gen = olists == 0
    ref, n = [], 0
    gen_row = {}
    for o in range(bsz):
        for i in range(olen):
            if gen[o, i]:
                gen_row[(o, i)] = n
                n += 1
    for o in range(bsz):
        for i in range(olen):
            if not gen[o, i]:
                continue
            acc = torch.zeros(4, H, W, dtype=torch.float64)
            for s in range(olen):
                if s == i:
                    continue
                acc += (prev[gen_row[(o, s)]] if gen[o, s] else all_lat[o * olen + s]).double().cpu()
            ref.append(acc)
    ref = torch.stack(ref).reshape(N, -1)
    assert rel_l2(out.cpu(), ref) < 4e-3, err_report(out.cpu(), ref, "gather")

    x, m, hist = _r((N, 4, H, W), 32).cuda(), _r((N, 4, H, W), 33).cuda(), _r((N, 4, H, W), 34).cuda()
    null = _r((4, H, W), 35).cuda()
    o = torch.empty(4 * N, H, W, 8, dtype=torch.bfloat16, device="cuda")
    ops.mutual_blend(x, m, hist, null, 0.1, [1, 1, 0, 0], [1, 0, 0, 0], o)
    torch.cuda.synchronize()
    nul = null[None].expand(N, -1, -1, -1)
    mm = torch.cat([m, m, nul, nul]).double()
    hh = torch.cat([hist, nul, nul, nul]).double()
    ref = torch.cat([0.9 * torch.cat([x] * 4).double() + 0.1 * mm, hh], 1).permute(0, 2, 3, 1)
    assert rel_l2(o, ref) < 4e-3, err_report(o.reshape(-1, 8), ref.reshape(-1, 8), "blend")


def test_layout_upsample_s2d_temb():
    from difashion_b200 import ops
    x = _r((3, 8, 16, 16), 40).cuda()
    o = torch.empty(3, 16, 16, 8, dtype=torch.bfloat16, device="cuda")
    ops.nchw_to_nhwc_bf16(x, o)
    assert torch.equal(o, x.permute(0, 2, 3, 1).bfloat16())
    y = _r((3, 16, 16, 4), 41).cuda()
    o2 = torch.empty(3, 4, 16, 16, dtype=torch.float32, device="cuda")
    ops.nhwc_to_nchw(y, o2)
    assert torch.equal(o2, y.permute(0, 3, 1, 2))
    z = _r((2, 8, 8, 64), 42).cuda()
    up = torch.empty(2, 16, 16, 64, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x(z, up)
    ref = F.interpolate(z.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(up, ref)
    # stride-2 conv through space-to-depth + tap table
    B, H, W, C = 2, 16, 16, 64
    xin = _r((B, H, W, C), 43).cuda()
    w = _r((C, C, 3, 3), 44, (9 * C) ** -0.5).cuda()
    bias = _r((C,), 45).cuda()
    s2d = torch.empty(B, H // 2, W // 2, 4 * C, dtype=torch.bfloat16, device="cuda")
    ops.space_to_depth(xin, s2d)
    out = torch.empty(B, H // 2, W // 2, C, dtype=torch.float32, device="cuda")
    ops.gemm([s2d], ops.pack_conv3x3(w), C, out=out, taps=[ops.s2d_taps(C)], a_c=[C], conv_geom=(B, H // 2, W // 2), bias=bias)
    torch.cuda.synchronize()
    ref = F.conv2d(xin.bfloat16().double().permute(0, 3, 1, 2), w.bfloat16().double(), bias.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(out, ref) < 1e-5, err_report(out.reshape(-1, C), ref.reshape(-1, C), "s2d conv")
    # timestep embedding
    t = torch.tensor([981.0, 1.0, 500.0]).cuda()
    te = torch.empty(3, 320, dtype=torch.bfloat16, device="cuda")
    ops.timestep_embedding(t, te)
    torch.cuda.synchronize()
    import math
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    arg = t.cpu()[:, None] * freqs[None]
    ref = torch.cat([torch.cos(arg), torch.sin(arg)], -1)
    assert (te.float().cpu() - ref).abs().max() < 1e-2


def test_gemm_activations():
    from difashion_b200 import ops
    a = _r((64, 256), 50).bfloat16().cuda()
    w = _r((512, 256), 51, 1 / 16).cuda()
    b = _r((512,), 52).cuda()
    wp = ops.pack_linear(w)
    y = a.double() @ wp.double().t() + b.double()
    for act, fn in [(ops.ACT_SILU, F.silu), (ops.ACT_LEAKY_RELU, lambda t: F.leaky_relu(t, 0.01)), (ops.ACT_TANH, torch.tanh)]:
        out = torch.empty(64, 512, dtype=torch.float32, device="cuda")
        ops.gemm([a], wp, 512, out=out, bias=b, act=act)
        torch.cuda.synchronize()
        assert rel_l2(out, fn(y)) < 1e-5


@pytest.mark.parametrize("B,H,W,c0,c1", [(2, 64, 64, 320, 0), (3, 16, 16, 640, 320), (2, 8, 8, 1280, 1280), (2, 32, 32, 1280, 640)])
def test_groupnorm_fused_statistics_from_gemm_epilogue(B, H, W, c0, c1):
    """The producing conv/GEMM epilogue emits per-(32-row block, channel pair) partial sums; GroupNorm then skips its
    statistics pass.  Checked against F.group_norm on the producers' actual fp32 outputs."""
    from difashion_b200 import ops
    srcs, parts = [], []
    for i, c in enumerate([c0, c1]):
        if c == 0:
            continue
        a = _r((B * H * W, 64), 60 + i).bfloat16().cuda()
        w = ops.pack_linear(_r((c, 64), 62 + i, 0.3).cuda())
        bias = _r((c,), 64 + i).cuda()
        out = torch.empty(B, H, W, c, dtype=torch.float32, device="cuda")
        part = torch.zeros(ops.gn_partial_shape(B * H * W, c), dtype=torch.float32, device="cuda")
        ops.gemm([a], w, c, out=out.view(B * H * W, c), bias=bias, gn_partial=part)
        torch.cuda.synchronize()
        # the partials are exact block sums of the stored output
        ref_s = out.view(B * H * W // 32, 32, c // 2, 2).double().sum(dim=(1, 3))
        ref_q = (out.view(B * H * W // 32, 32, c // 2, 2).double() ** 2).sum(dim=(1, 3))
        assert rel_l2(part[..., 0], ref_s) < 1e-5 and rel_l2(part[..., 1], ref_q) < 1e-5
        srcs.append(out)
        parts.append(part)
    C = c0 + c1
    gamma, beta = _r((C,), 66).cuda(), _r((C,), 67).cuda()
    ws = torch.empty(ops.groupnorm_ws_floats(B, 32), dtype=torch.float32, device="cuda")
    o = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    ops.groupnorm(srcs[0], srcs[1] if c1 else None, gamma, beta, groups=32, eps=1e-5, silu=True, stats_ws=ws, out=o,
                  partials=(parts[0], parts[1] if c1 else None))
    torch.cuda.synchronize()
    x = srcs[0] if not c1 else torch.cat(srcs, -1)
    ref = F.silu(F.group_norm(x.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)).permute(0, 2, 3, 1)
    assert rel_l2(o, ref) < 4e-3, err_report(o.reshape(-1, C), ref.reshape(-1, C), "gn fused")


def test_mutual_blend_branches_with_equal_flags_hold_equal_bits():
    """CFG branches that get the same (mutual, history) flags must come out of the blend kernel bit-identical for EVERY branch
    count — the shared CFG prefix relies on it.  (Round 1's kernel computed `(1 - eta) * x + eta * m` inside the branch
    loop; nvcc unrolled the loop by two and contracted the expression into an FMA differently in the remainder iteration, so
    with 3 branches the last one differed in the last fp32 bit, flipping a bf16 rounding about once per 65k elements —
    hence 2M elements here.)  Also: the result is the two-products-and-a-sum the reference computes (difashion.py:513)."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(17)
    n, s = 128, 64
    x, m, h = (torch.randn(n, 4, s, s, generator=g).cuda() for _ in range(3))
    z = torch.randn(4, s, s, generator=g).cuda()
    eta = 0.1
    for use_m, use_h in (([1, 1, 1], [1, 0, 0]), ([1, 0, 0], [1, 1, 1]), ([1, 1, 0, 0], [1, 0, 0, 0]), ([1, 1], [1, 0]), ([1, 0, 0], [0, 0, 0])):
        nb = len(use_m)
        for dt in (torch.bfloat16, torch.float32):
            out = torch.empty(nb * n, s, s, 8, dtype=dt, device="cuda")
            ops.mutual_blend(x, m, h, z, eta, use_m, use_h, out)
            o = out.view(nb, n, s, s, 8)
            for a in range(nb):
                for b in range(a + 1, nb):
                    if use_m[a] == use_m[b]:
                        assert torch.equal(o[a, ..., :4], o[b, ..., :4]), (use_m, a, b, dt)
                    if use_h[a] == use_h[b]:
                        assert torch.equal(o[a, ..., 4:], o[b, ..., 4:]), (use_h, a, b, dt)
            if dt == torch.float32:
                for b in range(nb):
                    src = m if use_m[b] else z.expand_as(x)
                    want = ((1 - eta) * x + eta * src).permute(0, 2, 3, 1)        # torch fp32: two products, one sum
                    assert torch.equal(o[b, ..., :4], want), (use_m, b)
