"""fp32 verification path (``set_precision("fp32")``: dfb_gemm_f32 / dfb_attention_f32 / fp32 operands everywhere)
against the CPU oracle.  Tolerance is BASELINE.json's fp32 bar: per-step noise-prediction rel-L2 <= 1e-4,
final-latent cosine >= 0.999 (in practice 1 - 1e-9)."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_unet_gpu import _cos, _gen_inputs, _mk
from tests.util import err_report, rel_l2

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def test_gemm_f32_plain_bias_rowbias_residual_act():
    from difashion_b200 import ops
    M, N, K, S = 1000, 328, 700, 250
    a, w = _rand((M, K), 1).cuda(), _rand((N, K), 2, K ** -0.5).cuda()
    bias, rb, res = _rand((N,), 3).cuda(), _rand((4, N), 4).cuda(), _rand((M, N), 5).cuda()
    wp = ops.pack_linear(w, torch.float32)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm([a], wp, N, out=out, bias=bias, rowbias=rb, rows_per_batch=S, residual=res)
    ref = a.double() @ w.double().t() + bias.double() + rb.double().repeat_interleave(S, 0) + res.double()
    assert rel_l2(out, ref) < 2e-6, err_report(out, ref, "gemm_f32")
    ops.gemm([a], wp, N, out=out, bias=bias, act=ops.ACT_SILU)
    ref = F.silu(a.double() @ w.double().t() + bias.double())
    assert rel_l2(out, ref) < 2e-6, err_report(out, ref, "gemm_f32 silu")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 96), (3, 8, 8, 40, 64), (1, 32, 32, 8, 32)])
def test_conv3x3_f32_with_shortcut_segment(B, H, W, Cin, Cout):
    from difashion_b200 import ops
    x = _rand((B, Cin, H, W), 6).cuda()
    xs = _rand((B, 24, H, W), 7).cuda()
    w, ws, bias = _rand((Cout, Cin, 3, 3), 8, (9 * Cin) ** -0.5).cuda(), _rand((Cout, 24, 1, 1), 9, 0.2).cuda(), _rand((Cout,), 10).cuda()
    xn, xsn = x.permute(0, 2, 3, 1).contiguous(), xs.permute(0, 2, 3, 1).contiguous()
    wp = torch.cat([ops.pack_conv3x3(w, torch.float32), ops.pack_linear(ws, torch.float32)], 1).contiguous()
    out = torch.empty(B, H, W, Cout, dtype=torch.float32, device="cuda")
    ops.gemm([xn, xsn], wp, Cout, out=out, taps=[ops.TAPS_3X3, ops.TAP_CENTER], conv_geom=(B, H, W), bias=bias)
    ref = (F.conv2d(x.double(), w.double(), padding=1) + F.conv2d(xs.double(), ws.double()) + bias.double()[None, :, None, None])
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 2e-6, err_report(out.permute(0, 3, 1, 2).reshape(B * Cout, -1), ref.reshape(B * Cout, -1), "conv_f32")


def test_geglu_f32_and_stride2_conv():
    from difashion_b200 import ops
    M, C = 300, 64
    a, w, b = _rand((M, C), 11).cuda(), _rand((8 * C, C), 12, C ** -0.5).cuda(), _rand((8 * C,), 13, 0.1).cuda()
    wp, bp = ops.pack_geglu(w, b, torch.float32)
    out = torch.empty(M, 4 * C, dtype=torch.float32, device="cuda")
    ops.gemm([a], wp, 8 * C, out=out, bias=bp, geglu=True)
    y = a.double() @ w.double().t() + b.double()
    ref = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
    assert rel_l2(out, ref) < 2e-6, err_report(out, ref, "geglu_f32")
    # stride-2 conv through space-to-depth (fp32 operands)
    B, H, W, c = 2, 16, 16, 64
    x = _rand((B, c, H, W), 14).cuda()
    wc, bc = _rand((c, c, 3, 3), 15, (9 * c) ** -0.5).cuda(), _rand((c,), 16).cuda()
    s2d = torch.empty(B, H // 2, W // 2, 4 * c, dtype=torch.float32, device="cuda")
    ops.space_to_depth(x.permute(0, 2, 3, 1).contiguous(), s2d)
    o = torch.empty(B, H // 2, W // 2, c, dtype=torch.float32, device="cuda")
    ops.gemm([s2d], ops.pack_conv3x3(wc, torch.float32), c, out=o, taps=[ops.s2d_taps(c)], a_c=[c], conv_geom=(B, H // 2, W // 2), bias=bc)
    ref = F.conv2d(x.double(), wc.double(), bc.double(), stride=2, padding=1)
    assert rel_l2(o.permute(0, 3, 1, 2), ref) < 2e-6


@pytest.mark.parametrize("B,H,Sq,Skv,d", [(2, 8, 256, 256, 40), (1, 4, 200, 77, 80), (2, 2, 64, 85, 160), (1, 3, 130, 33, 8)])
def test_attention_f32(B, H, Sq, Skv, d):
    from difashion_b200 import ops
    dp = ops.pad16(d)
    q, k, v = _rand((B, H, Sq, d), 20), _rand((B, H, Skv, d), 21), _rand((B, H, Skv, d), 22)
    def lay(t):                      # [B,H,S,d] -> [B,S,H*dp] with zero padding columns
        o = torch.zeros(B, t.shape[2], H, dp)
        o[..., :d] = t.permute(0, 2, 1, 3)
        return o.reshape(B, t.shape[2], H * dp).cuda()
    out = torch.empty(B, Sq, H * dp, dtype=torch.float32, device="cuda")
    ops.attention(lay(q), lay(k), lay(v), out, heads=H, dp=dp, scale=d ** -0.5)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double())
    got = out.view(B, Sq, H, dp)[..., :d].permute(0, 2, 1, 3).cpu()
    assert rel_l2(got, ref) < 2e-6


@pytest.mark.parametrize("which,B,S", [("tiny", 4, 77), ("tiny", 3, 85), ("full", 1, 77)])
def test_unet_forward_fp32_matches_oracle(which, B, S):
    oracle, unet = _mk(which)
    unet.set_precision("fp32")
    cfg = oracle.cfg
    g = torch.Generator().manual_seed(123)
    x = torch.randn(B, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    ctx = torch.randn(B, S, cfg.cross_attention_dim, generator=g)
    ref = oracle(x, torch.tensor(981), ctx)
    got = unet(x.cuda(), 981, ctx.cuda()).sample
    torch.cuda.synchronize()
    e = rel_l2(got.cpu(), ref)
    print(f"\n[fp32 {which} B={B}] eps rel-L2 {e:.3e}")
    assert e <= FP32_TOL
    # and the default precision still works on the same module afterwards (separate pack / workspace)
    unet.set_precision("bf16")
    got_bf16 = unet(x.cuda(), 981, ctx.cuda()).sample
    assert FP32_TOL < rel_l2(got_bf16.cpu(), ref) <= 1e-2


def test_generation_fp32_tiny_gor_and_fitb():
    from oracle.generation_oracle import make_oracle_mutual_encoder, oracle_generation
    from oracle.schedulers_oracle import OracleDDIMScheduler
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler
    oracle, unet = _mk("tiny")
    unet.set_precision("fp32")
    cfg = oracle.cfg
    ome = make_oracle_mutual_encoder(seed=1, latent_size=cfg.sample_size, hid_dim=64)
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64)
    me.load_state_dict(ome.state_dict())
    for olists, steps in ((torch.zeros(2, 4, dtype=torch.long), 50), (torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2]]), 10)):
        inp = _gen_inputs(cfg, olists)
        rec_o, rec_g = [], []
        lat_o = oracle_generation(oracle, ome, OracleDDIMScheduler(), **inp, num_inference_steps=50, record=rec_o, max_steps=steps)
        pipe = B200DiFashionPipeline(unet, me.cuda(), B200DDIMScheduler(), eta_mutual=0.1, use_cuda_graph=False)
        dev_inp = {k: (v.cuda() if k != "olists" else v) for k, v in inp.items()}
        lat_g = pipe.generate(**dev_inp, num_inference_steps=50, record=rec_g, max_steps=steps).cpu()
        worst = max(rel_l2(rg["eps_branches"][0].permute(0, 3, 1, 2).cpu(), ro["noise_pred_branches"]) for rg, ro in zip(rec_g, rec_o))
        print(f"\n[fp32 generation, {steps} steps] worst per-step eps rel-L2 {worst:.3e}; final latents rel-L2 {rel_l2(lat_g, lat_o):.3e}")
        assert worst <= FP32_TOL
        assert _cos(lat_g, lat_o) >= 0.999 and rel_l2(lat_g, lat_o) <= 1e-3
