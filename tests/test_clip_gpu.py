"""GPU parity of the text stage (``B200CLIPTextModel``; reference ``difashion.py:339-352``) against the CPU oracle
(oracle/clip_oracle.py, itself pinned to transformers' CLIPTextModel) and against the committed golden fixture that
transformers produced (tests/golden/clip_tiny.pt).  BASELINE.json states tolerances for the UNet's noise prediction
only; this stage is held to rel-L2 <= 1e-4 on the fp32 verification path and <= 1e-2 on the bf16 tensor-core path."""
import os

import pytest
import torch

from tests.util import err_report, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _ref_attention(q, k, v, heads, scale, causal):
    B, S, _ = q.shape
    hd = q.shape[-1] // heads
    sp = lambda t: t.double().view(B, t.shape[1], heads, hd).transpose(1, 2)
    s = sp(q) @ sp(k).transpose(-1, -2) * scale
    if causal:
        s = s + torch.full((S, S), float("-inf"), dtype=torch.float64).triu(1)
    return (torch.softmax(s, -1) @ sp(v)).transpose(1, 2).reshape(B, S, heads * hd)


@pytest.mark.parametrize("B,H,S,d,dtype", [(3, 12, 77, 64, torch.bfloat16), (2, 4, 77, 16, torch.bfloat16), (2, 2, 200, 64, torch.bfloat16),
                                          (1, 3, 33, 32, torch.bfloat16), (2, 4, 77, 64, torch.float32), (1, 2, 150, 32, torch.float32)])
def test_causal_attention(B, H, S, d, dtype):
    """dfb_attention / dfb_attention_f32 with causal = 1 (fused q|k|v operand, one or several key tiles)."""
    from difashion_b200 import ops
    D = H * d
    qkv = _rand((B, S, 3 * D), 21).to(dtype).cuda()
    out = torch.empty(B, S, D, dtype=dtype, device="cuda")
    ops.attention(qkv, qkv, qkv, out, heads=H, dp=d, scale=d ** -0.5, q_col0=0, k_col0=D, v_col0=2 * D, causal=True)
    torch.cuda.synchronize()
    c = qkv.cpu()
    ref = _ref_attention(c[..., :D], c[..., D:2 * D], c[..., 2 * D:], H, d ** -0.5, True)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_l2(out.cpu(), ref) < tol, err_report(out.cpu().reshape(B * S, D), ref.reshape(B * S, D), "causal attention")
    # row 0 attends to key 0 only
    assert rel_l2(out[:, 0].cpu(), c[:, 0, 2 * D:]) < (1e-6 if dtype == torch.float32 else 1e-2)
    # non-causal call on the same operands is unchanged by the new flag
    ops.attention(qkv, qkv, qkv, out, heads=H, dp=d, scale=d ** -0.5, q_col0=0, k_col0=D, v_col0=2 * D)
    ref = _ref_attention(c[..., :D], c[..., D:2 * D], c[..., 2 * D:], H, d ** -0.5, False)
    assert rel_l2(out.cpu(), ref) < tol


def test_embed_tokens_quick_gelu_and_uint8():
    from difashion_b200 import ops
    tok, pos = _rand((500, 64), 1).cuda(), _rand((77, 64), 2).cuda()
    ids = torch.randint(0, 500, (5, 77), generator=torch.Generator().manual_seed(3))
    out = torch.empty(5 * 77, 64, dtype=torch.float32, device="cuda")
    ops.embed_tokens(ids.to(torch.int32).cuda(), tok, pos, out)
    assert torch.equal(out.cpu().view(5, 77, 64), tok.cpu()[ids] + pos.cpu()[None])
    # quick-GELU epilogue (tensor-core and fp32 paths)
    M, N, K = 300, 256, 128
    a, w, b = _rand((M, K), 4), _rand((N, K), 5, K ** -0.5), _rand((N,), 6)
    for dt, tol in ((torch.bfloat16, 4e-3), (torch.float32, 2e-6)):
        o = torch.empty(M, N, dtype=dt, device="cuda")
        ops.gemm([a.to(dt).cuda()], ops.pack_linear(w.cuda(), dt), N, out=o, bias=b.cuda(), act=ops.ACT_QUICK_GELU)
        y = a.to(dt).double() @ w.to(dt).double().t() + b.double()
        assert rel_l2(o.cpu(), y * torch.sigmoid(1.702 * y)) < tol
    # VaeImageProcessor.postprocess arithmetic, bit-exact against numpy (round half to even)
    img = _rand((2, 16, 24, 4), 7, 0.8)
    img[0, 0, 0, :3] = torch.tensor([-1.0 + 1.0 / 255.0, 0.0, 1.0 / 255.0])        # exact .5 cases
    u8 = torch.empty(2, 16, 24, 3, dtype=torch.uint8, device="cuda")
    ops.image_to_uint8(img.cuda(), u8)
    ref = ((img[..., :3] / 2 + 0.5).clamp(0, 1).numpy() * 255).round().astype("uint8")
    assert (u8.cpu().numpy() == ref).all()


def _mk(which, seed=0):
    from difashion_b200.clip import B200CLIPTextModel
    from oracle.clip_oracle import CLIPTextConfigLite, make_oracle_clip, tiny_clip_config
    # "*gelu": SD-2-base's text encoder activation (exact erf GELU; the reference's default base model, train.py:44)
    cfg = {"tiny": tiny_clip_config, "full": CLIPTextConfigLite,
           "tiny_gelu": lambda: tiny_clip_config(hidden_act="gelu", num_attention_heads=2)}[which]()
    o = make_oracle_clip(cfg, seed=seed)
    m = B200CLIPTextModel(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                          num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                          hidden_act=cfg.hidden_act)
    m.load_transformers_state_dict(o.state_dict())
    return cfg, o, m.cuda()


@pytest.mark.parametrize("which,B,precision,tol", [("tiny", 5, "bf16", 1e-2), ("tiny", 3, "fp32", 1e-4), ("full", 3, "bf16", 1e-2),
                                                   ("full", 2, "fp32", 1e-4), ("tiny_gelu", 4, "bf16", 1e-2),
                                                   ("tiny_gelu", 3, "fp32", 1e-4)])
def test_clip_text_model_matches_oracle(which, B, precision, tol):
    cfg, oracle, m = _mk(which)
    m.set_precision(precision)
    ids = torch.randint(0, cfg.vocab_size, (B, 77), generator=torch.Generator().manual_seed(9))
    ids[-1] = m.null_input_ids()[0].clamp_max(cfg.vocab_size - 1)                    # the empty prompt's shape
    ref = oracle(ids)[0]
    out = m(ids.cuda())
    got, got2 = out[0], m(ids, return_dict=False)[0]                               # device ids / host ids
    torch.cuda.synchronize()
    e = rel_l2(got.cpu(), ref)
    print(f"\n[clip {which} {precision} B={B}] last_hidden_state rel-L2 {e:.3e}")
    assert got.shape == (B, 77, cfg.hidden_size) and got.dtype == torch.float32 and out.last_hidden_state is got
    assert e <= tol, err_report(got.cpu().reshape(B * 77, -1), ref.reshape(B * 77, -1), "clip")
    assert torch.equal(got, got2)
    for i in range(B):
        assert rel_l2(got[i].cpu(), ref[i]) <= 2 * tol
    # chunked batches and shorter sequences
    assert torch.equal(m(ids.cuda(), max_batch=2)[0], got)
    short = m(ids[:, :20].cuda())[0]
    assert rel_l2(short.cpu(), oracle(ids[:, :20])[0]) <= tol


@pytest.mark.parametrize("fixture", ["clip_tiny.pt", "clip_tiny_gelu.pt"])
def test_clip_against_transformers_golden_fixture(fixture):
    """tests/golden/clip_tiny*.pt: weights, ids and last_hidden_state produced by transformers.CLIPTextModel."""
    from difashion_b200.clip import B200CLIPTextModel
    gold = torch.load(os.path.join(GOLD, fixture))
    m = B200CLIPTextModel(**gold["config"])
    m.load_transformers_state_dict(gold["state_dict"])
    m.cuda()
    for precision, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        m.set_precision(precision)
        got = m(gold["input_ids"].cuda())[0]
        e = rel_l2(got.cpu(), gold["last_hidden_state"])
        print(f"\n[clip golden {precision}] rel-L2 vs transformers {e:.3e}")
        assert e <= tol
    with pytest.raises(IndexError):
        m(torch.full((1, 77), gold["config"]["vocab_size"], dtype=torch.long).cuda())
    with pytest.raises(ValueError):
        m(torch.zeros(1, 78, dtype=torch.long).cuda())
