"""GPU parity of the tcgen05 flash-attention kernel vs plain softmax(QK^T*scale)V in fp64 on the same
bf16 operands.  Tolerance rel-L2 <= 1e-2 (P and the output are rounded to bf16; measured ~3e-3)."""
import pytest
import torch

from tests.util import err_report, rel_l2

pytestmark = pytest.mark.gpu


def _ref(q, k, v, heads, d, scale):
    b, sq, _ = q.shape
    skv = k.shape[1]
    qh = q.double().reshape(b, sq, heads, d).permute(0, 2, 1, 3)
    kh = k.double().reshape(b, skv, heads, d).permute(0, 2, 1, 3)
    vh = v.double().reshape(b, skv, heads, d).permute(0, 2, 1, 3)
    w = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (w @ vh).permute(0, 2, 1, 3).reshape(b, sq, heads * d)


def _pad_heads(x, heads, d, dp):
    b, s, _ = x.shape
    out = torch.zeros(b, s, heads, dp, dtype=x.dtype, device=x.device)
    out[..., :d] = x.reshape(b, s, heads, d)
    return out.reshape(b, s, heads * dp)


CASES = [
    (2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160), (3, 8, 64, 64, 160),
    (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80), (2, 8, 256, 77, 160), (2, 8, 64, 85, 160),
    (2, 2, 16, 16, 32), (3, 2, 4, 4, 64), (3, 2, 256, 77, 32), (1, 1, 200, 300, 16),
    # dp <= 64 with more than one 64-key tile: ragged tails of 1 / 12 / 33 / 63 keys, dp = 64
    (1, 2, 130, 97, 32), (2, 3, 256, 160, 64), (1, 2, 100, 225, 40), (2, 2, 300, 127, 48), (1, 4, 128, 65, 40),
]


@pytest.mark.parametrize("B,H,Sq,Skv,d", CASES)
def test_attention(B, H, Sq, Skv, d):
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(Sq * 7 + Skv + d)
    q = torch.randn(B, Sq, H * d, generator=g).bfloat16().cuda()
    k = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    v = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    dp = ops.pad16(d)
    scale = d ** -0.5
    qp, kp, vp = (_pad_heads(t, H, d, dp).contiguous() for t in (q, k, v))
    out = torch.full((B, Sq, H * dp), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.attention(qp, kp, vp, out, heads=H, dp=dp, scale=scale)
    torch.cuda.synchronize()
    got = out.reshape(B, Sq, H, dp)[..., :d].reshape(B, Sq, H * d)
    ref = _ref(q, k, v, H, d, scale)
    assert rel_l2(got, ref) < 1e-2, err_report(got.reshape(-1, H * d), ref.reshape(-1, H * d), f"attn {B},{H},{Sq},{Skv},{d}")


def test_attention_fused_qkv_layout_and_large_logits():
    """q/k/v as column slices of one [B,S,3*H*dp] buffer (the fused projection output); logits scaled
    up so the online-softmax rescale path (running max grows across KV tiles) is exercised."""
    from difashion_b200 import ops
    B, H, S, d = 2, 8, 1024, 40
    dp = ops.pad16(d)
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B, S, H * d, generator=g) * 3).bfloat16().cuda()
    k = (torch.randn(B, S, H * d, generator=g) * 3).bfloat16().cuda()
    k[:, S // 2:] *= 2.0                       # later KV tiles carry larger logits
    v = torch.randn(B, S, H * d, generator=g).bfloat16().cuda()
    qkv = torch.cat([_pad_heads(t, H, d, dp) for t in (q, k, v)], dim=-1).contiguous()
    C = H * dp
    out = torch.zeros(B, S, C, dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out, heads=H, dp=dp, scale=d ** -0.5,
                  q_col0=0, k_col0=0, v_col0=0)
    torch.cuda.synchronize()
    got = out.reshape(B, S, H, dp)[..., :d].reshape(B, S, H * d)
    ref = _ref(q, k, v, H, d, d ** -0.5)
    assert rel_l2(got, ref) < 1e-2, err_report(got.reshape(-1, H * d), ref.reshape(-1, H * d), "attn fused")


@pytest.mark.parametrize("B,H,Sq,Skv,d,bkv", [(2, 8, 1024, 1024, 40, 64), (2, 8, 512, 512, 40, 128), (1, 1, 200, 300, 16, 64),
                                               (2, 8, 256, 256, 160, 64), (2, 4, 384, 1000, 80, 64), (1, 2, 130, 97, 32, 64),
                                               (2, 3, 256, 160, 64, 64), (2, 2, 300, 127, 48, 64), (1, 4, 128, 65, 40, 64)])
def test_attention_variants_smem_p_and_single_buffer(B, H, Sq, Skv, d, bkv):
    """Non-default variants behind the tuning hooks: P through shared memory (bit4), the single-buffer kernel (bit3)
    and the split-KV kernel with eight softmax warps (bit5)."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(Sq + Skv + d)
    q = torch.randn(B, Sq, H * d, generator=g).bfloat16().cuda()
    k = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    v = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    dp = ops.pad16(d)
    qp, kp, vp = (_pad_heads(t, H, d, dp).contiguous() for t in (q, k, v))
    out = torch.full((B, Sq, H * dp), float("nan"), dtype=torch.bfloat16, device="cuda")
    ref = _ref(q, k, v, H, d, d ** -0.5)
    for flags in (16, 8, 32):            # 32: the experimental split-KV kernel (8 softmax warps) where dp <= 64
        out.fill_(float("nan"))
        ops.attention(qp, kp, vp, out, heads=H, dp=dp, scale=d ** -0.5, block_kv=bkv, dbg_flags=flags)
        torch.cuda.synchronize()
        got = out.reshape(B, Sq, H, dp)[..., :d].reshape(B, Sq, H * d)
        assert rel_l2(got, ref) < 1e-2, err_report(got.reshape(-1, H * d), ref.reshape(-1, H * d), f"attn flags={flags}")


@pytest.mark.parametrize("B,H,Sq,Skv,d", [(2, 8, 4096, 77, 40), (3, 4, 1000, 85, 80), (1, 2, 640, 33, 32), (2, 2, 300, 128, 64)])
def test_cross_attention_short_kv_kernel_streams_query_tiles(B, H, Sq, Skv, d):
    """Short-KV kernel (K/V resident, query tiles streamed through the S/O double buffers): 1, 3 and 8 tiles per CTA
    (3 does not divide the tile count: ragged last CTA), ragged last query tile, and the generic kernel (bit6) as control."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(Sq + Skv + d)
    q = torch.randn(B, Sq, H * d, generator=g).bfloat16().cuda()
    k = (1.5 * torch.randn(B, Skv, H * d, generator=g)).bfloat16().cuda()
    v = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    dp = ops.pad16(d)
    qp, kp, vp = (_pad_heads(t, H, d, dp).contiguous() for t in (q, k, v))
    ref = _ref(q, k, v, H, d, d ** -0.5)
    outs = []
    for flags in (1 << 8, 3 << 8, 8 << 8, 64):
        out = torch.full((B, Sq, H * dp), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.attention(qp, kp, vp, out, heads=H, dp=dp, scale=d ** -0.5, dbg_flags=flags)
        torch.cuda.synchronize()
        got = out.reshape(B, Sq, H, dp)[..., :d].reshape(B, Sq, H * d)
        assert rel_l2(got, ref) < 1e-2, err_report(got.reshape(-1, H * d), ref.reshape(-1, H * d), f"short-kv flags={flags}")
        outs.append(out.reshape(B, Sq, H, dp)[..., :d].clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])      # tiles per CTA must not change a bit


def test_attention_no_max_fast_path_variant():
    """dbg bit7: fast path without max tracking (row-tile sum as the overflow sentinel) — a measured-slower experiment
    kept for the record; must stay correct, including when the running maximum keeps growing."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, S, H, d = 2, 1024, 8, 40
    dp = 48
    q, k, v = (torch.randn(B, S, H * dp, generator=g) for _ in range(3))
    k = k * torch.linspace(0.2, 6.0, S)[None, :, None]                     # later keys score higher: frequent rescales
    for t in (q, k, v):
        t.view(B, S, H, dp)[..., d:] = 0
    qb, kb, vb = q.bfloat16().cuda(), k.bfloat16().cuda(), v.bfloat16().cuda()
    outs = []
    for flags in (0, 128):
        out = torch.empty(B, S, H * dp, dtype=torch.bfloat16, device="cuda")
        ops.attention(qb, kb, vb, out, heads=H, dp=dp, scale=d ** -0.5, block_kv=64, dbg_flags=flags)
        outs.append(out.cpu())
    sp = lambda t: t.cpu().double().view(B, S, H, dp).transpose(1, 2)
    ref = (torch.softmax(sp(qb) @ sp(kb).transpose(-1, -2) * d ** -0.5, -1) @ sp(vb)).transpose(1, 2).reshape(B, S, H * dp)
    assert rel_l2(outs[0], ref) < 5e-3 and rel_l2(outs[1], ref) < 5e-3
    # the two variants round P (and the output) to bf16 against different reference maxima, so with these sharply peaked
    # rows they differ at the bf16-rounding level (measured 2.0e-3 on B200, the size of either one's error vs fp64)
    assert rel_l2(outs[1], outs[0]) < 4e-3


@pytest.mark.skipif(__import__("os").environ.get("DFB_TEST_PP") != "1",
                    reason="attn_fwd_pp_kernel (ping-pong softmax warpgroups, dbg bit12) was written after round 1's GPU budget "
                           "was spent and has never run: opt in with DFB_TEST_PP=1 (a faulty kernel can poison the CUDA context)")
@pytest.mark.parametrize("B,H,Sq,Skv,d", [(2, 8, 4096, 4096, 40), (1, 2, 512, 320, 64), (2, 3, 300, 1000, 48), (1, 1, 256, 128, 16)])
def test_attention_ping_pong_variant(B, H, Sq, Skv, d):
    """dbg bit12: one CTA per SM, two query tiles, two softmax warpgroups taking turns on the MUFU (named barriers), K/V
    tiles shared — must agree with the shipped double-buffered kernel to bf16-rounding level and with fp64 to 1e-2."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(Sq + Skv + d)
    q = torch.randn(B, Sq, H * d, generator=g).bfloat16().cuda()
    k = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    v = torch.randn(B, Skv, H * d, generator=g).bfloat16().cuda()
    dp = ops.pad16(d)
    qp, kp, vp = (_pad_heads(t, H, d, dp).contiguous() for t in (q, k, v))
    outs = []
    for flags in (0, 4096):
        out = torch.full((B, Sq, H * dp), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.attention(qp, kp, vp, out, heads=H, dp=dp, scale=d ** -0.5, block_kv=64, dbg_flags=flags)
        torch.cuda.synchronize()
        outs.append(out.reshape(B, Sq, H, dp)[..., :d].reshape(B, Sq, H * d))
    ref = _ref(q, k, v, H, d, d ** -0.5)
    assert rel_l2(outs[1], ref) < 1e-2 and rel_l2(outs[0], ref) < 1e-2
    assert rel_l2(outs[1], outs[0]) < 4e-3
