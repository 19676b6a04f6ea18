"""CPU self-checks pinning the oracle: structure, closed-form identities, golden vectors."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle.generation_oracle import (guidance_flags, make_oracle_mutual_encoder, mutual_indices, oracle_generation)
from oracle.schedulers_oracle import OracleDDIMScheduler, OraclePNDMScheduler
from oracle.unet_oracle import (Attention, OracleUNet2DConditionModel, UNetConfig, make_oracle_unet, tiny_config,
                                timestep_embedding)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _count(cfg):
    with torch.device("meta"):
        m = OracleUNet2DConditionModel(cfg)
    return sum(p.numel() for p in m.parameters()), list(m.state_dict().keys())


def test_sd15_structure_param_and_tensor_counts():
    n4, keys4 = _count(UNetConfig(in_channels=4))
    n8, keys8 = _count(UNetConfig(in_channels=8))
    assert n4 == 859_520_964                       # published SD-1.5 UNet size
    assert n8 == 859_532_484 == n4 + 320 * 4 * 9   # DiFashion's widened conv_in (difashion.py:82-93)
    assert len(keys4) == len(keys8) == 686
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.time_emb_proj.bias",
              "down_blocks.1.resnets.0.conv_shortcut.weight", "down_blocks.2.attentions.1.transformer_blocks.0.attn2.to_k.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight", "down_blocks.2.downsamplers.0.conv.bias",
              "mid_block.attentions.0.proj_out.weight", "up_blocks.0.upsamplers.0.conv.weight",
              "up_blocks.3.attentions.2.transformer_blocks.0.attn1.to_out.0.bias", "up_blocks.3.resnets.2.conv_shortcut.bias",
              "conv_norm_out.weight", "conv_out.bias"):
        assert k in keys8, k
    assert not any(k.startswith("down_blocks.3.attentions") or k.startswith("up_blocks.0.attentions") for k in keys8)
    assert not any(k.startswith("down_blocks.3.downsamplers") or k.startswith("up_blocks.3.upsamplers") for k in keys8)


def test_layer_counts():
    with torch.device("meta"):
        m = OracleUNet2DConditionModel(UNetConfig())
    import oracle.unet_oracle as U
    res = [x for x in m.modules() if isinstance(x, U.ResnetBlock2D)]
    assert len(res) == 22 and sum(r.conv_shortcut is not None for r in res) == 14
    assert sum(isinstance(x, U.Transformer2DModel) for x in m.modules()) == 16
    assert sum(isinstance(x, U.Attention) for x in m.modules()) == 32
    assert sum(isinstance(x, torch.nn.GroupNorm) for x in m.modules()) == 61
    assert sum(isinstance(x, torch.nn.LayerNorm) for x in m.modules()) == 48
    up_in = [r.conv1.weight.shape[1] for b in m.up_blocks for r in b.resnets]
    assert up_in == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]


def test_tiny_unet_golden_and_determinism():
    gold = torch.load(os.path.join(GOLD, "tiny_unet.pt"))
    unet = make_oracle_unet(tiny_config(), seed=0)
    y = unet(gold["x"], torch.tensor(gold["t"]), gold["ctx"])
    assert y.shape == (2, 4, 16, 16)
    assert torch.allclose(y, gold["y"], rtol=1e-4, atol=1e-5)
    assert torch.equal(y, unet(gold["x"], gold["t"], gold["ctx"]))          # python-int timestep == 0-d tensor


def test_zero_init_history_channels_have_no_effect():
    """difashion.py:91-92: the widened conv_in starts with zeros on the 4 history channels."""
    unet = make_oracle_unet(tiny_config(), seed=0)
    unet.conv_in.weight[:, 4:].zero_()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 8, 16, 16, generator=g)
    ctx = torch.randn(1, 77, 64, generator=g)
    x2 = x.clone()
    x2[:, 4:] = torch.randn(1, 4, 16, 16, generator=g)
    assert torch.equal(unet(x, 10, ctx), unet(x2, 10, ctx))


def test_attention_matches_sdpa():
    torch.manual_seed(0)
    a = Attention(64, 48, 4, 16).eval()
    x, ctx = torch.randn(2, 10, 64), torch.randn(2, 7, 48)
    q, k, v = a.to_q(x), a.to_k(ctx), a.to_v(ctx)
    sp = lambda t: t.reshape(2, -1, 4, 16).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(2, 10, 64)
    ref = a.to_out[0](ref)
    assert torch.allclose(a(x, ctx), ref, atol=1e-6)


def test_timestep_embedding_is_cos_then_sin():
    e = timestep_embedding(torch.tensor([3.0]), 320)
    f0 = 1.0
    f1 = math.exp(-math.log(10000) * 1 / 160)
    assert abs(float(e[0, 0]) - math.cos(3.0 * f0)) < 1e-6 and abs(float(e[0, 160]) - math.sin(3.0 * f0)) < 1e-6
    assert abs(float(e[0, 1]) - math.cos(3.0 * f1)) < 1e-6


def test_ddim_closed_forms():
    s = OracleDDIMScheduler()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert ts[0] == 981 and ts[-1] == 1 and len(ts) == 50 and ts[1] == 961
    x = torch.randn(2, 4, 8, 8)
    # eps == 0 -> pure rescale by sqrt(a_prev / a_t)
    a_t, a_p = s.alphas_cumprod[981], s.alphas_cumprod[961]
    out = s.step(torch.zeros_like(x), 981, x)[0]
    assert torch.allclose(out, x * (a_p / a_t) ** 0.5, rtol=1e-5, atol=1e-6)
    # exact eps of a known x0 -> analytic x_{t-1}
    x0, eps = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    xt = a_t ** 0.5 * x0 + (1 - a_t) ** 0.5 * eps
    out = s.step(eps, 981, xt)[0]
    assert torch.allclose(out, a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps, rtol=1e-4, atol=1e-5)
    # last step uses final_alpha_cumprod = alphas_cumprod[0] (set_alpha_to_one=False)
    out = s.step(eps, 1, xt)[0]
    assert torch.isfinite(out).all()


def test_pndm_timesteps_and_start():
    s = OraclePNDMScheduler()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-1] == 1
    x = torch.randn(1, 4, 8, 8)
    e1, e2 = torch.randn_like(x), torch.randn_like(x)
    x1 = s.step(e1, 981, x)[0]
    # second call (counter == 1) redoes the first step from the stored sample with the averaged eps
    x2 = s.step(e2, 961, x1)[0]
    ref = s._prev_sample(x, 981, 961, (e1 + e2) / 2)
    assert torch.allclose(x2, ref, atol=1e-6)
    assert s.counter == 2 and len(s.ets) == 1


def test_scheduler_golden():
    gold = torch.load(os.path.join(GOLD, "schedulers.pt"))
    for name, cls in (("ddim", OracleDDIMScheduler), ("pndm", OraclePNDMScheduler)):
        s = cls()
        s.set_timesteps(6)
        assert torch.equal(s.timesteps, gold[name]["timesteps"])
        x = gold["x0"].clone()
        for i in range(gold[name]["traj"].shape[0]):
            x = s.step(gold["eps"][i], s.timesteps[i], x)[0]
            assert torch.allclose(x, gold[name]["traj"][i], rtol=1e-5, atol=1e-6), (name, i)


def test_guidance_flags_and_indices():
    assert guidance_flags(True, True, 12, 4, 5) == (True, True, True, True)
    assert guidance_flags(True, True, 12, 1.0, 5) == (False, True, True, False)
    assert guidance_flags(False, True, 12, 4, 5) == (False, True, True, False)
    mi = mutual_indices(torch.tensor([[0, 0, 5, 0], [7, 0, 8, 9]]))
    assert mi.tolist() == [[-1, -2, 2, -3], [4, -4, 6, 7]]


def test_generation_golden_and_cfg_identity():
    gold = torch.load(os.path.join(GOLD, "generation_tiny.pt"))
    unet = make_oracle_unet(tiny_config(), seed=0)
    me = make_oracle_mutual_encoder(seed=1, latent_size=16, hid_dim=64)
    rec = []
    lat = oracle_generation(unet, me, OracleDDIMScheduler(), **gold["inputs"], num_inference_steps=50, max_steps=3, record=rec)
    assert torch.allclose(lat, gold["latents"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(rec[0]["noise_pred"], gold["eps_step0"], rtol=1e-4, atol=1e-5)
    # all scales == 1 -> no guidance branches: eps is the single fully-conditioned prediction
    rec1 = []
    oracle_generation(unet, me, OracleDDIMScheduler(), **gold["inputs"], num_inference_steps=50, max_steps=1, record=rec1,
                      category_guidance_scale=1.0, hist_guidance_scale=1.0, mutual_guidance_scale=1.0)
    n = gold["inputs"]["init_latents"].shape[0]
    assert rec1[0]["noise_pred_branches"].shape[0] == n
    assert torch.allclose(rec1[0]["noise_pred"], gold["eps_branches_step0"][:n], rtol=1e-4, atol=1e-5)


def test_vae_decoder_oracle_structure_and_shapes():
    """AutoencoderKL decode half: published SD-VAE parameter split, diffusers key names."""
    from oracle.vae_oracle import make_oracle_vae, tiny_vae_config
    m = make_oracle_vae()
    assert sum(p.numel() for p in m.decoder.parameters()) == 49_490_179
    assert sum(p.numel() for p in m.post_quant_conv.parameters()) == 20
    # 83 653 863 (published SD VAE total) = encoder 34 163 592 + quant_conv 72 + decoder + post_quant_conv
    assert 34_163_592 + 72 + 49_490_179 + 20 == 83_653_863
    sd = m.state_dict()
    assert len(sd) == 140
    for k in ("post_quant_conv.weight", "decoder.conv_in.bias", "decoder.mid_block.attentions.0.group_norm.weight",
              "decoder.mid_block.attentions.0.to_q.bias", "decoder.mid_block.attentions.0.to_out.0.weight",
              "decoder.up_blocks.2.resnets.0.conv_shortcut.weight", "decoder.up_blocks.0.upsamplers.0.conv.weight",
              "decoder.up_blocks.3.resnets.2.conv2.bias", "decoder.conv_norm_out.weight", "decoder.conv_out.bias"):
        assert k in sd, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd
    assert sd["decoder.up_blocks.2.resnets.0.conv1.weight"].shape == (256, 512, 3, 3)
    assert sd["decoder.up_blocks.3.resnets.0.conv1.weight"].shape == (128, 256, 3, 3)
    t = make_oracle_vae(tiny_vae_config(), seed=3)
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    y = t.decode_latents(z)
    assert y.shape == (2, 3, 128, 128) and torch.isfinite(y).all()
    assert torch.equal(y, t.decode(z / t.cfg.scaling_factor))
    # the attention's residual connection and the decoder are deterministic
    assert torch.equal(y, make_oracle_vae(tiny_vae_config(), seed=3).decode_latents(z))
