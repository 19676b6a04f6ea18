"""Host-side (no GPU): the reference's on-disk model layout — a Stable-Diffusion directory (difashion.py:64-79) plus a
DiFashion checkpoint (<ckpt>/unet, <ckpt>/fashion_encoder; inf4eval.py:543-581) — loads into the B200 modules, in both
weight formats diffusers 0.18.2 / transformers write, with the reference's conv_in surgery (difashion.py:82-93)."""
import json
import os

import pytest
import torch

TINY_UNET = dict(sample_size=16, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64, attention_head_dim=2)
TINY_CLIP = dict(vocab_size=120, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                 bos_token_id=118, eos_token_id=119, pad_token_id=119)
TINY_VAE = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=1, norm_num_groups=32)
SD_SCHED = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.8.0", "beta_end": 0.012, "beta_schedule": "scaled_linear",
            "beta_start": 0.00085, "clip_sample": False, "num_train_timesteps": 1000, "prediction_type": "epsilon",
            "set_alpha_to_one": False, "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None}


def _sd_dir(root, safe):
    """A tiny Stable-Diffusion-shaped directory: 4-channel pretrained UNet, VAE, CLIP text encoder, scheduler config."""
    from difashion_b200 import B200AutoencoderKL, B200CLIPTextModel, B200UNet2DConditionModel
    torch.manual_seed(5)
    unet = B200UNet2DConditionModel(in_channels=4, **TINY_UNET)
    unet.save_pretrained(os.path.join(root, "unet"), safe_serialization=safe)
    vae = B200AutoencoderKL(**TINY_VAE)
    vae.save_pretrained(os.path.join(root, "vae"), safe_serialization=safe)
    clip = B200CLIPTextModel(**TINY_CLIP)
    clip.save_pretrained(os.path.join(root, "text_encoder"), safe_serialization=safe)
    os.makedirs(os.path.join(root, "scheduler"))
    with open(os.path.join(root, "scheduler", "scheduler_config.json"), "w") as f:
        json.dump(SD_SCHED, f)
    return unet, vae, clip


@pytest.mark.parametrize("safe", [False, True])
def test_assemble_from_sd_directory_and_checkpoint(tmp_path, safe):
    from difashion_b200 import B200DDIMScheduler, B200DiFashion, B200PNDMScheduler
    root = str(tmp_path / "sd")
    unet0, vae0, clip0 = _sd_dir(root, safe)
    ext = "safetensors" if safe else "bin"
    assert os.path.exists(os.path.join(root, "unet", f"diffusion_pytorch_model.{ext}"))
    assert os.path.exists(os.path.join(root, "text_encoder", "model.safetensors" if safe else "pytorch_model.bin"))

    m = B200DiFashion.from_pretrained(root, cate_num=11, category_emb_size=8, hid_dim=32, eta=0.2)
    assert isinstance(m.noise_scheduler, B200PNDMScheduler) and m.noise_scheduler.config.steps_offset == 1
    assert isinstance(B200DiFashion.from_pretrained(root, scheduler="ddim", hid_dim=32).noise_scheduler, B200DDIMScheduler)
    # conv_in surgery: 8 input channels, the pretrained 4 copied, the history half zero (difashion.py:82-93)
    w = m.unet.conv_in.weight
    assert w.shape[1] == 8 and m.unet.config.in_channels == 8 and m.unet.conv_in.out_channels == 64
    assert torch.equal(w[:, :4], unet0.conv_in.weight) and float(w[:, 4:].abs().max()) == 0.0
    sd0 = unet0.state_dict()
    assert all(torch.equal(v, sd0[k]) for k, v in m.unet.state_dict().items() if not k.startswith("conv_in."))
    assert all(torch.equal(a, b) for a, b in zip(m.vae.state_dict().values(), vae0.state_dict().values()))
    assert all(torch.equal(a, b) for a, b in zip(m.text_encoder.state_dict().values(), clip0.state_dict().values()))
    # MutualEncoder sized from the VAE / UNet configs (difashion.py:95-101)
    fe = m.fashion_encoder
    assert dict(fe.config) == dict(cate_num=11, cate_emb_size=8, latent_channels=4, latent_size=16, hid_dim=32)
    assert fe.mlp[0].weight.shape == (32, 4 * 16 * 16) and m.pipe.eta_mutual == 0.2 and m.vae_scale_factor == 8

    # save_model_hook / load_model_hook round trip (inf4eval.py:543-581)
    ckpt = str(tmp_path / "checkpoint-100")
    m.save_checkpoint(ckpt, safe_serialization=safe)
    assert sorted(os.listdir(ckpt)) == ["fashion_encoder", "unet"]
    with open(os.path.join(ckpt, "unet", "config.json")) as f:
        cfg = json.load(f)
    assert cfg["_class_name"] == "UNet2DConditionModel" and cfg["in_channels"] == 8 and cfg["block_out_channels"] == [64, 128, 128, 128]
    want_u = {k: v.clone() for k, v in m.unet.state_dict().items()}
    want_f = {k: v.clone() for k, v in fe.state_dict().items()}
    m2 = B200DiFashion.from_pretrained(root, checkpoint=None, cate_num=11, category_emb_size=8, hid_dim=32)
    with torch.no_grad():
        for p in list(m2.unet.parameters()) + list(m2.fashion_encoder.parameters()):
            p.add_(1.0)
    m2.load_checkpoint(ckpt)
    assert all(torch.equal(v, want_u[k]) for k, v in m2.unet.state_dict().items())
    assert all(torch.equal(v, want_f[k]) for k, v in m2.fashion_encoder.state_dict().items())
    m3 = B200DiFashion.from_pretrained(root, checkpoint=ckpt, cate_num=11, category_emb_size=8, hid_dim=32)
    assert all(torch.equal(v, want_u[k]) for k, v in m3.unet.state_dict().items())

    # EMA copies (--use_ema / --use_ema_fashion): EMAModel.save_pretrained writes ordinary model directories + extra config keys
    with torch.no_grad():
        for q in list(m3.unet.parameters()) + list(m3.fashion_encoder.parameters()):
            q.mul_(0.5)
    m3.unet.save_pretrained(os.path.join(ckpt, "unet_ema"), safe_serialization=safe)
    m3.fashion_encoder.save_pretrained(os.path.join(ckpt, "fashion_encoder_ema"), safe_serialization=safe)
    with open(os.path.join(ckpt, "unet_ema", "config.json")) as f:
        ecfg = json.load(f)
    ecfg.update(decay=0.9999, min_decay=0.0, optimization_step=1234, update_after_step=0, use_ema_warmup=False, inv_gamma=1.0, power=0.75)
    with open(os.path.join(ckpt, "unet_ema", "config.json"), "w") as f:
        json.dump(ecfg, f)
    m2.load_checkpoint(ckpt, use_ema=True, use_ema_fashion=True)
    assert all(torch.equal(v, 0.5 * want_u[k]) for k, v in m2.unet.state_dict().items())
    assert all(torch.equal(v, 0.5 * want_f[k]) for k, v in m2.fashion_encoder.state_dict().items())
    assert "decay" not in m2.unet.config and m2.unet.config.in_channels == 8

    # the reference ctor's signature: DiFashion(args, logger, cate_num, device) (difashion.py:52-58)
    class Args:
        pretrained_model_name_or_path, category_emb_size, hid_dim, eta = root, 8, 32, 0.1
    msgs = []

    class Log:
        def info(self, s):
            msgs.append(s)
    m4 = B200DiFashion.from_args(Args(), Log(), 11, None)
    assert m4.fashion_encoder.config.cate_num == 11 and msgs


def test_weight_file_resolution_and_errors(tmp_path):
    from difashion_b200 import B200UNet2DConditionModel, MutualEncoder
    from difashion_b200 import checkpoint as ck
    d = str(tmp_path / "fashion_encoder")
    torch.manual_seed(0)
    a = MutualEncoder(cate_num=5, cate_emb_size=4, latent_size=8, hid_dim=16)
    a.save_pretrained(d)                                        # .bin
    b = MutualEncoder(cate_num=5, cate_emb_size=4, latent_size=8, hid_dim=16)
    b.save_pretrained(d, safe_serialization=True)               # .safetensors next to it: wins, as in diffusers
    assert ck.find_weights(d)[1] is True
    got = MutualEncoder.from_pretrained(str(tmp_path), subfolder="fashion_encoder")
    assert torch.equal(got.mlp[0].weight, b.mlp[0].weight) and not torch.equal(got.mlp[0].weight, a.mlp[0].weight)
    assert list(got.state_dict().keys()) == ["category_embedding.weight", "mlp.0.weight", "mlp.0.bias", "mlp.3.weight", "mlp.3.bias"]
    got.register_to_config(hid_dim=99, _class_name="x")
    assert got.config.hid_dim == 99 and "_class_name" not in got.config
    with pytest.raises(OSError):
        MutualEncoder.from_pretrained(str(tmp_path), subfolder="nope")
    os.makedirs(str(tmp_path / "empty"))
    with pytest.raises(OSError, match="config.json"):
        B200UNet2DConditionModel.from_pretrained(str(tmp_path / "empty"))
    ck.write_config(str(tmp_path / "cfgonly"), dict(sample_size=16), "UNet2DConditionModel")
    with pytest.raises(OSError, match="no weights file"):
        B200UNet2DConditionModel.from_pretrained(str(tmp_path / "cfgonly"))
    ck.write_config(str(tmp_path / "odd"), dict(sample_size=16, dual_cross_attention=True), "UNet2DConditionModel")
    with pytest.raises(NotImplementedError, match="dual_cross_attention"):
        B200UNet2DConditionModel.from_pretrained(str(tmp_path / "odd"))
    # a state dict that does not match the config is an error, never a silent partial load
    u = B200UNet2DConditionModel(**TINY_UNET)
    u.save_pretrained(str(tmp_path / "u"))
    sd = torch.load(str(tmp_path / "u" / "diffusion_pytorch_model.bin"))
    sd.pop("conv_out.bias")
    torch.save(sd, str(tmp_path / "u" / "diffusion_pytorch_model.bin"))
    with pytest.raises(RuntimeError, match="conv_out.bias"):
        B200UNet2DConditionModel.from_pretrained(str(tmp_path / "u"))


def test_scheduler_config_file(tmp_path):
    from difashion_b200 import B200DDIMScheduler, B200PNDMScheduler
    d = str(tmp_path / "scheduler")
    os.makedirs(d)
    cfg = dict(SD_SCHED)
    with open(os.path.join(d, "scheduler_config.json"), "w") as f:
        json.dump(cfg, f)
    p = B200PNDMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    p.set_timesteps(50)
    assert len(p.timesteps) == 51 and int(p.timesteps[0]) == 981           # PLMS: n + 1 model calls
    q = B200DDIMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    q.set_timesteps(50)
    assert len(q.timesteps) == 50 and int(q.timesteps[0]) == 981 and int(q.timesteps[-1]) == 1
    assert torch.equal(p.alphas_cumprod, B200PNDMScheduler().alphas_cumprod)
    for bad in (dict(prediction_type="v_prediction"), dict(trained_betas=[0.1]), dict(timestep_spacing="trailing"),
                dict(skip_prk_steps=False)):
        with open(os.path.join(d, "scheduler_config.json"), "w") as f:
            json.dump(dict(cfg, **bad), f)
        with pytest.raises(NotImplementedError):
            B200PNDMScheduler.from_pretrained(d)


def test_history_latent_format(tmp_path):
    """data_utils.py:114-147: all_item_latents.npy cache + hist_latents[uid][cate] = mean of the interacted items' latents,
    hist_latents["null"] = the white image's latent; what fashion_generation's `history` argument holds."""
    import numpy as np
    from difashion_b200 import B200AutoencoderKL, build_history_latents, encode_all_item_latents
    g = torch.Generator().manual_seed(0)
    all_latents = torch.randn(6, 4, 8, 8, generator=g)
    np.save(str(tmp_path / "all_item_latents.npy"), np.array(all_latents))           # the reference's own cache file
    vae = B200AutoencoderKL(**TINY_VAE)                                               # on the host: the cache must make it unnecessary
    got = encode_all_item_latents(vae, img_dataset=None, data_path=str(tmp_path))
    assert torch.equal(got, all_latents)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        encode_all_item_latents(vae, img_dataset=[torch.zeros(3, 64, 64)], data_path=None)
    history = {7: {3: [1, 2, 5], 9: [4]}, 8: {3: [0]}}
    hist = build_history_latents(history, all_latents)
    assert set(hist) == {7, 8, "null"} and torch.equal(hist["null"], all_latents[0])
    assert torch.allclose(hist[7][3], (all_latents[1] + all_latents[2] + all_latents[5]) / 3) and torch.equal(hist[7][9], all_latents[4])
    assert hist[8][3].shape == (4, 8, 8)


def test_category_prompts_and_tokenisation():
    """data_utils.py:88-111: prompt template ("a pair of" for pants / earrings) and the [outfits, olen, 77] id tensor."""
    from types import SimpleNamespace
    from difashion_b200 import category_prompt, tokenize_categories
    assert category_prompt("dress") == "A photo of a dress, on white background, high quality"
    assert category_prompt("skinny pants") == "A photo of a pair of skinny pants, on white background, high quality"
    assert category_prompt("hoop earrings").startswith("A photo of a pair of hoop earrings")

    class Tok:                                   # stands in for CLIPTokenizer: one id per character, padded to model_max_length
        model_max_length = 77

        def __call__(self, prompts, max_length, padding, truncation, return_tensors):
            assert padding == "max_length" and truncation and return_tensors == "pt" and max_length == 77
            ids = torch.zeros(len(prompts), max_length, dtype=torch.long)
            for i, p in enumerate(prompts):
                t = torch.tensor([ord(c) for c in p][:max_length])
                ids[i, :len(t)] = t
            return SimpleNamespace(input_ids=ids)
    ids = tokenize_categories(Tok(), [[1, 2, 3, 4], [4, 4, 1, 2]], {1: "dress", 2: "pants", 3: "bag", 4: "earrings"})
    assert ids.shape == (2, 4, 77) and ids.dtype == torch.long
    assert torch.equal(ids[0, 0], ids[1, 2]) and torch.equal(ids[0, 3], ids[1, 0]) and not torch.equal(ids[0, 0], ids[0, 1])
    assert "".join(chr(int(c)) for c in ids[0, 1] if c) == category_prompt("pants")[:77]
    tr = pytest.importorskip("transformers")     # the real tokenizer class accepts the same call (no vocabulary files offline)
    assert hasattr(tr, "CLIPTokenizer")
