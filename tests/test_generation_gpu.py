"""The whole reference method: ``B200DiFashion.fashion_generation`` (prompt encoding -> VAE encode -> history lookup ->
denoising loop -> VAE decode -> uint8 images -> result dictionary -> inf4eval's on-disk layout) against
``oracle.generation_oracle.oracle_fashion_generation`` (DiFashion/models/difashion.py:277-616) on identical random-init
weights, ids, images and noise.  Small configurations of every model so that the CPU oracle finishes in seconds."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _build(precision, sched_name="ddim", **kw):
    from difashion_b200 import (B200AutoencoderKL, B200CLIPTextModel, B200DDIMScheduler, B200DiFashion, B200PNDMScheduler,
                                B200UNet2DConditionModel, MutualEncoder)
    from oracle.clip_oracle import make_oracle_clip, tiny_clip_config
    from oracle.generation_oracle import make_oracle_mutual_encoder
    from oracle.schedulers_oracle import OracleDDIMScheduler, OraclePNDMScheduler
    from oracle.unet_oracle import make_oracle_unet, tiny_config
    from oracle.vae_oracle import make_oracle_vae, tiny_vae_config
    ucfg = tiny_config()                                                    # sample_size 16, cross_attention_dim 64
    ccfg = tiny_clip_config(hidden_size=ucfg.cross_attention_dim, intermediate_size=128)
    vcfg = tiny_vae_config(block_out_channels=(64, 64, 128, 128))           # 128x128 images <-> 16x16 latents
    o = dict(unet=make_oracle_unet(ucfg, seed=0), me=make_oracle_mutual_encoder(seed=1, latent_size=ucfg.sample_size, hid_dim=64),
             clip=make_oracle_clip(ccfg, seed=2), vae=make_oracle_vae(vcfg, seed=3, with_encoder=True),
             sched=OracleDDIMScheduler() if sched_name == "ddim" else OraclePNDMScheduler())
    unet = B200UNet2DConditionModel(sample_size=ucfg.sample_size, in_channels=ucfg.in_channels, out_channels=ucfg.out_channels,
                                    block_out_channels=tuple(ucfg.block_out_channels), cross_attention_dim=ucfg.cross_attention_dim,
                                    attention_head_dim=ucfg.attention_head_dim)
    unet.load_state_dict(o["unet"].state_dict())
    me = MutualEncoder(latent_size=ucfg.sample_size, hid_dim=64)
    me.load_state_dict(o["me"].state_dict())
    clip = B200CLIPTextModel(vocab_size=ccfg.vocab_size, hidden_size=ccfg.hidden_size, intermediate_size=ccfg.intermediate_size,
                             num_hidden_layers=ccfg.num_hidden_layers, num_attention_heads=ccfg.num_attention_heads,
                             bos_token_id=ccfg.vocab_size - 2, eos_token_id=ccfg.vocab_size - 1, pad_token_id=ccfg.vocab_size - 1)
    clip.load_transformers_state_dict(o["clip"].state_dict())
    vae = B200AutoencoderKL(block_out_channels=tuple(vcfg.block_out_channels), layers_per_block=vcfg.layers_per_block,
                            norm_num_groups=vcfg.norm_num_groups, scaling_factor=vcfg.scaling_factor)
    vae.load_diffusers_state_dict(o["vae"].state_dict())
    for m in (unet, clip, vae):
        m.cuda()
        m.set_precision(precision)
    me.cuda()
    sched = B200DDIMScheduler() if sched_name == "ddim" else B200PNDMScheduler()
    return o, B200DiFashion(unet, vae, clip, me, sched, eta=0.1, **kw), ucfg, ccfg


def _inputs(ucfg, ccfg, olists, seed=31):
    g = torch.Generator().manual_seed(seed)
    bsz, olen = olists.shape
    n = int((olists == 0).sum())
    input_ids = torch.randint(0, ccfg.vocab_size - 2, (bsz, olen, 77), generator=g)
    input_ids[..., 0] = ccfg.vocab_size - 2
    input_ids[..., 9:] = ccfg.vocab_size - 1
    category = torch.randint(1, 6, (bsz, olen), generator=g)
    s = ucfg.sample_size
    uids, oids = torch.arange(bsz) % 2 + 40, torch.arange(bsz) + 900           # two users, distinct outfits
    history = {40: {c: 0.9 * torch.randn(4, s, s, generator=g) for c in (1, 2, 3)}, 41: {4: 0.9 * torch.randn(4, s, s, generator=g)}}
    return dict(uids=uids, oids=oids, input_ids=input_ids, olists=olists, category=category, history=history,
                outfit_images=torch.randn(bsz * olen, 3, 8 * s, 8 * s, generator=g).clamp(-1, 1),
                null_img=torch.ones(3, 8 * s, 8 * s), init_latents=torch.randn(n, 4, s, s, generator=g))


# the oracle swaps the CLIP null ids for the tiny vocabulary
def _oracle_run(o, inp, steps, scales, **kw):
    import oracle.clip_oracle as co
    from oracle.generation_oracle import oracle_fashion_generation
    rec = {}
    vocab = o["clip"].cfg.vocab_size
    orig = co.null_input_ids

    def tiny_null(max_length=77):
        ids = torch.full((1, max_length), vocab - 1, dtype=torch.long)
        ids[0, 0] = vocab - 2
        return ids
    co.null_input_ids = tiny_null
    try:
        res, init = oracle_fashion_generation(o["unet"], o["me"], o["sched"], o["clip"], o["vae"], **inp, num_inference_steps=steps,
                                              category_guidance_scale=scales[0], hist_guidance_scale=scales[1],
                                              mutual_guidance_scale=scales[2], record=rec, **kw)
    finally:
        co.null_input_ids = orig
    return res, rec


@pytest.mark.parametrize("precision,task,sched", [("fp32", "FITB", "ddim"), ("bf16", "FITB", "ddim"), ("bf16", "GOR", "pndm"), ("fp32", "GOR", "ddim")])
def test_fashion_generation_matches_oracle(precision, task, sched, tmp_path):
    from difashion_b200 import save_batch_outputs, save_outputs_npy
    # GOR runs the reference's history lookup (never matches: null latents), FITB the integer-key intent (history rows active)
    int_keys = task == "FITB"
    o, model, ucfg, ccfg = _build(precision, sched, reference_history_lookup=not int_keys)
    olists = torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0]]) if task == "FITB" else torch.zeros(2, 4, dtype=torch.long)
    inp = _inputs(ucfg, ccfg, olists)
    steps, scales = 6, (12.0, 4.0, 5.0)
    ref, rec = _oracle_run(o, inp, steps, scales, history_int_keys=int_keys)
    got, init = model.fashion_generation(**inp, num_inference_steps=steps, category_guidance_scale=scales[0],
                                         hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2], output_type="uint8",
                                         return_dict=False)
    lat = model.fashion_generation(**inp, num_inference_steps=steps, category_guidance_scale=scales[0], hist_guidance_scale=scales[1],
                                   mutual_guidance_scale=scales[2], output_type="latent")[0].images
    torch.cuda.synchronize()
    e_lat = rel_l2(lat.cpu(), rec["latents"])
    assert torch.equal(init.cpu(), inp["init_latents"])
    assert sorted(got) == sorted(ref) and all(sorted(got[u]) == sorted(ref[u]) for u in ref)
    diffs = []
    for u in ref:
        for oid in ref[u]:
            a, b = got[u][oid], ref[u][oid]
            assert len(a["images"]) == len(b["images"]) and torch.equal(a["full_cates"], b["full_cates"])
            assert [int(c) for c in a["cates"]] == [int(c) for c in b["cates"]] and torch.equal(a["outfits"], b["outfits"])
            for x, y in zip(a["images"], b["images"]):
                assert x.shape == y.shape == (8 * ucfg.sample_size, 8 * ucfg.sample_size, 3) and x.dtype == np.uint8
                diffs.append(np.abs(x.astype(np.int16) - y.astype(np.int16)))
    d = np.stack(diffs)
    print(f"\n[fashion_generation {precision} {task} {sched}] final-latent rel-L2 {e_lat:.3e}; uint8 images: max |diff| {d.max()}, "
          f"mean {d.mean():.4f}, differing pixels {float((d != 0).mean()):.4%}")
    if precision == "fp32":
        assert e_lat <= 1e-4 and d.max() <= 1 and float((d != 0).mean()) < 2e-3
    else:
        assert e_lat <= 5e-2 and d.mean() < 3.0
    # on-disk layout of inf4eval.save_batch_outputs
    gen = str(tmp_path / f"{task}-checkpoint-1-cate12.0-mutual5.0-hist4.0")
    all_out, _ = save_batch_outputs({}, {}, got, gen, task, save_grd=False)
    back = np.load(save_outputs_npy(gen, all_out), allow_pickle=True).item()
    for u in ref:
        for oid in ref[u]:
            assert len(back[u][oid]["image_paths"]) == len(ref[u][oid]["images"]) and all(os.path.exists(p) for p in back[u][oid]["image_paths"])
            assert os.path.exists(os.path.join(gen, "images", str(u), str(oid), "all.jpg")) == (task == "GOR")


def test_fashion_generation_history_lookup_and_prompt_cache():
    """`reference_history_lookup=True` (default) reproduces the reference's tensor-keyed membership test, which never matches
    (every item gets the null latent); False uses history rows for (uid, category) pairs present in `history` (integer keys);
    prompts are encoded once."""
    o, model, ucfg, ccfg = _build("fp32")
    assert model.reference_history_lookup is True
    model.reference_history_lookup = False
    olists = torch.tensor([[0, 0, 7, 9], [0, 5, 0, 2]])
    inp = _inputs(ucfg, ccfg, olists)
    inp["category"] = torch.tensor([[1, 2, 5, 5], [4, 5, 5, 5]])              # (uid 40: 1, 2) and (uid 41: 4) are in `history`
    kw = dict(num_inference_steps=3, category_guidance_scale=12.0, hist_guidance_scale=4.0, mutual_guidance_scale=5.0, output_type="latent")
    a = model.fashion_generation(**inp, **kw)[0].images.clone()
    n_cached = len(model._prompt_cache)
    b = model.fashion_generation(**inp, **kw)[0].images.clone()
    assert torch.equal(a, b) and len(model._prompt_cache) == n_cached            # deterministic; no re-encoding
    _, rec = _oracle_run(o, inp, 3, (12.0, 4.0, 5.0), history_int_keys=True)
    assert rel_l2(a.cpu(), rec["latents"]) <= 1e-4
    model.reference_history_lookup = True
    c = model.fashion_generation(**inp, **kw)[0].images
    _, rec_ref = _oracle_run(o, inp, 3, (12.0, 4.0, 5.0))                       # the literal restatement: lookup never matches
    _, rec_nohist = _oracle_run(o, dict(inp, history={}), 3, (12.0, 4.0, 5.0))
    assert torch.equal(rec_ref["latents"], rec_nohist["latents"])
    assert rel_l2(c.cpu(), rec_ref["latents"]) <= 1e-4 and rel_l2(c.cpu(), rec["latents"]) > 1e-3
