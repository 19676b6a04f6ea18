"""Restatement of DiFashion's ``MutualEncoder`` and of the ``fashion_generation`` denoising loop.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows ``DiFashion/models/difashion.py``:
* ``MutualEncoder``                 ``:21-46``  (Linear -> LeakyReLU(0.01) -> Dropout -> Linear -> Tanh)
* xavier-normal init               ``:731-746``
* condition-flag logic             ``:309-325``
* history / prompt CFG layouts     ``:388-431``
* ``mutual_indicies`` bookkeeping  ``:439-451``
* loop body                        ``:456-577`` (latent expansion, mutual sum of the *other*
  slots with weight 1.0 — NOT the mean —, MLP, CFG layout, blend with ``args.eta``,
  history concat, UNet, 4/3/2-branch CFG combine, scheduler step, prev_latents hand-off)

``oracle_generation`` is the hot path (the loop): it takes the VAE / CLIP outputs (latents, prompt embeddings) as
inputs.  ``oracle_fashion_generation`` restates the whole method with the stages around the loop:
CLIP prompt encoding ``:339-353``, null-latent / given-item VAE encode ``:375-376``, ``:435-437``, history lookup
``:378-386``, VAE decode + ``VaeImageProcessor.postprocess`` ``:579-592``, result dictionary ``:598-614``.
"""
from __future__ import annotations

import inspect
from typing import Optional

import torch
import torch.nn as nn


class OracleMutualEncoder(nn.Module):
    def __init__(self, latent_channels=4, latent_size=64, hid_dim=256, cate_num=50, cate_emb_size=64):
        super().__init__()
        self.category_embedding = nn.Embedding(cate_num, cate_emb_size)   # unused in the reference too
        self.latent_channels, self.latent_size = latent_channels, latent_size
        d = latent_channels * latent_size * latent_size
        self.mlp = nn.Sequential(nn.Linear(d, hid_dim), nn.LeakyReLU(), nn.Dropout(0.1),
                                 nn.Linear(hid_dim, d), nn.Tanh())

    def forward(self, x):
        b = x.shape[0]
        return self.mlp(x.reshape(b, -1)).reshape(b, self.latent_channels, self.latent_size, self.latent_size)


def make_oracle_mutual_encoder(seed: int = 1, **kw) -> OracleMutualEncoder:
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = OracleMutualEncoder(**kw)
    for mod in m.modules():                        # difashion.py:731-746
        if isinstance(mod, nn.Embedding):
            nn.init.xavier_normal_(mod.weight)
        elif isinstance(mod, nn.Linear):
            nn.init.xavier_normal_(mod.weight)
            nn.init.constant_(mod.bias, 0.0)
    torch.random.set_rng_state(st)
    for p in m.parameters():
        p.requires_grad_(False)
    return m.eval()


def guidance_flags(use_history, use_mutual, s_cate, s_hist, s_mutual):
    """difashion.py:309-325."""
    do_h = bool(use_history and s_hist > 1.0)
    do_m = bool(use_mutual and s_mutual > 1.0)
    do_c = bool(s_cate > 1.0)
    return do_h, do_m, do_c, (do_h and do_m and do_c)


def mutual_indices(olists: torch.Tensor) -> torch.Tensor:
    """difashion.py:439-451: >=0 -> row of ``all_latents``; negative -> -(row of generated latents)-1."""
    bsz, olen = olists.shape
    gen_masks = olists == 0
    rows, n = [], 0
    for i in range(bsz):
        idx = torch.arange(olen) + i * olen
        g = int(gen_masks[i].sum())
        idx[gen_masks[i]] = -torch.arange(n, n + g) - 1
        rows.append(idx)
        n += g
    return torch.stack(rows)


@torch.no_grad()
def oracle_generation(unet, mutual_encoder, scheduler, *, olists, all_latents, category_prompts,
                      null_prompt, hist_latents, null_latent, init_latents, num_inference_steps=50,
                      category_guidance_scale=12.0, hist_guidance_scale=4.0, mutual_guidance_scale=5.0,
                      eta_mutual=0.1, use_history=True, use_mutual_guidance=True, ddim_eta=0.0,
                      generator=None, record: Optional[list] = None, max_steps: Optional[int] = None):
    """Returns final latents [N,4,h,w].  ``record`` (list) receives per-step dicts
    {t, unet_in, noise_pred_branches, noise_pred, latents} for per-step parity checks."""
    do_h, do_m, do_c, do_all = guidance_flags(use_history, use_mutual_guidance, category_guidance_scale,
                                              hist_guidance_scale, mutual_guidance_scale)
    bsz, olen = olists.shape
    fill_idx = torch.nonzero(olists == 0)
    n = fill_idx.shape[0]
    gen_masks = olists == 0
    null_prompts = torch.cat([null_prompt] * n, dim=0)

    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    latents = init_latents.clone() * scheduler.init_noise_sigma

    null_stack = torch.stack([null_latent] * n)
    if not use_history:                      # difashion.py:381-385: without history every item gets the null latent
        hist_latents = null_stack
    if do_all:
        hist = torch.cat([hist_latents, null_stack, null_stack, null_stack], 0)
        ctx = torch.cat([category_prompts, category_prompts, category_prompts, null_prompts], 0)
        nb = 4
    elif do_c:
        if do_h:
            hist = torch.cat([hist_latents, null_stack, null_stack], 0)
            ctx = torch.cat([category_prompts, category_prompts, null_prompts], 0)
            nb = 3
        elif do_m:
            hist = torch.cat([hist_latents] * 3, 0)
            ctx = torch.cat([category_prompts, category_prompts, null_prompts], 0)
            nb = 3
        else:
            hist = torch.cat([hist_latents] * 2, 0)
            ctx = torch.cat([category_prompts, null_prompts], 0)
            nb = 2
    else:
        if do_h:
            hist = torch.cat([hist_latents, null_stack], 0)
            ctx = torch.cat([category_prompts] * 2, 0)
            nb = 2
        elif do_m:
            hist = torch.cat([hist_latents] * 2, 0)
            ctx = torch.cat([category_prompts] * 2, 0)
            nb = 2
        else:
            hist, ctx, nb = hist_latents, category_prompts, 1

    step_params = set(inspect.signature(scheduler.step).parameters.keys())
    extra = {}
    if "eta" in step_params:
        extra["eta"] = ddim_eta
    if "generator" in step_params:
        extra["generator"] = generator

    mi = mutual_indices(olists)
    prev_latents = latents.clone()

    for i, t in enumerate(timesteps):
        if max_steps is not None and i >= max_steps:
            break
        x = scheduler.scale_model_input(torch.cat([latents] * nb), t)

        if use_mutual_guidance:
            conds = []
            for (o_idx, i_idx) in fill_idx.tolist():
                w = torch.ones(olen)
                w[i_idx] = 0.0
                slots = torch.zeros((olen,) + tuple(null_latent.shape), dtype=null_latent.dtype)
                g = gen_masks[o_idx]
                slots[~g] = all_latents[mi[o_idx][~g]]
                slots[g] = prev_latents[-mi[o_idx][g] - 1]
                conds.append(sum(wk * s for wk, s in zip(w, slots)))
            m = mutual_encoder(torch.stack(conds))
        else:
            m = null_stack.clone()

        if do_all:
            m = torch.cat([m, m, null_stack, null_stack], 0)
        elif do_c:
            if do_m:
                m = torch.cat([m, null_stack, null_stack], 0)
            elif do_h:
                m = torch.cat([m] * 3, 0)
            else:
                m = torch.cat([m] * 2, 0)
        else:
            if do_m:
                m = torch.cat([m, null_stack], 0)
            elif do_h:
                m = torch.cat([m] * 2, 0)

        x = (1 - eta_mutual) * x + eta_mutual * m
        unet_in = torch.cat([x, hist], dim=1)
        eps_b = unet(unet_in, t, ctx)

        s_h, s_m, s_c = hist_guidance_scale, mutual_guidance_scale, category_guidance_scale
        if do_all:
            e0, e1, e2, e3 = eps_b.chunk(4)
            eps = e3 + s_h * (e0 - e1) + s_m * (e1 - e2) + s_c * (e2 - e3)
        elif do_c:
            if do_h:
                e0, e1, e2 = eps_b.chunk(3)
                eps = e2 + s_h * (e0 - e1) + s_c * (e1 - e2)
            elif do_m:
                e0, e1, e2 = eps_b.chunk(3)
                eps = e2 + s_m * (e0 - e1) + s_c * (e1 - e2)
            else:
                e0, e1 = eps_b.chunk(2)
                eps = e1 + s_c * (e0 - e1)
        else:
            if do_h:
                e0, e1 = eps_b.chunk(2)
                eps = e1 + s_h * (e0 - e1)
            elif do_m:
                e0, e1 = eps_b.chunk(2)
                eps = e1 + s_m * (e0 - e1)
            else:
                eps = eps_b

        latents = scheduler.step(eps, t, latents, **extra, return_dict=False)[0]
        prev_latents = latents
        if record is not None:
            record.append(dict(t=int(t), unet_in=unet_in, noise_pred_branches=eps_b, noise_pred=eps,
                               latents=latents.clone()))
    return latents


def postprocess_uint8(image: torch.Tensor):
    """diffusers ``VaeImageProcessor.postprocess(image, output_type="np"/"pil")`` with ``do_denormalize`` all True
    (difashion.py:586-592): ``(image / 2 + 0.5).clamp(0, 1)`` -> NHWC float32 numpy -> ``(x * 255).round().astype("uint8")``."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.cpu().permute(0, 2, 3, 1).float().numpy()
    return (x * 255).round().astype("uint8")


@torch.no_grad()
def oracle_fashion_generation(unet, mutual_encoder, scheduler, text_encoder, vae, *, uids, oids, input_ids, olists,
                              outfit_images, category, history, null_img, init_latents, num_inference_steps=50,
                              category_guidance_scale=7.5, hist_guidance_scale=7.5, mutual_guidance_scale=7.5,
                              eta_mutual=0.1, use_history=True, use_mutual_guidance=True, ddim_eta=0.0, generator=None,
                              history_int_keys=False,
                              max_steps: Optional[int] = None, record: Optional[dict] = None):
    """``DiFashion.fashion_generation(..., return_dict=False)`` (difashion.py:277-616) on the CPU oracles.

    uids/oids [bsz], input_ids [bsz, olen, 77], olists [bsz, olen] (0 = blank), outfit_images [bsz*olen, 3, H, W],
    category [bsz, olen], history {uid: {cate: latent[4,h,w]}}, null_img [3, H, W], init_latents [N, 4, h, w].
    Returns (all_results, init_latents) with images as uint8 HWC numpy arrays (what ``output_type="pil"`` wraps)."""
    fill_idx = torch.nonzero(olists == 0)                                   # :332-337
    fill_cate = category[fill_idx[:, 0], fill_idx[:, 1]]
    fill_uids, fill_oids = uids[fill_idx[:, 0]], oids[fill_idx[:, 0]]
    full_cate = category[fill_idx[:, 0]]
    fill_input_ids = input_ids[fill_idx[:, 0], fill_idx[:, 1]]
    category_prompts = text_encoder(fill_input_ids)[0]                      # :339-341
    from .clip_oracle import null_input_ids
    null_prompt = text_encoder(null_input_ids(category_prompts.shape[1]))[0]      # :343-352
    null_latent = vae.encode_mode_scaled(null_img.unsqueeze(0))[0]          # :375-376
    hist = []                                                               # :378-386
    # difashion.py:380-382 iterates the TENSOR ``fill_cate`` and tests ``cate in history[uid]`` with the 0-d tensor: tensors
    # hash by identity, so the membership test never matches an integer key and every item gets the null latent.  Restated
    # literally (default); ``history_int_keys=True`` is the evident intent (integer category keys), kept for the product's
    # ``reference_history_lookup=False`` option.
    for i, cate in enumerate(fill_cate):
        uid = int(uids[fill_idx[i][0]])
        if history_int_keys:
            cate = int(cate)
        if use_history and cate in history.get(uid, {}):
            hist.append(history[uid][cate])
        else:
            hist.append(null_latent)
    hist_latents = torch.stack(hist)
    all_latents = vae.encode_mode_scaled(outfit_images)                      # :435-437
    latents = oracle_generation(unet, mutual_encoder, scheduler, olists=olists, all_latents=all_latents,
                                category_prompts=category_prompts, null_prompt=null_prompt, hist_latents=hist_latents,
                                null_latent=null_latent, init_latents=init_latents, num_inference_steps=num_inference_steps,
                                category_guidance_scale=category_guidance_scale, hist_guidance_scale=hist_guidance_scale,
                                mutual_guidance_scale=mutual_guidance_scale, eta_mutual=eta_mutual, use_history=use_history,
                                use_mutual_guidance=use_mutual_guidance, ddim_eta=ddim_eta, generator=generator,
                                max_steps=max_steps)
    image = vae.decode_latents(latents)                                      # :579
    images = postprocess_uint8(image)                                        # :586-592
    if record is not None:
        record.update(category_prompts=category_prompts, null_prompt=null_prompt, null_latent=null_latent,
                      hist_latents=hist_latents, all_latents=all_latents, latents=latents, image=image)
    all_results = {}                                                         # :598-614
    for i, uid in enumerate(fill_uids.tolist()):
        oid = int(fill_oids[i])
        ent = all_results.setdefault(uid, {}).setdefault(oid, dict(images=[], cates=[], full_cates=full_cate[i]))
        ent["images"].append(images[i])
        ent["cates"].append(fill_cate[i])
        ent["outfits"] = olists[fill_idx[i][0]]
    return all_results, init_latents
