"""The caller-side data format of the history condition (``DiFashion/data_utils.py:114-147``).

The reference precomputes the VAE latent of every catalogue item once (``all_item_latents.npy``: ``vae.encode(imgs)
.latent_dist.mode() * vae.config.scaling_factor`` in batches of 64, ``data_utils.py:114-135``), then builds the history
dictionary ``hist_latents[uid][category] = mean of the latents of the items the user interacted with in that category``
plus ``hist_latents["null"] = all_latents[0]`` (the white image, ``:137-147``); ``fashion_generation`` receives it as its
``history`` argument (``inf4eval.py:742``, ``difashion.py:378-386``).  Here the encoding runs on ``B200AutoencoderKL``'s
kernels; the dictionary arithmetic is the reference's own host-side mean over a handful of tensors.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional, Sequence

import numpy as np
import torch

ALL_LATENTS_FILE = "all_item_latents.npy"


@torch.no_grad()
def encode_all_item_latents(vae, img_dataset, data_path: Optional[str] = None, batch_size: int = 64,
                            device=None) -> torch.Tensor:
    """``all_latents`` fp32 ``[n_items, 4, h, w]`` on the host: loaded from ``<data_path>/all_item_latents.npy`` when that file
    exists (the reference's cache, same format: ``np.save`` of the array), else encoded batch by batch with
    ``vae.encode_latents`` (mode of the posterior times the scaling factor) and saved there."""
    path = os.path.join(data_path, ALL_LATENTS_FILE) if data_path else None
    if path and os.path.exists(path):
        return torch.tensor(np.load(path, allow_pickle=False))
    device = torch.device(device) if device is not None else vae.device
    if device.type != "cuda":
        raise RuntimeError("encode_all_item_latents needs the VAE on a CUDA device: there is no CPU fallback")
    out = []
    n = len(img_dataset)
    for start in range(0, n, batch_size):
        imgs = torch.stack([img_dataset[i] for i in range(start, min(start + batch_size, n))], dim=0)
        imgs = imgs.to(memory_format=torch.contiguous_format).float().to(device)
        out.append(vae.encode_latents(imgs).float().cpu())
    all_latents = torch.cat(out, dim=0)
    if path:
        np.save(path, np.array(all_latents))
    return all_latents


def build_history_latents(history: Mapping[int, Mapping[int, Sequence[int]]], all_latents: torch.Tensor) -> Dict:
    """``hist_latents[uid][cate] = all_latents[iids].mean(0)`` and ``hist_latents["null"] = all_latents[0]``
    (``data_utils.py:137-147``): the ``history`` argument of ``fashion_generation``."""
    hist: Dict = {}
    for uid in history:
        hist.setdefault(uid, {})
        for cate in history[uid]:
            iids = torch.as_tensor(list(history[uid][cate]), dtype=torch.long)
            hist[uid][cate] = all_latents[iids].mean(dim=0)
    hist["null"] = all_latents[0]
    return hist
