"""On-disk model layout of the reference: read / write the directories DiFashion loads and saves.

The reference assembles its model from a Stable-Diffusion directory (``DiFashion/models/difashion.py:64-79``:
``scheduler/``, ``text_encoder/``, ``vae/``, ``unet/`` of ``--pretrained_model_name_or_path``) and restores a fine-tuned
checkpoint from ``<ckpt>/unet`` and ``<ckpt>/fashion_encoder`` (``DiFashion/inf4eval.py:556-581``, written by
``save_model_hook`` ``:543-554`` through diffusers' ``save_pretrained``).  A diffusers 0.18.2 model directory is
``config.json`` + ``diffusion_pytorch_model.safetensors`` (preferred when present) or ``diffusion_pytorch_model.bin``;
a transformers one is ``config.json`` + ``model.safetensors`` or ``pytorch_model.bin``.  Host-side only: no kernels here.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional, Sequence, Tuple

import torch

DIFFUSERS_STEM = "diffusion_pytorch_model"
TRANSFORMERS_STEMS = ("model", "pytorch_model")


def model_dir(path: str, subfolder: Optional[str] = None) -> str:
    d = os.path.join(path, subfolder) if subfolder else path
    if not os.path.isdir(d):
        raise OSError(f"{d} is not a directory (model hubs are unreachable here: pass a local path)")
    return d


def read_config(d: str, name: str = "config.json") -> Dict[str, Any]:
    """``config.json`` without the ``_class_name`` / ``_diffusers_version`` bookkeeping keys; lists become tuples."""
    p = os.path.join(d, name)
    if not os.path.isfile(p):
        raise OSError(f"no {name} in {d}")
    with open(p) as f:
        cfg = json.load(f)
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if not k.startswith("_")}


def write_config(d: str, cfg: Dict[str, Any], class_name: str, name: str = "config.json") -> None:
    os.makedirs(d, exist_ok=True)
    out = {"_class_name": class_name}
    out.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()})
    with open(os.path.join(d, name), "w") as f:
        json.dump(out, f, indent=2, sort_keys=True)


def find_weights(d: str, stems: Sequence[str] = (DIFFUSERS_STEM,)) -> Tuple[str, bool]:
    """(file, is_safetensors): ``<stem>.safetensors`` wins over ``<stem>.bin`` as in diffusers / transformers."""
    for stem in stems:
        p = os.path.join(d, stem + ".safetensors")
        if os.path.isfile(p):
            return p, True
    for stem in stems:
        p = os.path.join(d, stem + ".bin")
        if os.path.isfile(p):
            return p, False
    raise OSError(f"no weights file ({' | '.join(stems)}).(safetensors | bin) in {d}")


def read_state_dict(d: str, stems: Sequence[str] = (DIFFUSERS_STEM,)) -> Dict[str, torch.Tensor]:
    p, safe = find_weights(d, stems)
    if safe:
        from safetensors.torch import load_file
        return load_file(p, device="cpu")
    sd = torch.load(p, map_location="cpu", weights_only=True)
    if not isinstance(sd, dict):
        raise RuntimeError(f"{p} does not hold a state dict")
    return sd


def write_state_dict(d: str, sd: Dict[str, torch.Tensor], stem: str = DIFFUSERS_STEM, safe_serialization: bool = False) -> str:
    os.makedirs(d, exist_ok=True)
    sd = {k: v.detach().cpu().contiguous() for k, v in sd.items()}
    if safe_serialization:
        from safetensors.torch import save_file
        p = os.path.join(d, stem + ".safetensors")
        save_file(sd, p, metadata={"format": "pt"})
    else:
        p = os.path.join(d, stem + ".bin")
        torch.save(sd, p)
    return p


def widen_conv_in(unet, in_channels: int = 8):
    """DiFashion's surgery on the pretrained UNet (``difashion.py:82-93``): ``conv_in`` gets ``in_channels`` inputs,
    the pretrained 4 are copied, the new (history-latent) channels start at ZERO; bias is re-initialised by ``nn.Conv2d``
    in the reference (a fresh module) — the checkpoint restored afterwards overwrites it."""
    import torch.nn as nn
    old = unet.conv_in
    if old.weight.shape[1] == in_channels:
        return unet
    if old.weight.shape[1] > in_channels:
        raise ValueError(f"conv_in already has {old.weight.shape[1]} input channels")
    unet.register_to_config(in_channels=in_channels)
    with torch.no_grad():
        new = nn.Conv2d(in_channels, old.out_channels, old.kernel_size, old.stride, old.padding)
        new.weight.zero_()
        new.weight[:, :old.weight.shape[1]].copy_(old.weight)
        new.weight.requires_grad_(False)
        new.bias.requires_grad_(False)
        unet.conv_in = new
    return unet
