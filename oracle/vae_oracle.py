"""Pure-PyTorch fp32 restatement of diffusers 0.18.2 ``AutoencoderKL`` (SD-1.5 VAE): decoder, and (``with_encoder``)
the encoder half used before the loop — ``vae.encode(images).latent_dist.mode() * scaling_factor`` for the given items
(``DiFashion/models/difashion.py:435-437``), the white ``null_img`` (``:375-376``) and the history images (``:129-144``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) — parity unpinned: diffusers is absent.

Reference call site: ``image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]``
(``DiFashion/models/difashion.py:579``; SURVEY.md §8f row 1 — the step right after the denoising loop).

Restated from memory of the pinned release (SD-1.5 ``vae/config.json``: ``block_out_channels (128, 256, 512,
512)``, ``layers_per_block 2``, ``latent_channels 4``, ``norm_num_groups 32``, ``act_fn silu``,
``scaling_factor 0.18215``):

* ``decode(z) = Decoder(post_quant_conv(z))``; ``post_quant_conv = Conv2d(4, 4, 1)``;
* ``Decoder``: ``conv_in`` 3x3 4->512; ``mid_block`` = ``UNetMidBlock2D`` (resnet, single-head attention with
  GroupNorm(32, eps 1e-6) / biased q,k,v / residual, resnet); 4 ``UpDecoderBlock2D`` (3 resnets each,
  channels 512, 512, 256, 128, nearest-2x ``Upsample2D`` + 3x3 conv after the first three);
  ``conv_norm_out`` GroupNorm(32, 128, eps 1e-6) -> SiLU -> ``conv_out`` 3x3 128->3;
* ``ResnetBlock2D`` without time embedding (``temb_channels=None``), eps 1e-6, 1x1 ``conv_shortcut`` when
  the channel count changes.

Pinned by builder-made checks only (``tests/test_oracle_cpu.py``): decoder parameter count 49 490 179 + 20
for ``post_quant_conv`` (the published SD VAE total 83 653 863 = encoder 34 163 592 + quant_conv 72 + these),
138 + 2 state-dict tensors, diffusers key names.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    sample_size: int = 512


def tiny_vae_config(**kw) -> VAEConfig:
    base = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=1, norm_num_groups=8, sample_size=128)
    base.update(kw)
    return VAEConfig(**base)


class VAEResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VAEAttention(nn.Module):
    """diffusers 0.18.2 ``Attention`` as built by ``UNetMidBlock2D`` for the VAE (one head of ``channels`` dims,
    biased projections, GroupNorm, residual connection, ``rescale_output_factor`` 1)."""

    def __init__(self, channels: int, groups: int, eps: float = 1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])
        self.scale = channels ** -0.5

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        y = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)           # [B, HW, C]
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * self.scale, dim=-1)
        o = self.to_out[1](self.to_out[0](torch.matmul(p, v)))
        return o.transpose(1, 2).reshape(b, c, h, w) + res


class VAEUpsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self, resnets, attentions=None, upsamplers=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attentions is not None:
            self.attentions = nn.ModuleList(attentions)
        if upsamplers is not None:
            self.upsamplers = nn.ModuleList(upsamplers)


class VAEDecoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = tuple(cfg.block_out_channels), cfg.norm_num_groups
        top = boc[-1]
        self.conv_in = nn.Conv2d(cfg.latent_channels, top, 3, padding=1)
        self.mid_block = _Block([VAEResnetBlock2D(top, top, g), VAEResnetBlock2D(top, top, g)], [VAEAttention(top, g)])
        ups, prev = [], top
        rev = list(reversed(boc))
        for i, c in enumerate(rev):
            res = [VAEResnetBlock2D(prev if j == 0 else c, c, g) for j in range(cfg.layers_per_block + 1)]
            ups.append(_Block(res, upsamplers=[VAEUpsample2D(c)] if i < len(rev) - 1 else None))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        m = self.mid_block
        h = m.resnets[1](m.attentions[0](m.resnets[0](h)))
        for blk in self.up_blocks:
            for r in blk.resnets:
                h = r(h)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VAEDownsample2D(nn.Module):
    """diffusers ``Downsample2D(use_conv=True, padding=0)`` as built by ``DownEncoderBlock2D``: asymmetric zero pad
    (right / bottom only) then a stride-2, pad-0 3x3 conv."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class _DownBlock(nn.Module):
    def __init__(self, resnets, downsamplers=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if downsamplers is not None:
            self.downsamplers = nn.ModuleList(downsamplers)


class VAEEncoder(nn.Module):
    """diffusers 0.18.2 ``Encoder`` (``double_z=True``): ``conv_in`` 3x3 3->128; 4 ``DownEncoderBlock2D`` (2 resnets each,
    channels 128, 256, 512, 512, ``Downsample2D`` after the first three); the same ``UNetMidBlock2D`` as the decoder;
    ``conv_norm_out`` GroupNorm(32, 512, eps 1e-6) -> SiLU -> ``conv_out`` 3x3 512 -> 2*latent_channels."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = tuple(cfg.block_out_channels), cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        downs, prev = [], boc[0]
        for i, c in enumerate(boc):
            res = [VAEResnetBlock2D(prev if j == 0 else c, c, g) for j in range(cfg.layers_per_block)]
            downs.append(_DownBlock(res, [VAEDownsample2D(c)] if i < len(boc) - 1 else None))
            prev = c
        self.down_blocks = nn.ModuleList(downs)
        top = boc[-1]
        self.mid_block = _Block([VAEResnetBlock2D(top, top, g), VAEResnetBlock2D(top, top, g)], [VAEAttention(top, g)])
        self.conv_norm_out = nn.GroupNorm(g, top, eps=1e-6)
        self.conv_out = nn.Conv2d(top, 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            for r in blk.resnets:
                h = r(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
        m = self.mid_block
        h = m.resnets[1](m.attentions[0](m.resnets[0](h)))
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class OracleVAE(nn.Module):
    """``AutoencoderKL`` (state-dict keys ``post_quant_conv.*`` / ``decoder.*`` and, ``with_encoder=True``,
    ``quant_conv.*`` / ``encoder.*`` as in diffusers)."""

    def __init__(self, cfg: VAEConfig = None, with_encoder: bool = False):
        super().__init__()
        self.cfg = cfg or VAEConfig()
        if with_encoder:
            self.encoder = VAEEncoder(self.cfg)
            self.quant_conv = nn.Conv2d(2 * self.cfg.latent_channels, 2 * self.cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(self.cfg.latent_channels, self.cfg.latent_channels, 1)
        self.decoder = VAEDecoder(self.cfg)

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor):
        """``AutoencoderKL.encode(x).latent_dist`` -> (mean, logvar): ``moments = quant_conv(encoder(x))``,
        ``mean, logvar = chunk(moments, 2, dim=1)``, ``logvar = clamp(logvar, -30, 20)``
        (``DiagonalGaussianDistribution``)."""
        mean, logvar = torch.chunk(self.quant_conv(self.encoder(x)), 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    @torch.no_grad()
    def encode_mode_scaled(self, x: torch.Tensor) -> torch.Tensor:
        """The reference call ``vae.encode(x).latent_dist.mode() * vae.config.scaling_factor``
        (difashion.py:129-130, :375-376, :435-437): the mode of the diagonal Gaussian is its mean."""
        return self.encode_moments(x)[0] * self.cfg.scaling_factor

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.post_quant_conv(z))

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """The reference call ``vae.decode(latents / scaling_factor)`` (difashion.py:579)."""
        return self.decode(latents / self.cfg.scaling_factor)


def make_oracle_vae(cfg: VAEConfig = None, seed: int = 0, with_encoder: bool = False) -> OracleVAE:
    torch.manual_seed(seed)
    m = OracleVAE(cfg, with_encoder=with_encoder).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m
