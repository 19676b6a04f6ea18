"""Pure-PyTorch fp32 restatement of diffusers 0.18.2 ``UNet2DConditionModel`` (SD-1.5 shape).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) — parity unpinned: diffusers is absent.

What it restates (reference call site ``DiFashion/models/difashion.py:518-523``; model
assembly / 8-channel ``conv_in`` surgery ``difashion.py:82-93``):

* module tree and state-dict key names of diffusers' ``UNet2DConditionModel`` so a diffusers
  checkpoint loads with ``load_state_dict``;
* forward numerics of App. A.2: fp32 sinusoidal timestep projection ``[cos | sin]``,
  ``ResnetBlock2D``, ``Transformer2DModel`` (1x1-conv projections, or Linear when
  ``use_linear_projection``), ``BasicTransformerBlock`` with exact-erf GEGLU, nearest-2x
  ``Upsample2D``, stride-2 ``Downsample2D``.

Everything is plain ``torch.nn`` on CPU; no dependency on the product package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 8            # DiFashion widens 4 -> 8 (difashion.py:83-93)
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = (
        "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = (
        "UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: object = 8   # diffusers 0.18.2 uses this as the NUMBER of heads
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0

    def heads(self, i: int) -> int:
        a = self.attention_head_dim
        return a[i] if isinstance(a, (list, tuple)) else a


def tiny_config(**kw) -> UNetConfig:
    """A structurally identical but small UNet for fast CPU/GPU parity tests."""
    base = dict(sample_size=16, in_channels=8, out_channels=4,
                block_out_channels=(64, 128, 128, 128), cross_attention_dim=64,
                attention_head_dim=2)
    base.update(kw)
    return UNetConfig(**base)


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True,
                       freq_shift: float = 0.0, max_period: int = 10000) -> torch.Tensor:
    """diffusers ``get_timestep_embedding`` (App. A.2 item 2): fp32, returns [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, s, _ = x.shape
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        h = self.heads

        def split(t):
            return t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3)

        q, k, v = split(q), split(k), split(v)
        # plain AttnProcessor math: softmax(q k^T * scale) v   (App. A.2 "Attention")
        w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * self.scale, dim=-1)
        o = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(b, s, -1)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)          # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, cross_dim, groups, use_linear_projection=False):
        super().__init__()
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        if use_linear_projection:
            self.proj_in = nn.Linear(channels, channels)
            self.proj_out = nn.Linear(channels, channels)
        else:
            self.proj_in = nn.Conv2d(channels, channels, 1)
            self.proj_out = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, channels // heads, cross_dim)])

    def forward(self, x, ctx):
        b, c, hh, ww = x.shape
        res = x
        h = self.norm(x)
        if not self.use_linear_projection:
            h = self.proj_in(h)
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        else:
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if not self.use_linear_projection:
            h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
            h = self.proj_out(h)
        else:
            h = self.proj_out(h)
            h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, i: int, cin: int, cout: int, temb_dim: int):
        super().__init__()
        has_attn = cfg.down_block_types[i] == "CrossAttnDownBlock2D"
        final = i == len(cfg.block_out_channels) - 1
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if j == 0 else cout, cout, temb_dim, cfg.norm_num_groups, cfg.norm_eps)
            for j in range(cfg.layers_per_block)])
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, cfg.heads(i), cfg.cross_attention_dim, cfg.norm_num_groups,
                                   cfg.use_linear_projection)
                for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.downsamplers = None if final else nn.ModuleList([Downsample2D(cout)])

    def forward(self, h, temb, ctx):
        outs = []
        for j, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, temb_dim: int):
        super().__init__()
        c = cfg.block_out_channels[-1]
        self.resnets = nn.ModuleList([
            ResnetBlock2D(c, c, temb_dim, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(c, cfg.heads(len(cfg.block_out_channels) - 1), cfg.cross_attention_dim,
                               cfg.norm_num_groups, cfg.use_linear_projection)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, i: int, temb_dim: int):
        super().__init__()
        rev = list(reversed(cfg.block_out_channels))
        n = len(rev)
        out_c = rev[i]
        prev_c = rev[max(i - 1, 0)]
        in_c = rev[min(i + 1, n - 1)]
        has_attn = cfg.up_block_types[i] == "CrossAttnUpBlock2D"
        nl = cfg.layers_per_block + 1
        res = []
        for j in range(nl):
            skip_c = in_c if j == nl - 1 else out_c
            rin = prev_c if j == 0 else out_c
            res.append(ResnetBlock2D(rin + skip_c, out_c, temb_dim, cfg.norm_num_groups, cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        rev_heads = n - 1 - i
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(out_c, cfg.heads(rev_heads), cfg.cross_attention_dim,
                                   cfg.norm_num_groups, cfg.use_linear_projection)
                for _ in range(nl)])
        else:
            self.attentions = None
        self.upsamplers = None if i == n - 1 else nn.ModuleList([Upsample2D(out_c)])

    def forward(self, h, skips: List[torch.Tensor], temb, ctx):
        for j, r in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = r(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class OracleUNet2DConditionModel(nn.Module):
    """diffusers ``UNet2DConditionModel`` restated; ``forward`` returns the sample tensor."""

    def __init__(self, cfg: Optional[UNetConfig] = None):
        super().__init__()
        cfg = cfg or UNetConfig()
        self.cfg = cfg
        boc = cfg.block_out_channels
        temb_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        downs, cin = [], boc[0]
        for i, c in enumerate(boc):
            downs.append(DownBlock(cfg, i, cin, c, temb_dim))
            cin = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, temb_dim)
        self.up_blocks = nn.ModuleList([UpBlock(cfg, i, temb_dim) for i in range(len(boc))])
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def time_embed(self, timestep, batch: int) -> torch.Tensor:
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32 if isinstance(t, float) else torch.int64)
        elif t.dim() == 0:
            t = t[None]
        t = t.expand(batch)
        t_emb = timestep_embedding(t, self.cfg.block_out_channels[0], self.cfg.flip_sin_to_cos,
                                   self.cfg.freq_shift)
        return self.time_embedding(t_emb)

    def forward(self, sample, timestep, encoder_hidden_states, taps: Optional[dict] = None):
        """``taps`` (optional dict) receives per-block intermediates for error localisation."""
        temb = self.time_embed(timestep, sample.shape[0])
        h = self.conv_in(sample)
        if taps is not None:
            taps["conv_in"] = h
        skips = [h]
        for i, blk in enumerate(self.down_blocks):
            h, outs = blk(h, temb, encoder_hidden_states)
            skips.extend(outs)
            if taps is not None:
                taps[f"down{i}"] = h
        h = self.mid_block(h, temb, encoder_hidden_states)
        if taps is not None:
            taps["mid"] = h
        for i, blk in enumerate(self.up_blocks):
            h = blk(h, skips, temb, encoder_hidden_states)
            if taps is not None:
                taps[f"up{i}"] = h
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return h


def make_oracle_unet(cfg: Optional[UNetConfig] = None, seed: int = 0) -> OracleUNet2DConditionModel:
    """Random-init (PyTorch defaults, like ``UNet2DConditionModel(**config)``), fixed seed."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = OracleUNet2DConditionModel(cfg).eval()
    torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m
