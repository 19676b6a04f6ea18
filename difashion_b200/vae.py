"""``B200AutoencoderKL`` — diffusers 0.18.2 ``AutoencoderKL`` (SD-1.5 VAE) on the hand-written sm_100a kernels.

``decode``: the step right after the denoising loop,
``image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False)[0]``
(``DiFashion/models/difashion.py:579``).
``encode``: the step right before it, ``vae.encode(images).latent_dist.mode() * vae.config.scaling_factor`` for the
given items of an outfit (``:435-437``), the white ``null_img`` (``:375-376``) and the history images
(``data_utils.py:132``) — SURVEY.md §8f row 3.  The encoder's ``Downsample2D`` (zero pad right / bottom, stride-2
pad-0 3x3 conv) is the implicit-GEMM conv over space-to-depth planes with its own tap table (the pad is the TMA
out-of-bounds fill); ``quant_conv`` (1x1) is folded into ``conv_out``'s packed weights in fp32.

Same conventions as ``unet.py``: diffusers module tree / state-dict keys (``post_quant_conv.*``, ``decoder.*``, ``quant_conv.*``, ``encoder.*``; a decoder-only
state dict loads too, ``encode`` then raises), plain
``nn`` modules as parameter containers (``encoder.*`` / ``quant_conv.*`` too), NHWC activations, bf16 tensor-core operands with an fp32 residual
stream (or fp32 operands on the verification path), no PyTorch arithmetic, no CPU fallback.

* 3x3 convolutions (up to 512x512x128) run on the implicit-GEMM tcgen05 kernel (tiles of 128 pixels of one
  image row when the image is wider than 128), GroupNorm statistics come from the producing epilogue;
* the single-head d=512 mid-block attention (one layer, 4096 tokens) exceeds the flash kernel's TMEM budget
  (2 x 64 score columns + 512 output columns > 512), so it runs per image as GEMMs around a row-softmax
  kernel: S = Q K^T, P = softmax(S / sqrt(C)), O = P V with V^T produced directly by a GEMM whose "weight"
  operand is the normalised activation; the value bias is folded into the output projection
  (rows of P sum to 1).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .config import FrozenConfig
from .unet import Workspace, _f32

SD15_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, act_fn="silu", sample_size=512, scaling_factor=0.18215)


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class DiagonalGaussianDistribution:
    """diffusers ``DiagonalGaussianDistribution`` over the encoder moments: ``mode()`` (= mean) is what inference uses
    (difashion.py:376, :437); ``sample()`` (training, :144, out of scope) is host glue on the [B, 4, h, w] latents."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor):
        self.mean, self.logvar = mean, logvar

    @property
    def std(self):
        return torch.exp(0.5 * self.logvar)

    @property
    def var(self):
        return torch.exp(self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator=None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


class _Resnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class _Attention(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])


class _Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _Downsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)


class _Block(nn.Module):
    def __init__(self, resnets, attentions=None, upsamplers=None, downsamplers=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attentions is not None:
            self.attentions = nn.ModuleList(attentions)
        if upsamplers is not None:
            self.upsamplers = nn.ModuleList(upsamplers)
        if downsamplers is not None:
            self.downsamplers = nn.ModuleList(downsamplers)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = tuple(cfg["block_out_channels"]), cfg["norm_num_groups"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        downs, prev = [], boc[0]
        for i, c in enumerate(boc):
            res = [_Resnet(prev if j == 0 else c, c, g) for j in range(cfg["layers_per_block"])]
            downs.append(_Block(res, downsamplers=[_Downsample(c)] if i < len(boc) - 1 else None))
            prev = c
        self.down_blocks = nn.ModuleList(downs)
        top = boc[-1]
        self.mid_block = _Block([_Resnet(top, top, g), _Resnet(top, top, g)], [_Attention(top, g)])
        self.conv_norm_out = nn.GroupNorm(g, top, eps=1e-6)
        self.conv_out = nn.Conv2d(top, 2 * cfg["latent_channels"], 3, padding=1)


class _Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, g = tuple(cfg["block_out_channels"]), cfg["norm_num_groups"]
        top = boc[-1]
        self.conv_in = nn.Conv2d(cfg["latent_channels"], top, 3, padding=1)
        self.mid_block = _Block([_Resnet(top, top, g), _Resnet(top, top, g)], [_Attention(top, g)])
        ups, prev, rev = [], top, list(reversed(boc))
        for i, c in enumerate(rev):
            res = [_Resnet(prev if j == 0 else c, c, g) for j in range(cfg["layers_per_block"] + 1)]
            ups.append(_Block(res, upsamplers=[_Upsample(c)] if i < len(rev) - 1 else None))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)


_DEPRECATED_ATTN_KEYS = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


class B200AutoencoderKL(nn.Module):
    def __init__(self, **config):
        super().__init__()
        cfg = dict(SD15_VAE_CONFIG)
        cfg.update(config)
        self._config = FrozenConfig(cfg)
        self.encoder = _Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)
        self.post_quant_conv = nn.Conv2d(cfg["latent_channels"], cfg["latent_channels"], 1)
        self.decoder = _Decoder(cfg)
        self._encoder_loaded = True        # False after loading a decoder-only state dict
        self._pack_enc: Optional[Dict[str, Any]] = None
        self._pack_enc_key = None
        for p in self.parameters():
            p.requires_grad_(False)
        self._op_dtype = torch.bfloat16
        self._pack: Optional[Dict[str, Any]] = None
        self._pack_key = None
        self._ws: Dict[Any, Workspace] = {}
        self.max_images = 16            # images per pass (bounds the workspace: ~0.55 GB per image at 512x512)

    @property
    def config(self) -> FrozenConfig:
        return self._config

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def set_precision(self, precision: str):
        """``"bf16"`` (tensor cores) or ``"fp32"`` (verification path on the CUDA cores); see the UNet."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._op_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        return self

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw):
        """``AutoencoderKL.from_pretrained(dir, subfolder="vae")`` (difashion.py:74): ``config.json`` +
        ``diffusion_pytorch_model.(safetensors | bin)``; old attention key names are accepted."""
        from . import checkpoint as ck
        d = ck.model_dir(path, subfolder)
        cfg = ck.read_config(d)
        if cfg.get("act_fn", "silu") != "silu":
            raise NotImplementedError(f"VAE act_fn={cfg['act_fn']!r}")
        m = cls(**{k: v for k, v in cfg.items() if k in SD15_VAE_CONFIG})
        m.load_diffusers_state_dict(ck.read_state_dict(d))
        return m

    def save_pretrained(self, save_directory: str, safe_serialization: bool = False, **kw):
        from . import checkpoint as ck
        ck.write_config(save_directory, dict(self._config), "AutoencoderKL")
        ck.write_state_dict(save_directory, self.state_dict(), ck.DIFFUSERS_STEM, safe_serialization)

    def load_diffusers_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Load a diffusers ``AutoencoderKL`` state dict (the pre-0.18 attention names query/key/value/proj_attn are
        mapped to to_q/to_k/to_v/to_out.0).  A decoder-only dict (no ``encoder.*`` keys) loads the decode half and
        leaves ``encode`` unavailable."""
        out = {}
        has_enc = any(k.startswith("encoder.") for k in sd)
        for k, v in sd.items():
            if not (k.startswith(("decoder.", "post_quant_conv.")) or (has_enc and k.startswith(("encoder.", "quant_conv.")))):
                continue
            parts = k.split(".")
            if "attentions" in parts:
                parts = [p for p in ".".join(_DEPRECATED_ATTN_KEYS.get(p, p) for p in parts).split(".")]
                if v.dim() == 4 and parts[-1] == "weight" and parts[-2] in ("to_q", "to_k", "to_v", "0"):
                    v = v.reshape(v.shape[0], v.shape[1])        # very old checkpoints stored 1x1 convs
            out[".".join(parts)] = v
        if has_enc:
            res = self.load_state_dict(out, strict=True)
        else:
            res = self.load_state_dict(out, strict=False)
            bad = [k for k in res.missing_keys if not k.startswith(("encoder.", "quant_conv."))] + list(res.unexpected_keys)
            if bad:
                raise RuntimeError(f"AutoencoderKL state dict mismatch: {bad[:8]}")
        self._encoder_loaded = has_enc
        return res

    # ---------------------------------------------------------------- packing
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    @staticmethod
    def _pack_resnet(rb: _Resnet, device, dt):
        r = dict(cin=rb.conv1.weight.shape[1], cout=rb.conv1.weight.shape[0])
        r["n1"] = (_f32(rb.norm1.weight, device), _f32(rb.norm1.bias, device), rb.norm1.eps, rb.norm1.num_groups)
        r["n2"] = (_f32(rb.norm2.weight, device), _f32(rb.norm2.bias, device), rb.norm2.eps, rb.norm2.num_groups)
        r["w1"], r["b1"] = ops.pack_conv3x3(rb.conv1.weight.to(device), dt), _f32(rb.conv1.bias, device)
        w2, b2 = ops.pack_conv3x3(rb.conv2.weight.to(device), dt), _f32(rb.conv2.bias, device)
        r["shortcut"] = rb.conv_shortcut is not None
        if r["shortcut"]:
            w2 = torch.cat([w2, ops.pack_linear(rb.conv_shortcut.weight.to(device), dt)], dim=1).contiguous()
            b2 = (b2 + _f32(rb.conv_shortcut.bias, device)).contiguous()
        r["w2"], r["b2"] = w2, b2
        return r

    @staticmethod
    def _pack_attention(at: _Attention, device, dt):
        c = at.to_q.weight.shape[0]
        wo, bo = at.to_out[0].weight.detach().to(device).float(), at.to_out[0].bias.detach().to(device).float()
        bv = at.to_v.bias.detach().to(device).float()
        return dict(
            c=c, gn=(_f32(at.group_norm.weight, device), _f32(at.group_norm.bias, device), at.group_norm.eps, at.group_norm.num_groups),
            wq=ops.pack_linear(at.to_q.weight.to(device), dt), bq=_f32(at.to_q.bias, device),
            wk=ops.pack_linear(at.to_k.weight.to(device), dt), bk=_f32(at.to_k.bias, device),
            wv_act=at.to_v.weight.detach().to(device=device, dtype=dt).contiguous(),       # A operand of the V^T GEMM
            wo=ops.pack_linear(wo, dt), bo=(bo + wo @ bv).contiguous(), scale=float(c) ** -0.5)

    def pack_encoder(self, device=None):
        """Pack the encode half: ``quant_conv`` (1x1, 8 -> 8) is folded into ``conv_out`` in fp32
        (``W' = Wq . Wc``, ``b' = Wq bc + bq``: one conv emits the moments, no intermediate rounding)."""
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("B200AutoencoderKL needs a CUDA device: there is no CPU fallback")
        if not self._encoder_loaded:
            raise RuntimeError("this B200AutoencoderKL was loaded from a decoder-only state dict: encode() is unavailable")
        dt = self._op_dtype
        key = (str(device), str(dt), self._weights_key())
        if self._pack_enc is not None and self._pack_enc_key == key:
            return self._pack_enc
        e = self.encoder
        P: Dict[str, Any] = {}
        ic = e.conv_in.weight.shape[1]
        assert ic <= 8
        win = torch.zeros(e.conv_in.weight.shape[0], 8, 3, 3, device=device)          # 3 -> 8 input channels (zero columns)
        win[:, :ic] = e.conv_in.weight.detach().to(device).float()
        P["conv_in"] = (ops.pack_conv3x3(win, dt), _f32(e.conv_in.bias, device), e.conv_in.weight.shape[0])
        P["down"] = []
        for blk in e.down_blocks:
            ent = dict(resnets=[self._pack_resnet(r, device, dt) for r in blk.resnets], down=None)
            if hasattr(blk, "downsamplers"):
                conv = blk.downsamplers[0].conv
                c = conv.weight.shape[0]
                if c % 64:
                    raise NotImplementedError("Downsample2D needs a channel count that is a multiple of 64")
                ent["down"] = (ops.pack_conv3x3(conv.weight.to(device), dt), _f32(conv.bias, device), c)
            P["down"].append(ent)
        P["mid"] = [self._pack_resnet(r, device, dt) for r in e.mid_block.resnets]
        P["attn"] = self._pack_attention(e.mid_block.attentions[0], device, dt)
        n = e.conv_norm_out
        P["norm_out"] = (_f32(n.weight, device), _f32(n.bias, device), n.eps, n.num_groups)
        wq = self.quant_conv.weight.detach().to(device).float().reshape(self.quant_conv.weight.shape[0], -1)     # [2lc, 2lc]
        wc = e.conv_out.weight.detach().to(device).float()                                                      # [2lc, top, 3, 3]
        wfold = torch.einsum("om,mikl->oikl", wq, wc)
        bfold = wq @ e.conv_out.bias.detach().to(device).float() + self.quant_conv.bias.detach().to(device).float()
        P["conv_out"] = (ops.pack_conv3x3(wfold, dt), bfold.contiguous(), wfold.shape[0])
        # "conv_out_sf": the mean rows also carry the reference's `* vae.config.scaling_factor` (difashion.py:376, :437)
        lc, sf = self.config.latent_channels, float(self.config.scaling_factor)
        wsf, bsf = wfold.clone(), bfold.clone()
        wsf[:lc] *= sf
        bsf[:lc] *= sf
        P["conv_out_sf"] = (ops.pack_conv3x3(wsf, dt), bsf.contiguous(), wfold.shape[0])
        self._pack_enc, self._pack_enc_key = P, key
        return P

    def pack(self, device=None):
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("B200AutoencoderKL needs a CUDA device: there is no CPU fallback")
        dt = self._op_dtype
        key = (str(device), str(dt), self._weights_key())
        if self._pack is not None and self._pack_key == key:
            return self._pack
        P: Dict[str, Any] = {}
        lc = self.config.latent_channels
        assert lc <= 8
        sf = float(self.config.scaling_factor)
        # post_quant_conv (1x1) over the 8-channel padded latent operand; "pq_sf" also carries the reference's
        # `latents / scaling_factor` (difashion.py:579) so decode_latents needs no separate scaling pass
        wpq = torch.zeros(lc, 8, device=device)
        wpq[:, :lc] = self.post_quant_conv.weight.detach().to(device).float().reshape(lc, lc)
        P["pq"] = (ops.pack_linear(wpq, dt), _f32(self.post_quant_conv.bias, device))
        P["pq_sf"] = (ops.pack_linear(wpq / sf, dt), P["pq"][1])
        d = self.decoder
        win = torch.zeros(d.conv_in.weight.shape[0], 8, 3, 3, device=device)
        win[:, :lc] = d.conv_in.weight.detach().to(device).float()
        P["conv_in"] = (ops.pack_conv3x3(win, dt), _f32(d.conv_in.bias, device))

        pack_resnet = lambda rb: self._pack_resnet(rb, device, dt)
        P["attn"] = self._pack_attention(d.mid_block.attentions[0], device, dt)
        P["mid"] = [pack_resnet(r) for r in d.mid_block.resnets]
        P["up"] = []
        for blk in d.up_blocks:
            e = dict(resnets=[pack_resnet(r) for r in blk.resnets], up=None)
            if hasattr(blk, "upsamplers"):
                conv = blk.upsamplers[0].conv
                e["up"] = (ops.pack_conv3x3(conv.weight.to(device), dt), _f32(conv.bias, device), conv.weight.shape[0])
            P["up"].append(e)
        n = d.conv_norm_out
        P["norm_out"] = (_f32(n.weight, device), _f32(n.bias, device), n.eps, n.num_groups)
        oc = d.conv_out.weight.shape[0]
        assert oc <= 4
        wout = torch.zeros(4, d.conv_out.weight.shape[1], 3, 3, device=device)       # 3 -> 4 output channels (zero row)
        wout[:oc] = d.conv_out.weight.detach().to(device).float()
        bout = torch.zeros(4, device=device)
        bout[:oc] = d.conv_out.bias.detach().to(device).float()
        P["conv_out"] = (ops.pack_conv3x3(wout, dt), bout.contiguous(), oc)
        self._pack, self._pack_key = P, key
        return P

    # ---------------------------------------------------------------- kernel sequencing
    def _gnp_new(self, ws: Workspace, tag: str, out: torch.Tensor):
        hw, n = out.shape[1] * out.shape[2], out.shape[3]
        self._gnp.pop(out.data_ptr(), None)
        if hw % 32 != 0 or n % 4 != 0 or self._op_dtype != torch.bfloat16:
            return None
        part = ws.get(tag + "_gnp", ops.gn_partial_shape(out.shape[0] * hw, n), torch.float32)
        self._gnp[out.data_ptr()] = part
        return part

    def _gn(self, x, norm, silu, ws, out, raw_out=None):
        g, b, eps, groups = norm
        stats = ws.get("gn_stats", (ops.groupnorm_ws_floats(x.shape[0], groups),), torch.float32)
        ops.groupnorm(x, None, g, b, groups=groups, eps=eps, silu=silu, stats_ws=stats, out=out, raw_out=raw_out,
                      partials=(self._gnp.get(x.data_ptr()), None))

    def _resnet(self, pk, x, ws: Workspace, out_tag: str):
        B, H, W, _ = x.shape
        cin, cout, dt = pk["cin"], pk["cout"], self._op_dtype
        xn = ws.get("xn", (B, H, W, cin), dt)
        xraw = ws.get("xraw", (B, H, W, cin), dt) if pk["shortcut"] else None
        self._gn(x, pk["n1"], True, ws, xn, xraw)
        h1 = ws.get("h1", (B, H, W, cout), torch.float32)
        ops.gemm([xn], pk["w1"], cout, out=h1, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=pk["b1"],
                 gn_partial=self._gnp_new(ws, "h1", h1))
        hn = ws.get("hn", (B, H, W, cout), dt)
        self._gn(h1, pk["n2"], True, ws, hn)
        out = ws.get(out_tag, (B, H, W, cout), torch.float32)
        outp = self._gnp_new(ws, out_tag, out)
        if pk["shortcut"]:
            ops.gemm([hn, xraw], pk["w2"], cout, out=out, taps=[ops.TAPS_3X3, ops.TAP_CENTER], conv_geom=(B, H, W),
                     bias=pk["b2"], gn_partial=outp)
        else:
            ops.gemm([hn], pk["w2"], cout, out=out, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=pk["b2"], residual=x,
                     gn_partial=outp)
        return out

    def _attention(self, pk, x, ws: Workspace, out_tag: str):
        B, H, W, C = x.shape
        S, M, dt = H * W, B * H * W, self._op_dtype
        xn = ws.get("xn", (M, C), dt)
        self._gn(x, pk["gn"], False, ws, xn.view(B, H, W, C))
        q, k = ws.get("att_q", (M, C), dt), ws.get("att_k", (M, C), dt)
        ops.gemm([xn], pk["wq"], C, out=q, bias=pk["bq"])
        ops.gemm([xn], pk["wk"], C, out=k, bias=pk["bk"])
        o = ws.get("att_o", (M, C), dt)
        scores = ws.get("att_s", (S, S), torch.float32)
        prob = ws.get("att_p", (S, S), dt)
        vt = ws.get("att_vt", (C, S), dt)
        for b in range(B):
            rows = slice(b * S, (b + 1) * S)
            ops.gemm([pk["wv_act"]], xn[rows], S, out=vt)                       # V^T = Wv xn^T   (bias folded into to_out)
            ops.gemm([q[rows]], k[rows], S, out=scores)                        # S = Q K^T
            ops.softmax_rows(scores, prob, pk["scale"])
            ops.gemm([prob], vt, C, out=o[rows])                               # O = P V
        out = ws.get(out_tag, (B, H, W, C), torch.float32)
        ops.gemm([o], pk["wo"], C, out=out.view(M, C), bias=pk["bo"], residual=x.view(M, C),
                 gn_partial=self._gnp_new(ws, out_tag, out))
        return out

    def _decode_chunk(self, z: torch.Tensor, ws: Workspace, pq: str = "pq") -> torch.Tensor:
        P = self.pack(z.device)
        dt = self._op_dtype
        B, lc, H, W = z.shape
        self._gnp = {}
        z8 = ws.get("z8", (B, 8, H, W), torch.float32)
        z8.zero_()
        z8[:, :lc].copy_(z)
        x0 = ws.get("x0", (B, H, W, 8), dt)
        ops.nchw_to_nhwc(z8, x0)
        x1 = ws.get("x1", (B, H, W, 8), dt)
        x1.zero_()
        ops.gemm([x0.view(B * H * W, 8)], P[pq][0], lc, out=x1.view(B * H * W, 8)[:, :lc], bias=P[pq][1])
        top = P["mid"][0]["cin"]
        h = ws.get("h_in", (B, H, W, top), torch.float32)
        ops.gemm([x1], P["conv_in"][0], top, out=h, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=P["conv_in"][1],
                 gn_partial=self._gnp_new(ws, "h_in", h))
        h = self._resnet(P["mid"][0], h, ws, "mid_r0")
        h = self._attention(P["attn"], h, ws, "mid_a")
        h = self._resnet(P["mid"][1], h, ws, "mid_r1")
        par = 0
        for e in P["up"]:
            for rp in e["resnets"]:
                par ^= 1
                h = self._resnet(rp, h, ws, f"up_r{par}")
            if e["up"] is not None:
                w, b, c = e["up"]
                Bh, Hh, Wh = h.shape[0], h.shape[1], h.shape[2]
                up = ws.get("upx", (Bh, 2 * Hh, 2 * Wh, c), dt)
                ops.upsample2x(h, up)
                h = ws.get("up_conv", (Bh, 2 * Hh, 2 * Wh, c), torch.float32)
                ops.gemm([up], w, c, out=h, taps=[ops.TAPS_3X3], conv_geom=(Bh, 2 * Hh, 2 * Wh), bias=b,
                         gn_partial=self._gnp_new(ws, "up_conv", h))
        Bh, Hh, Wh, c0 = h.shape
        xn = ws.get("xn", (Bh, Hh, Wh, c0), dt)
        self._gn(h, P["norm_out"], True, ws, xn)
        img = ws.get("img", (Bh, Hh, Wh, 4), torch.float32)
        ops.gemm([xn], P["conv_out"][0], 4, out=img, taps=[ops.TAPS_3X3], conv_geom=(Bh, Hh, Wh), bias=P["conv_out"][1])
        return img

    # ---------------------------------------------------------------- encode
    def _encode_chunk(self, x: torch.Tensor, ws: Workspace, co: str = "conv_out") -> torch.Tensor:
        """x fp32 NCHW [B, 3, H, W] -> moments fp32 NHWC [B, H/8, W/8, 2*latent_channels]."""
        P = self.pack_encoder(x.device)
        dt = self._op_dtype
        B, ic, H, W = x.shape
        self._gnp = {}
        x8 = ws.get("enc_x8", (B, 8, H, W), torch.float32)
        x8.zero_()
        x8[:, :ic].copy_(x)
        x0 = ws.get("enc_x0", (B, H, W, 8), dt)
        ops.nchw_to_nhwc(x8, x0)
        w, b, c0 = P["conv_in"]
        h = ws.get("enc_in", (B, H, W, c0), torch.float32)
        ops.gemm([x0], w, c0, out=h, taps=[ops.TAPS_3X3], conv_geom=(B, H, W), bias=b, gn_partial=self._gnp_new(ws, "enc_in", h))
        par = 0
        for ent in P["down"]:
            for rp in ent["resnets"]:
                par ^= 1
                h = self._resnet(rp, h, ws, f"enc_r{par}")
            if ent["down"] is not None:
                w, b, c = ent["down"]
                Bh, Hh, Wh = h.shape[0], h.shape[1], h.shape[2]
                if Hh % 2 or Wh % 2:
                    raise ValueError("AutoencoderKL.encode needs image sides that are multiples of 8")
                s2d = ws.get("enc_s2d", (Bh, Hh // 2, Wh // 2, 4 * c), dt)
                ops.space_to_depth(h, s2d)
                h = ws.get("enc_down", (Bh, Hh // 2, Wh // 2, c), torch.float32)
                ops.gemm([s2d], w, c, out=h, taps=[ops.s2d_taps_pad0(c)], a_c=[c], conv_geom=(Bh, Hh // 2, Wh // 2), bias=b,
                         gn_partial=self._gnp_new(ws, "enc_down", h))
        h = self._resnet(P["mid"][0], h, ws, "mid_r0")
        h = self._attention(P["attn"], h, ws, "mid_a")
        h = self._resnet(P["mid"][1], h, ws, "mid_r1")
        Bh, Hh, Wh, top = h.shape
        xn = ws.get("xn", (Bh, Hh, Wh, top), dt)
        self._gn(h, P["norm_out"], True, ws, xn)
        w, b, nm = P[co]
        mom = ws.get("enc_mom", (Bh, Hh, Wh, nm), torch.float32)
        ops.gemm([xn], w, nm, out=mom, taps=[ops.TAPS_3X3], conv_geom=(Bh, Hh, Wh), bias=b)      # conv_out with quant_conv folded in
        return mom

    def _encode(self, x: torch.Tensor, co: str) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("B200AutoencoderKL needs CUDA tensors: there is no CPU fallback")
        B, ic, H, W = x.shape
        if ic != self.config.in_channels:
            raise ValueError(f"images have {ic} channels, the encoder expects {self.config.in_channels}")
        ndown = len(self.config.block_out_channels) - 1
        if H % (1 << ndown) or W % (1 << ndown):
            raise ValueError(f"image sides must be multiples of {1 << ndown}")
        xf = x.float().contiguous()
        mom = torch.empty(B, 2 * self.config.latent_channels, H >> ndown, W >> ndown, dtype=torch.float32, device=x.device)
        ws = self.workspace(("enc", min(B, self.max_images), H, W, str(self._op_dtype)), x.device)
        for b0 in range(0, B, self.max_images):
            xc = xf[b0:b0 + self.max_images]
            ops.nhwc_to_nchw(self._encode_chunk(xc, ws, co), mom[b0:b0 + xc.shape[0]])
        return mom

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """diffusers ``AutoencoderKL.encode``: image [B, 3, H, W] in [-1, 1] -> ``AutoencoderKLOutput`` whose
        ``latent_dist`` is the diagonal Gaussian over [B, 4, H/8, W/8] (``.mode()`` = mean, as the reference reads it)."""
        lc = self.config.latent_channels
        mom = self._encode(x, "conv_out")
        dist = DiagonalGaussianDistribution(mom[:, :lc].to(x.dtype), torch.clamp(mom[:, lc:], -30.0, 20.0).to(x.dtype))
        return AutoencoderKLOutput(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def encode_latents(self, images: torch.Tensor) -> torch.Tensor:
        """``vae.encode(images).latent_dist.mode() * vae.config.scaling_factor`` (difashion.py:376, :435-437,
        data_utils.py:132) -> [B, 4, H/8, W/8]; the scaling factor is folded into the packed ``conv_out`` weights."""
        return self._encode(images, "conv_out_sf")[:, :self.config.latent_channels].to(images.dtype)

    def workspace(self, key, device) -> Workspace:
        ws = self._ws.get(key)
        if ws is None or ws.device != device:
            ws = Workspace(device)
            self._ws[key] = ws
        return ws

    def _decode(self, z: torch.Tensor, pq: str) -> torch.Tensor:
        if not z.is_cuda:
            raise RuntimeError("B200AutoencoderKL needs CUDA tensors: there is no CPU fallback")
        B, lc, H, W = z.shape
        if lc != self.config.latent_channels:
            raise ValueError(f"latents have {lc} channels, the decoder expects {self.config.latent_channels}")
        oc = self.config.out_channels
        nup = len(self.config.block_out_channels) - 1
        zf = z.float().contiguous()
        out = torch.empty(B, oc, H << nup, W << nup, dtype=torch.float32, device=z.device)
        ws = self.workspace(("dec", min(B, self.max_images), H, W, str(self._op_dtype)), z.device)
        for b0 in range(0, B, self.max_images):
            zc = zf[b0:b0 + self.max_images]
            img = self._decode_chunk(zc, ws, pq)
            n = zc.shape[0]
            tmp = ws.get("img_nchw", (n, 4, H << nup, W << nup), torch.float32)
            ops.nhwc_to_nchw(img, tmp)
            out[b0:b0 + n].copy_(tmp[:, :oc])
        return out.to(z.dtype)

    @torch.no_grad()
    def decode_latents_uint8(self, latents: torch.Tensor) -> torch.Tensor:
        """``vae.decode(latents / scaling_factor)`` followed by ``VaeImageProcessor.postprocess(..., "pil")``'s
        arithmetic (difashion.py:579-592): uint8 ``[B, 8h, 8w, 3]`` (HWC RGB) on the device, written by
        ``dfb_image_to_uint8`` straight from the decoder's NHWC output (no NCHW round trip)."""
        if not latents.is_cuda:
            raise RuntimeError("B200AutoencoderKL needs CUDA tensors: there is no CPU fallback")
        B, lc, H, W = latents.shape
        if lc != self.config.latent_channels:
            raise ValueError(f"latents have {lc} channels, the decoder expects {self.config.latent_channels}")
        nup = len(self.config.block_out_channels) - 1
        zf = latents.float().contiguous()
        out = torch.empty(B, H << nup, W << nup, 3, dtype=torch.uint8, device=latents.device)
        ws = self.workspace(("dec", min(B, self.max_images), H, W, str(self._op_dtype)), latents.device)
        for b0 in range(0, B, self.max_images):
            zc = zf[b0:b0 + self.max_images]
            ops.image_to_uint8(self._decode_chunk(zc, ws, "pq_sf"), out[b0:b0 + zc.shape[0]])
        return out

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """diffusers ``AutoencoderKL.decode``: z [B, 4, h, w] (already divided by ``config.scaling_factor`` by the
        caller, as the reference does) -> image [B, 3, 8h, 8w] in z's dtype."""
        out = self._decode(z, "pq")
        return DecoderOutput(sample=out) if return_dict else (out,)

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """``vae.decode(latents / scaling_factor)`` of difashion.py:579 in one call: the division is folded into the
        packed ``post_quant_conv`` weights."""
        return self._decode(latents, "pq_sf")
