"""``Attention`` module (diffusers names) and the B200 attention processor.

``B200AttnProcessor`` implements the diffusers-0.18.2 attention-processor protocol
(``proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None)``; the
plug-in point xformers uses, reference ``DiFashion/models/difashion.py:109-118``) on the tcgen05
GEMM + flash-attention kernels.  It works with this package's ``Attention`` *and* with a diffusers
``Attention`` module (same attribute names: ``to_q/to_k/to_v/to_out/heads/scale``).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


def pack_head_rows(w: torch.Tensor, heads: int, dp: int) -> torch.Tensor:
    """``[heads*d, K]`` projection weight -> ``[heads*dp, K]`` with zero rows padding each head."""
    n, k = w.shape
    d = n // heads
    out = torch.zeros(heads, dp, k, dtype=w.dtype, device=w.device)
    out[:, :d] = w.reshape(heads, d, k)
    return out.reshape(heads * dp, k)


def pack_head_cols(w: torch.Tensor, heads: int, dp: int) -> torch.Tensor:
    """``[N, heads*d]`` output-projection weight -> ``[N, heads*dp]`` (zero columns in the padding)."""
    n, k = w.shape
    d = k // heads
    out = torch.zeros(n, heads, dp, dtype=w.dtype, device=w.device)
    out[:, :, :d] = w.reshape(n, heads, d)
    return out.reshape(n, heads * dp)


class AttnPack:
    """Pre-packed bf16 weights of one Attention layer (padded-head layout)."""

    def __init__(self, attn, device, dtype: torch.dtype = torch.bfloat16):
        heads = attn.heads
        inner = attn.to_q.weight.shape[0]
        d = inner // heads
        self.heads, self.d, self.dp = heads, d, ops.pad16(d)
        self.cp = heads * self.dp
        self.scale = float(getattr(attn, "scale", d ** -0.5))
        f = lambda t: t.detach().to(device=device, dtype=torch.float32)
        wq, wk, wv = f(attn.to_q.weight), f(attn.to_k.weight), f(attn.to_v.weight)
        self.is_cross = wk.shape[1] != wq.shape[1] or getattr(attn, "is_cross_attention", False)
        hp = lambda w: pack_head_rows(w, heads, self.dp)
        self.w_q = ops.pack_linear(hp(wq), dtype)
        self.w_kv = ops.pack_linear(torch.cat([hp(wk), hp(wv)], 0), dtype)
        self.w_qkv = ops.pack_linear(torch.cat([hp(wq), hp(wk), hp(wv)], 0), dtype) if wk.shape[1] == wq.shape[1] else None
        # Self-attention with head padding (d = 40 -> 48): the fused q|k|v projection gets a bias that is 1.0 in the first
        # padding column of every head of V and 0 elsewhere (to_q / to_k / to_v have no bias of their own), so V carries a
        # ones column and the P V MMA of the long-sequence kernel accumulates the softmax denominator (ops.attention
        # ``ones_col``).  The padding columns of Q / K stay zero (scores unchanged) and to_out's packed weight has zero
        # columns there (the 1.0 the attention writes into that output column is multiplied by zero).
        self.ones_col = d if (self.w_qkv is not None and self.dp > d and dtype == torch.bfloat16) else None
        self.b_qkv = None
        if self.ones_col is not None:
            bq = torch.zeros(3, heads, self.dp, dtype=torch.float32, device=device)
            bq[2, :, d] = 1.0
            self.b_qkv = bq.reshape(-1).contiguous()
        self.w_o = ops.pack_linear(pack_head_cols(f(attn.to_out[0].weight), heads, self.dp), dtype)
        b = attn.to_out[0].bias
        self.b_o = f(b).contiguous() if b is not None else None
        self.c_out = attn.to_out[0].weight.shape[0]
        self.c_in = wq.shape[1]
        self.c_ctx = wk.shape[1]
        self.key = _weights_key(attn)


def _weights_key(attn):
    ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight)
    return tuple((w.data_ptr(), w._version) for w in ws)


class B200AttnProcessor:
    """softmax(QK^T/sqrt(d))V on tcgen05 (see ``dfb_attention`` in include/dfb200.h)."""

    def _pack(self, attn, device) -> AttnPack:
        pk = getattr(attn, "_dfb_pack", None)
        if pk is None or pk.key != _weights_key(attn) or pk.w_q.device != device:
            pk = AttnPack(attn, device)
            attn._dfb_pack = pk
        return pk

    @torch.no_grad()
    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None, **kw):
        if attention_mask is not None:
            raise NotImplementedError("B200AttnProcessor: attention_mask is not used on the DiFashion path")
        if not hidden_states.is_cuda:
            raise RuntimeError("B200AttnProcessor needs CUDA tensors (there is no CPU fallback)")
        pk = self._pack(attn, hidden_states.device)
        b, s, c = hidden_states.shape
        x = hidden_states.to(torch.bfloat16).contiguous()
        dev = x.device
        if encoder_hidden_states is None:
            qkv = torch.empty(b, s, 3 * pk.cp, dtype=torch.bfloat16, device=dev)
            ops.gemm([x.view(b * s, c)], pk.w_qkv, 3 * pk.cp, out=qkv.view(b * s, 3 * pk.cp), bias=pk.b_qkv)
            q, k, v = qkv[..., :pk.cp], qkv[..., pk.cp:2 * pk.cp], qkv[..., 2 * pk.cp:]
            ones_col = pk.ones_col
        else:
            ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
            skv = ctx.shape[1]
            q = torch.empty(b, s, pk.cp, dtype=torch.bfloat16, device=dev)
            ops.gemm([x.view(b * s, c)], pk.w_q, pk.cp, out=q.view(b * s, pk.cp))
            kv = torch.empty(b, skv, 2 * pk.cp, dtype=torch.bfloat16, device=dev)
            ops.gemm([ctx.view(b * skv, ctx.shape[2])], pk.w_kv, 2 * pk.cp, out=kv.view(b * skv, 2 * pk.cp))
            k, v = kv[..., :pk.cp], kv[..., pk.cp:]
            ones_col = None
        o = torch.empty(b, s, pk.cp, dtype=torch.bfloat16, device=dev)
        ops.attention(q, k, v, o, heads=pk.heads, dp=pk.dp, scale=pk.scale, ones_col=ones_col)
        out = torch.empty(b, s, pk.c_out, dtype=torch.float32, device=dev)
        ops.gemm([o.view(b * s, pk.cp)], pk.w_o, pk.c_out, out=out.view(b * s, pk.c_out), bias=pk.b_o)
        return out.to(hidden_states.dtype)        # to_out[1] is Dropout(0.0): identity


class Attention(nn.Module):
    """Parameter container with diffusers' ``Attention`` attribute names."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.is_cross_attention = cross_attention_dim is not None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = B200AttnProcessor()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
