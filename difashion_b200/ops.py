"""Thin PyTorch-tensor wrappers over the C ABI (``include/dfb200.h``) + weight packing.

PyTorch is used for device memory and the current stream only; all arithmetic happens in the
hand-written kernels of ``libdfb200.so``.  Activations are NHWC (``[B, H, W, C]`` / ``[M, C]``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DTYPE_BF16, DTYPE_F32, GemmParams, check

TAPS_3X3: Tuple[Tuple[int, int, int], ...] = tuple((kh - 1, kw - 1, 0) for kh in range(3) for kw in range(3))
TAP_CENTER: Tuple[Tuple[int, int, int], ...] = ((0, 0, 0),)


_LAUNCHES = 0
PROFILE = None      # bench.py sets this to a list: every tensor-core launch is then bracketed by CUDA events


def _prof_begin():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    return e0


ALG_BYTES = None    # with PROFILE: algorithmic HBM bytes per kind (operand + result tensors of each launch counted once)


def _prof_end(e0, kind: str, flops: float, shape, nbytes: float = 0.0):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record(torch.cuda.current_stream())
    PROFILE.append((kind, flops, shape, e0, e1))
    if ALG_BYTES is not None:
        ALG_BYTES[kind] = ALG_BYTES.get(kind, 0.0) + nbytes


def launch_count() -> int:
    """Kernels of libdfb200.so launched so far by this process (bench.py's ``gpu_launches`` evidence)."""
    return _LAUNCHES


def _count(k: int = 1) -> None:
    global _LAUNCHES
    _LAUNCHES += k


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    if t.dtype == torch.float32:
        return DTYPE_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def ceil64(x: int) -> int:
    return (x + 63) // 64 * 64


# ----------------------------------------------------------------------------------------------
# weight packing (done once, on the host side of the boundary)
# ----------------------------------------------------------------------------------------------
def pack_linear(weight: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """nn.Linear / 1x1-conv weight ``[N, K(,1,1)]`` -> ``[N, ceil64(K)]`` (K-major); bf16 for the tensor-core
    path, fp32 for the fp32 verification path."""
    w = weight.detach().reshape(weight.shape[0], -1).float()
    n, k = w.shape
    out = torch.zeros(n, ceil64(k), dtype=dtype, device=w.device)
    out[:, :k] = w.to(dtype)
    return out.contiguous()


def upsample_phase_taps(a: int, b: int):
    """Tap table ``(dh, dw, channel_offset)`` of phase (a, b) of a fused nearest-2x upsample + 3x3 conv, pad 1: output pixel
    ``(2i + a, 2j + b)`` reads the upsampled rows ``2i + a + {-1, 0, 1}``, i.e. input rows ``{i - 1, i, i}`` (a = 0) or
    ``{i, i, i + 1}`` (a = 1) — two distinct rows; same for columns.  Out-of-range rows / columns are the conv's zero padding
    in both formulations (the upsampled image's border maps to the input's border)."""
    dhs = (-1, 0) if a == 0 else (0, 1)
    dws = (-1, 0) if b == 0 else (0, 1)
    return [(dh, dw, 0) for dh in dhs for dw in dws]


def pack_upsample_phases(weight: torch.Tensor, dtype: torch.dtype = torch.bfloat16):
    """OIHW 3x3 weight of ``Upsample2D.conv`` -> four packed ``[O, 4 * ceil64(I)]`` matrices, one per output phase
    (a, b) in the order (0,0), (0,1), (1,0), (1,1), for the taps of ``upsample_phase_taps(a, b)``: the 3x3 kernel rows that
    land on the same input row are summed (in fp32, before the operand rounding): phase a = 0 -> rows [w0, w1 + w2],
    a = 1 -> [w0 + w1, w2]; same for columns.  16 tap-GEMMs on the low-resolution grid replace 9 on the 4x larger one."""
    o, i, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    w = weight.detach().float()
    ip = ceil64(i)
    rows = {0: [w[:, :, 0:1].sum(2), w[:, :, 1:3].sum(2)], 1: [w[:, :, 0:2].sum(2), w[:, :, 2:3].sum(2)]}    # [O, I, 3] each
    out = []
    for a in (0, 1):
        for b in (0, 1):
            pk = torch.zeros(o, 4, ip, dtype=torch.float32, device=w.device)
            t = 0
            for r in rows[a]:
                cols = [r[:, :, 0:1].sum(2), r[:, :, 1:3].sum(2)] if b == 0 else [r[:, :, 0:2].sum(2), r[:, :, 2:3].sum(2)]
                for c in cols:
                    pk[:, t, :i] = c
                    t += 1
            out.append(pk.reshape(o, 4 * ip).to(dtype).contiguous())
    return out


def pack_conv3x3(weight: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """OIHW 3x3 weight -> ``[O, 9 * ceil64(I)]`` with k = (kh*3+kw) * ceil64(I) + i."""
    o, i, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    ip = ceil64(i)
    out = torch.zeros(o, 9, ip, dtype=dtype, device=weight.device)
    out[:, :, :i] = weight.detach().float().permute(0, 2, 3, 1).reshape(o, 9, i).to(dtype)
    return out.reshape(o, 9 * ip).contiguous()


def pack_geglu(weight: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> Tuple[torch.Tensor, torch.Tensor]:
    """GEGLU ``proj`` ``[2*I, K]`` (rows [0,I) values, [I,2I) gates) -> rows interleaved in groups of
    16 (16 value rows, then their 16 gate rows) so one 32-column epilogue chunk holds both."""
    two_i, k = weight.shape
    inner = two_i // 2
    assert inner % 16 == 0
    w = weight.detach().float()
    b = bias.detach().float()
    wv, wg = w[:inner].reshape(inner // 16, 16, k), w[inner:].reshape(inner // 16, 16, k)
    bv, bg = b[:inner].reshape(inner // 16, 16), b[inner:].reshape(inner // 16, 16)
    wi = torch.stack([wv, wg], dim=1).reshape(two_i, k)
    bi = torch.stack([bv, bg], dim=1).reshape(two_i)
    return pack_linear(wi, dtype), bi.contiguous()


# ----------------------------------------------------------------------------------------------
# GEMM / implicit-GEMM convolution
# ----------------------------------------------------------------------------------------------
def gemm(a: Sequence[torch.Tensor], w: torch.Tensor, n: int, *, out: torch.Tensor,
         taps: Optional[Sequence[Sequence[Tuple[int, int, int]]]] = None,
         a_c: Optional[Sequence[int]] = None, conv_geom: Optional[Tuple[int, int, int]] = None,
         bias: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None,
         rows_per_batch: int = 1, residual: Optional[torch.Tensor] = None, geglu: bool = False,
         block_n: int = 0, act: int = 0, gn_partial: Optional[torch.Tensor] = None, cta_group: int = 0,
         up_phase: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """``out = epilogue(A @ w.T)`` on tcgen05 tensor cores (see ``dfb_gemm`` in include/dfb200.h).

    a:    1 or 2 bf16 operands; each ``[M, C]`` (plain) or ``[B, H, W, C]`` (conv), last dim
          contiguous; the row pitch is taken from ``stride(-2)``.
    taps: per segment a list of ``(dh, dw, channel_offset)``; default one centre tap.
    a_c:  per segment channel extent read per tap (default: the tensor's last dim).
    up_phase: ``(a, b)``: this conv is phase (a, b) of a fused nearest-2x upsample + 3x3 conv (``pack_upsample_phases``):
          ``a`` is the LOW-resolution ``[B, H, W, C]`` input, ``out`` the whole ``[B, 2H, 2W, N]`` output of which this call
          writes the pixels ``(2i + a, 2j + b)``; ``gn_partial`` is the partial-statistics buffer of the whole output.
    """
    f32 = a[0].dtype == torch.float32          # fp32 verification path (CUDA cores): fp32 operands throughout
    if f32 and geglu:
        # projection with bias into a scratch buffer, then pair the packed (value | gate) columns
        tmp = torch.empty(out.numel() // out.shape[-1], n, dtype=torch.float32, device=out.device)
        gemm(a, w, n, out=tmp, taps=taps, a_c=a_c, conv_geom=conv_geom, bias=bias)
        check(_lib.load().dfb_geglu_f32(tmp.data_ptr(), tmp.stride(0), out.data_ptr(), out.stride(-2), tmp.shape[0], n, _stream()),
              "dfb_geglu_f32")
        _count(1)
        return out
    p = GemmParams()
    nseg = len(a)
    assert 1 <= nseg <= 2
    p.nseg = nseg
    m_rows = 1
    for d in out.shape[:-1]:
        m_rows *= d
    if up_phase is not None:
        assert conv_geom is not None and m_rows == 4 * conv_geom[0] * conv_geom[1] * conv_geom[2] and out.is_contiguous()
        m_rows //= 4
        p.up2x = 1 + 2 * int(up_phase[0]) + int(up_phase[1])
    op_dtype = torch.float32 if f32 else torch.bfloat16
    for s, t in enumerate(a):
        assert t.dtype == op_dtype and t.is_cuda and t.stride(-1) == 1
        p.a[s] = t.data_ptr()
        p.a_ld[s] = t.stride(-2)
        p.a_c[s] = a_c[s] if a_c is not None else t.shape[-1]
        tp = taps[s] if taps is not None else TAP_CENTER
        p.ntaps[s] = len(tp)
        for i, (dh, dw, co) in enumerate(tp):
            p.tap_dh[s][i], p.tap_dw[s][i], p.tap_coff[s][i] = dh, dw, co
    if conv_geom is not None:
        p.conv = 1
        p.B, p.H, p.W = conv_geom
    p.M, p.N = m_rows, n
    assert w.dtype == op_dtype and w.is_contiguous()
    p.w, p.w_ld = w.data_ptr(), w.shape[1]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        p.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(-1) == 1
        p.rowbias, p.rowbias_ld, p.rows_per_batch = rowbias.data_ptr(), rowbias.stride(0), rows_per_batch
    if residual is not None:
        assert residual.stride(-1) == 1
        p.residual, p.res_ld, p.res_dtype = residual.data_ptr(), residual.stride(-2), _dt(residual)
    assert out.stride(-1) == 1
    p.out, p.out_ld, p.out_dtype = out.data_ptr(), out.stride(-2), _dt(out)
    p.geglu = 1 if geglu else 0
    p.act = act
    p.block_n = block_n
    p.cta_group = cta_group
    if gn_partial is not None:
        assert gn_partial.dtype == torch.float32 and gn_partial.numel() >= ((4 if up_phase is not None else 1) * m_rows // 32) * (n // 2) * 2
        p.gn_partial = gn_partial.data_ptr()
    e0 = _prof_begin()
    if f32:
        check(_lib.load().dfb_gemm_f32(C.byref(p), _stream()), "dfb_gemm_f32")
    else:
        check(_lib.load().dfb_gemm(C.byref(p), _stream()), "dfb_gemm")
    if e0 is not None:
        k_exec = sum(int(p.ntaps[s]) * ceil64(int(p.a_c[s])) for s in range(nseg))
        # algorithmic bytes: every operand / result tensor once (a conv reads its input once, not once per tap)
        esz = 4 if f32 else 2
        nb = sum(m_rows * max(int(p.a_c[s]) + max(co for _, _, co in (taps[s] if taps is not None else TAP_CENTER)), 1) * esz
                 for s in range(nseg))
        nb += n * k_exec * esz + (n * 4 if bias is not None else 0)
        nb += (rowbias.shape[0] * n * 4 if rowbias is not None else 0) + (m_rows * n * residual.element_size() if residual is not None else 0)
        out_rows = m_rows * (4 if up_phase is not None else 1)
        nb += m_rows * (n // 2 if geglu else n) * out.element_size() + (gn_partial.numel() * 4 // (4 if up_phase is not None else 1) if gn_partial is not None else 0)
        _prof_end(e0, "conv" if conv_geom is not None else "gemm", 2.0 * m_rows * n * k_exec, (m_rows, n, k_exec), float(nb))
    _count(1)
    return out


ACT_NONE, ACT_SILU, ACT_LEAKY_RELU, ACT_TANH, ACT_QUICK_GELU, ACT_GELU = 0, 1, 2, 3, 4, 5


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
def pad16(d: int) -> int:
    return (d + 15) // 16 * 16


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, heads: int, dp: int,
              scale: float, q_col0: int = 0, k_col0: int = 0, v_col0: int = 0, out_col0: int = 0,
              block_kv: int = 0, dbg_v_lbo: int = 0, dbg_v_sbo: int = 0, dbg_flags: int = 0, dbg_timeline: Optional[torch.Tensor] = None,
              causal: bool = False, ones_col: Optional[int] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(Q K^T * scale) V per (batch, head); q/k/v/out are bf16 ``[B, S, ld]`` views whose head
    ``h`` lives in columns ``[col0 + h*dp, col0 + (h+1)*dp)`` (``dp`` = head dim padded to 16).  ``causal``: key j is
    visible to query i only when j <= i (CLIP text encoder).  ``ones_col`` = c: the caller promises that column c of every
    padded head of ``v`` holds 1.0 (``AttnPack.b_qkv`` writes it through the projection's bias), so the long-sequence kernel
    takes the softmax denominator from the P V product (see ``dfb_attn_params.ones_col``).  ``workspace``: caller-owned
    int32 scratch of ``attention_ws_elems(B, heads, Sq)`` elements (``dfb_attn_params.workspace``: the 8-softmax-warp kernel)."""
    from ._lib import AttnParams
    p = AttnParams()
    for t in (q, k, v, out):
        assert t.dtype == q.dtype and t.dtype in (torch.bfloat16, torch.float32) and t.is_cuda and t.dim() == 3 and t.stride(-1) == 1
        assert t.stride(0) == t.shape[1] * t.stride(1), "batch stride must be S * row pitch"
    p.q, p.k, p.v, p.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    p.q_ld, p.k_ld, p.v_ld, p.out_ld = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    p.q_col0, p.k_col0, p.v_col0, p.out_col0 = q_col0, k_col0, v_col0, out_col0
    p.B, p.heads, p.Sq, p.Skv, p.dp = q.shape[0], heads, q.shape[1], k.shape[1], dp
    p.scale = scale
    p.block_kv = block_kv
    p.dbg_v_lbo, p.dbg_v_sbo = dbg_v_lbo, dbg_v_sbo
    p.dbg_flags = dbg_flags
    p.dbg_timeline = _ptr(dbg_timeline)
    p.causal = 1 if causal else 0
    p.ones_col = 0 if ones_col is None else int(ones_col) + 1
    if workspace is not None:
        assert workspace.dtype == torch.int32 and workspace.is_cuda and workspace.is_contiguous()
        assert workspace.numel() >= attention_ws_elems(q.shape[0], heads, q.shape[1])
        p.workspace = workspace.data_ptr()
    e0 = _prof_begin()
    if q.dtype == torch.float32:
        check(_lib.load().dfb_attention_f32(C.byref(p), _stream()), "dfb_attention_f32")
    else:
        check(_lib.load().dfb_attention(C.byref(p), _stream()), "dfb_attention")
    if e0 is not None:
        _prof_end(e0, "attention", 4.0 * q.shape[0] * heads * q.shape[1] * k.shape[1] * dp, (q.shape[0], heads, q.shape[1], k.shape[1], dp))
    # the 8-softmax-warp self-attention path is two launches (kernel + the redo pass over flagged tiles); same condition as the host side
    two = (workspace is not None and ones_col is not None and q.dtype == torch.bfloat16 and k.shape[1] > 128 and 192 + dp <= 256
           and not causal and (dbg_flags & (8 | 16 | 4096 | 32768)) == 0 and block_kv in (0, 64))
    _count(2 if two else 1)
    return out


def attention_ws_elems(b: int, heads: int, sq: int) -> int:
    """int32 elements of the scratch ``attention(workspace=...)`` takes (one flag per 128-query tile of every (batch, head))."""
    return b * heads * ((sq + 127) // 128)


# ----------------------------------------------------------------------------------------------
# norms
# ----------------------------------------------------------------------------------------------
def _profiled(kind):
    """bench.py's per-launch CUDA-event profile also covers the streaming / norm kernels (bytes instead of flops)."""
    def deco(fn):
        def wrapper(*a, **k):
            e0 = _prof_begin()
            out = fn(*a, **k)
            if e0 is not None:
                nbytes = 0.0
                for t in list(a) + list(k.values()):
                    if torch.is_tensor(t):
                        nbytes += t.numel() * t.element_size()
                _prof_end(e0, kind, nbytes, tuple(out.shape) if torch.is_tensor(out) else ())
            return out
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def groupnorm_ws_floats(b: int, groups: int) -> int:
    return int(_lib.load().dfb_groupnorm_ws_floats(b, groups))


def gn_partial_shape(m_rows: int, n: int):
    """Shape of the GroupNorm partial-statistics buffer a GEMM epilogue emits for an fp32 ``[m_rows, n]`` output."""
    return (m_rows // 32, n // 2, 2)


@_profiled("groupnorm")
def groupnorm(src0: torch.Tensor, src1: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, *,
              groups: int, eps: float, silu: bool, stats_ws: torch.Tensor, out: torch.Tensor,
              raw_out: Optional[torch.Tensor] = None, partials: Optional[Sequence[Optional[torch.Tensor]]] = None) -> torch.Tensor:
    """GroupNorm(+SiLU) over the channel-concat of fp32 NHWC sources ``[B, H, W, C]`` -> bf16 ``out``.
    ``partials`` = per-source statistics emitted by the producers' GEMM epilogues (skips the statistics pass)."""
    if partials is not None and partials[0] is not None and (src1 is None or partials[1] is not None):
        b = src0.shape[0]
        hw = src0.shape[1] * src0.shape[2] if src0.dim() == 4 else src0.shape[1]
        c0 = src0.shape[-1]
        c1 = 0 if src1 is None else src1.shape[-1]
        assert hw % 32 == 0 and src0.dtype == torch.float32 and (raw_out is None or raw_out.dtype == out.dtype)
        check(_lib.load().dfb_groupnorm_fused(
            src0.data_ptr(), c0, src0.stride(-2), partials[0].data_ptr(), _ptr(src1), c1,
            0 if src1 is None else src1.stride(-2), None if src1 is None else partials[1].data_ptr(), b, hw, groups, eps,
            gamma.data_ptr(), beta.data_ptr(), 1 if silu else 0, stats_ws.data_ptr(), out.data_ptr(), _dt(out), out.stride(-2),
            _ptr(raw_out), 0 if raw_out is None else raw_out.stride(-2), _stream()), "dfb_groupnorm_fused")
        _count(2)
        return out
    b = src0.shape[0]
    hw = src0.shape[1] * src0.shape[2] if src0.dim() == 4 else src0.shape[1]
    c0 = src0.shape[-1]
    c1 = 0 if src1 is None else src1.shape[-1]
    assert src0.dtype == torch.float32 and (raw_out is None or raw_out.dtype == out.dtype)
    assert stats_ws.dtype == torch.float32 and stats_ws.numel() >= groupnorm_ws_floats(b, groups)
    check(_lib.load().dfb_groupnorm(
        src0.data_ptr(), c0, src0.stride(-2), _ptr(src1), c1, 0 if src1 is None else src1.stride(-2), b, hw,
        groups, eps, gamma.data_ptr(), beta.data_ptr(), 1 if silu else 0, stats_ws.data_ptr(), out.data_ptr(), _dt(out),
        out.stride(-2), _ptr(raw_out), 0 if raw_out is None else raw_out.stride(-2), _stream()), "dfb_groupnorm")
    _count(3)
    return out


@_profiled("layernorm")
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, eps: float = 1e-5):
    """LayerNorm over the last dim of fp32 ``[rows, C]`` -> ``out`` (bf16, or fp32 on the verification path)."""
    rows = x.numel() // x.shape[-1]
    assert x.dtype == torch.float32
    check(_lib.load().dfb_layernorm(x.data_ptr(), x.stride(-2), gamma.data_ptr(), beta.data_ptr(), eps,
                                    out.data_ptr(), _dt(out), out.stride(-2), rows, x.shape[-1], _stream()), "dfb_layernorm")
    _count(1)
    return out


@_profiled("softmax_rows")
def softmax_rows(x: torch.Tensor, out: torch.Tensor, scale: float = 1.0):
    """``out[r, :] = softmax(scale * x[r, :])`` for a materialised fp32 score matrix ``[rows, cols]``."""
    rows, cols = x.shape
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and out.stride(-1) == 1 and out.shape == x.shape
    check(_lib.load().dfb_softmax_rows(x.data_ptr(), x.stride(0), float(scale), out.data_ptr(), _dt(out), out.stride(0),
                                       rows, cols, _stream()), "dfb_softmax_rows")
    _count(1)
    return out


# ----------------------------------------------------------------------------------------------
# streaming kernels
# ----------------------------------------------------------------------------------------------
@_profiled("cfg_step")
def cfg_step(eps: torch.Tensor, weights: Sequence[float], x_src: torch.Tensor, cx: float, ck: Sequence[float],
             hist: Sequence[Optional[torch.Tensor]] = (None, None, None), noise: Optional[torch.Tensor] = None,
             cn: float = 0.0, x_out: Optional[torch.Tensor] = None, eps_out: Optional[torch.Tensor] = None,
             eps_nchw: bool = False):
    """Fused CFG combine + scheduler update.  eps: fp32 NHWC ``[nb*N, H, W, 4]`` (or NCHW ``[nb*N, 4, H, W]``
    with ``eps_nchw``); x: fp32 NCHW."""
    nb = len(weights)
    n_items, hw = x_src.shape[0], x_src.shape[2] * x_src.shape[3]
    assert eps.dtype == torch.float32 and eps.is_contiguous() and eps.numel() == nb * n_items * hw * 4
    assert x_src.dtype == torch.float32 and x_src.is_contiguous() and x_src.shape[1] == 4
    if x_out is None:
        x_out = torch.empty_like(x_src)
    w = (C.c_float * nb)(*[float(v) for v in weights])
    ckk = (C.c_float * 4)(*[float(v) for v in (list(ck) + [0.0] * 4)[:4]])
    h = list(hist) + [None] * 3
    check(_lib.load().dfb_cfg_step(eps.data_ptr(), 1 if eps_nchw else 0, nb, w, x_src.data_ptr(), float(cx), ckk, _ptr(h[0]), _ptr(h[1]),
                                   _ptr(h[2]), _ptr(noise), float(cn), x_out.data_ptr(), _ptr(eps_out), n_items, hw,
                                   _stream()), "dfb_cfg_step")
    _count(1)
    return x_out


@_profiled("mutual_gather_sum")
def mutual_gather_sum(all_latents: Optional[torch.Tensor], prev_latents: torch.Tensor, idx: torch.Tensor,
                      out: torch.Tensor):
    n_items, n_src = idx.shape
    d = prev_latents[0].numel()
    assert idx.dtype == torch.int32 and idx.is_contiguous()
    check(_lib.load().dfb_mutual_gather_sum(_ptr(all_latents), prev_latents.data_ptr(), idx.data_ptr(), n_items,
                                            n_src, d, out.data_ptr(), _dt(out), _stream()), "dfb_mutual_gather_sum")
    _count(1)
    return out


@_profiled("mutual_blend")
def mutual_blend(x: torch.Tensor, m: Optional[torch.Tensor], hist: Optional[torch.Tensor], null_latent: torch.Tensor,
                 eta: float, use_m: Sequence[int], use_h: Sequence[int], out: torch.Tensor):
    nb = len(use_m)
    n_items, hw = x.shape[0], x.shape[2] * x.shape[3]
    um = (C.c_int32 * nb)(*[int(v) for v in use_m])
    uh = (C.c_int32 * nb)(*[int(v) for v in use_h])
    assert out.numel() == nb * n_items * hw * 8
    check(_lib.load().dfb_mutual_blend(x.data_ptr(), _ptr(m), _ptr(hist), null_latent.data_ptr(), float(eta), nb, um,
                                       uh, n_items, hw, out.data_ptr(), _dt(out), _stream()), "dfb_mutual_blend")
    _count(1)
    return out


@_profiled("nchw_to_nhwc_bf16")
def nchw_to_nhwc_bf16(x: torch.Tensor, out: torch.Tensor):
    b, c, h, w = x.shape
    assert x.is_contiguous()
    check(_lib.load().dfb_nchw_to_nhwc(x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), b, c, h * w, _stream()),
          "dfb_nchw_to_nhwc")
    _count(1)
    return out


@_profiled("nhwc_to_nchw")
def nhwc_to_nchw(x: torch.Tensor, out: torch.Tensor):
    b, c, h, w = out.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    check(_lib.load().dfb_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), _dt(out), b, c, h * w, _stream()),
          "dfb_nhwc_to_nchw")
    _count(1)
    return out


nchw_to_nhwc = nchw_to_nhwc_bf16      # the output dtype follows ``out`` (bf16 operand path / fp32 verification path)


@_profiled("pad_cast_rows")
def pad_cast_rows(x: torch.Tensor, out: torch.Tensor):
    b, s, d = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    check(_lib.load().dfb_pad_cast_rows(x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), b, s, out.shape[1], d, _stream()),
          "dfb_pad_cast_rows")
    _count(1)
    return out


@_profiled("cast")
def cast_f32(x: torch.Tensor, out: torch.Tensor):
    """fp32 -> ``out.dtype`` (the MMA operand type), same shape, both contiguous."""
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and x.numel() == out.numel()
    check(_lib.load().dfb_cast_f32(x.data_ptr(), out.data_ptr(), _dt(out), x.numel(), _stream()), "dfb_cast_f32")
    _count(1)
    return out


@_profiled("upsample2x")
def upsample2x(x: torch.Tensor, out: torch.Tensor):
    b, h, w, c = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    check(_lib.load().dfb_upsample2x(x.data_ptr(), out.data_ptr(), _dt(out), b, h, w, c, _stream()), "dfb_upsample2x")
    _count(1)
    return out


@_profiled("space_to_depth")
def space_to_depth(x: torch.Tensor, out: torch.Tensor):
    b, h, w, c = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    check(_lib.load().dfb_space_to_depth(x.data_ptr(), out.data_ptr(), _dt(out), b, h, w, c, _stream()), "dfb_space_to_depth")
    _count(1)
    return out


def s2d_taps(c: int) -> Tuple[Tuple[int, int, int], ...]:
    """Tap table of a stride-2, pad-1 3x3 conv expressed over the space-to-depth planes (kh-major,
    matching ``pack_conv3x3``): input row 2*ho + kh - 1 -> (plane parity, shift)."""
    par = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
    taps = []
    for kh in range(3):
        ph, dh = par[kh]
        for kw in range(3):
            pw, dw = par[kw]
            taps.append((dh, dw, (ph * 2 + pw) * c))
    return tuple(taps)


@_profiled("timestep_embedding")
def timestep_embedding(t: torch.Tensor, out: torch.Tensor, flip_sin_to_cos: bool = True, freq_shift: float = 0.0):
    assert t.dtype == torch.float32 and t.is_contiguous() and out.is_contiguous()
    check(_lib.load().dfb_timestep_embedding(t.data_ptr(), out.data_ptr(), _dt(out), t.shape[0], out.shape[1],
                                             1 if flip_sin_to_cos else 0, float(freq_shift), _stream()),
          "dfb_timestep_embedding")
    _count(1)
    return out


@_profiled("embed_tokens")
def embed_tokens(ids: torch.Tensor, token_table: torch.Tensor, position_table: torch.Tensor, out: torch.Tensor):
    """CLIPTextEmbeddings: ``out[b*S+s] = token_table[ids[b, s]] + position_table[s]`` (fp32 ``[B*S, D]``)."""
    b, s = ids.shape
    d = token_table.shape[1]
    assert ids.dtype == torch.int32 and ids.is_contiguous() and ids.is_cuda
    assert token_table.dtype == torch.float32 and token_table.is_contiguous() and position_table.dtype == torch.float32
    assert position_table.is_contiguous() and position_table.shape[0] >= s and position_table.shape[1] == d
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == b * s * d
    check(_lib.load().dfb_embed_tokens(ids.data_ptr(), token_table.data_ptr(), position_table.data_ptr(), out.data_ptr(), b, s, d,
                                       token_table.shape[0], _stream()), "dfb_embed_tokens")
    _count(1)
    return out


@_profiled("image_to_uint8")
def image_to_uint8(img: torch.Tensor, out: torch.Tensor):
    """VaeImageProcessor.postprocess: fp32 NHWC ``[B, H, W, C>=3]`` in [-1, 1] -> uint8 ``[B, H, W, 3]``."""
    assert img.dtype == torch.float32 and img.is_contiguous() and img.is_cuda and img.shape[-1] >= 3
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == img.numel() // img.shape[-1] * 3
    check(_lib.load().dfb_image_to_uint8(img.data_ptr(), img.shape[-1], out.data_ptr(), img.numel() // img.shape[-1], _stream()),
          "dfb_image_to_uint8")
    _count(1)
    return out


def s2d_taps_pad0(c: int) -> Tuple[Tuple[int, int, int], ...]:
    """Tap table of the VAE encoder's Downsample2D: ``F.pad(x, (0, 1, 0, 1))`` then a stride-2, pad-0 3x3 conv, over the
    space-to-depth planes: input row 2*ho + kh -> (plane parity, shift); the zero row / column the pad appends is the
    TMA out-of-bounds fill."""
    par = {0: (0, 0), 1: (1, 0), 2: (0, 1)}
    taps = []
    for kh in range(3):
        ph, dh = par[kh]
        for kw in range(3):
            pw, dw = par[kw]
            taps.append((dh, dw, (ph * 2 + pw) * c))
    return tuple(taps)
