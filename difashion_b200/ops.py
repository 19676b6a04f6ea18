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
def pack_linear(weight: torch.Tensor) -> torch.Tensor:
    """nn.Linear / 1x1-conv weight ``[N, K(,1,1)]`` -> bf16 ``[N, ceil64(K)]`` (K-major)."""
    w = weight.detach().reshape(weight.shape[0], -1).float()
    n, k = w.shape
    out = torch.zeros(n, ceil64(k), dtype=torch.bfloat16, device=w.device)
    out[:, :k] = w.to(torch.bfloat16)
    return out.contiguous()


def pack_conv3x3(weight: torch.Tensor) -> torch.Tensor:
    """OIHW 3x3 weight -> bf16 ``[O, 9 * ceil64(I)]`` with k = (kh*3+kw) * ceil64(I) + i."""
    o, i, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    ip = ceil64(i)
    out = torch.zeros(o, 9, ip, dtype=torch.bfloat16, device=weight.device)
    out[:, :, :i] = weight.detach().float().permute(0, 2, 3, 1).reshape(o, 9, i).to(torch.bfloat16)
    return out.reshape(o, 9 * ip).contiguous()


def pack_geglu(weight: torch.Tensor, bias: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
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
    return pack_linear(wi), bi.contiguous()


# ----------------------------------------------------------------------------------------------
# GEMM / implicit-GEMM convolution
# ----------------------------------------------------------------------------------------------
def gemm(a: Sequence[torch.Tensor], w: torch.Tensor, n: int, *, out: torch.Tensor,
         taps: Optional[Sequence[Sequence[Tuple[int, int, int]]]] = None,
         a_c: Optional[Sequence[int]] = None, conv_geom: Optional[Tuple[int, int, int]] = None,
         bias: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None,
         rows_per_batch: int = 1, residual: Optional[torch.Tensor] = None, geglu: bool = False,
         block_n: int = 0) -> torch.Tensor:
    """``out = epilogue(A @ w.T)`` on tcgen05 tensor cores (see ``dfb_gemm`` in include/dfb200.h).

    a:    1 or 2 bf16 operands; each ``[M, C]`` (plain) or ``[B, H, W, C]`` (conv), last dim
          contiguous; the row pitch is taken from ``stride(-2)``.
    taps: per segment a list of ``(dh, dw, channel_offset)``; default one centre tap.
    a_c:  per segment channel extent read per tap (default: the tensor's last dim).
    """
    p = GemmParams()
    nseg = len(a)
    assert 1 <= nseg <= 2
    p.nseg = nseg
    m_rows = 1
    for d in out.shape[:-1]:
        m_rows *= d
    for s, t in enumerate(a):
        assert t.dtype == torch.bfloat16 and t.is_cuda and t.stride(-1) == 1
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
    assert w.dtype == torch.bfloat16 and w.is_contiguous()
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
    p.block_n = block_n
    check(_lib.load().dfb_gemm(C.byref(p), _stream()), "dfb_gemm")
    return out
