"""Shared helpers for the parity tests."""
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def err_report(got: torch.Tensor, ref: torch.Tensor, name: str = "") -> str:
    """Where is the error? (per 32-row / 32-col block maxima) — makes one GPU run informative."""
    g, r = got.double(), ref.double()
    d = (g - r).abs()
    lines = [f"[{name}] shape={tuple(g.shape)} rel_l2={rel_l2(g, r):.3e} max_abs={d.max():.3e} "
             f"ref_absmax={r.abs().max():.3e} got_absmax={g.abs().max():.3e} nan={int(torch.isnan(g).sum())}"]
    if d.dim() == 2:
        m, n = d.shape
        rb = d[: m // 32 * 32].reshape(-1, 32, n).amax(dim=(1, 2)) if m >= 32 else d.amax(dim=1)
        cb = d[:, : n // 16 * 16].reshape(m, -1, 16).amax(dim=(0, 2)) if n >= 16 else d.amax(dim=0)
        lines.append("  row-block(32) max err: " + " ".join(f"{x:.1e}" for x in rb[:24].tolist()))
        lines.append("  col-block(16) max err: " + " ".join(f"{x:.1e}" for x in cb[:24].tolist()))
    return "\n".join(lines)
