"""The denoising loop of ``DiFashion.fashion_generation`` on the B200 kernels.

Mirrors ``DiFashion/models/difashion.py:277-616`` for the part that is the hot path (the loop body
``:456-577`` and its static set-up ``:309-325``, ``:388-451``); VAE / CLIP stages are the caller's.
"""
from __future__ import annotations

import torch

INT32_MIN = -(2 ** 31)


def mutual_index_table(olists: torch.Tensor) -> torch.Tensor:
    """Static gather table for the mutual condition (difashion.py:439-451 + :477-487).

    Row n (n-th generated item in ``torch.nonzero(olists == 0)`` order) lists the sources of the other
    ``olen-1`` slots of its outfit: ``j >= 0`` -> row j of ``all_latents`` (a given item),
    ``j < 0`` -> row ``-j-1`` of the generated latents.  int32 ``[N, olen-1]`` (CPU)."""
    olists = olists.cpu()
    bsz, olen = olists.shape
    gen = olists == 0
    gen_row = torch.full((bsz, olen), -1, dtype=torch.int64)
    gen_row[gen] = torch.arange(int(gen.sum()))
    rows = []
    for o in range(bsz):
        for i in range(olen):
            if not gen[o, i]:
                continue
            row = []
            for s in range(olen):
                if s == i:
                    continue
                row.append(-int(gen_row[o, s]) - 1 if gen[o, s] else o * olen + s)
            rows.append(row)
    return torch.tensor(rows, dtype=torch.int32).reshape(len(rows), olen - 1)
