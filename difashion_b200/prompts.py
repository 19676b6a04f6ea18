"""Category prompts of DiFashion (``DiFashion/data_utils.py:88-111``): the text side of the caller's data format.

Every item slot is prompted with ``"A photo of a <category>, on white background, high quality"`` — ``"a pair of"`` for
categories containing ``pants`` or ``earrings`` — tokenised to ``model_max_length`` (77) ids; ``fashion_generation``
receives them as ``input_ids [outfits, olen, 77]``.  Only ~50 distinct prompts exist, which is why
``B200DiFashion.encode_prompts`` caches the encoder output per distinct id row.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch

SPECIAL_CATES = ("pants", "earrings")


def category_prompt(category: str) -> str:
    if any(s in category for s in SPECIAL_CATES):
        return "A photo of a pair of " + category + ", on white background, high quality"
    return "A photo of a " + category + ", on white background, high quality"


def tokenize_categories(tokenizer, outfit_categories: Sequence[Sequence[int]], id_cate_dict: Mapping[int, str]) -> torch.Tensor:
    """``data["input_ids"]`` of ``data_utils.tokenize_category``: int64 ``[outfits, olen, model_max_length]``."""
    rows = []
    for outfit in outfit_categories:
        prompts = [category_prompt(id_cate_dict[int(cid)]) for cid in outfit]
        ids = tokenizer(prompts, max_length=tokenizer.model_max_length, padding="max_length", truncation=True,
                        return_tensors="pt").input_ids
        rows.append(torch.as_tensor(ids, dtype=torch.long))
    return torch.stack(rows)
