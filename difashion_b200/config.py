"""diffusers-style ``config`` object for the B200 UNet (attribute *and* mapping access)."""
from __future__ import annotations

from typing import Any, Dict

SD15_UNET_CONFIG: Dict[str, Any] = dict(
    sample_size=64, in_channels=8, out_channels=4, center_input_sample=False, flip_sin_to_cos=True,
    freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    mid_block_type="UNetMidBlock2DCrossAttn",
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
    cross_attention_dim=768, attention_head_dim=8, use_linear_projection=False, upcast_attention=False,
    resnet_time_scale_shift="default", dual_cross_attention=False, class_embed_type=None,
)

# DiFashion's default base model (train.py:44, inf4eval.py:65) — same topology, different heads.
SD2_BASE_UNET_OVERRIDES: Dict[str, Any] = dict(
    cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20), use_linear_projection=True)


class FrozenConfig(dict):
    """Mimics diffusers' FrozenDict config: ``cfg.sample_size`` and ``cfg["sample_size"]``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen; use register_to_config(**kwargs)")
