"""difashion_b200 — B200-native (sm_100a) implementation of DiFashion's conditional denoising step.

Public surface (mirrors what ``DiFashion/models/difashion.py`` consumes from diffusers):
``B200UNet2DConditionModel``, ``B200AttnProcessor``, ``B200DDIMScheduler``, ``B200PNDMScheduler``,
``MutualEncoder``, ``B200DiFashionPipeline`` (the loop), and the stages around it: ``B200AutoencoderKL`` (encode / decode),
``B200CLIPTextModel``, ``B200DiFashion.fashion_generation`` (the whole reference method), ``save_batch_outputs``.  All arithmetic runs in ``libdfb200.so`` (hand-written CUDA,
C ABI in ``include/dfb200.h``); importing this package never imports the test oracle.
"""
from .attention import Attention, B200AttnProcessor  # noqa: F401
from .mutual import MutualEncoder  # noqa: F401
from .pipeline import (B200DiFashionPipeline, gather_item_rows, guidance_plan, mutual_index_table,  # noqa: F401
                       shard_generation_inputs, shard_outfits)
from .schedulers import B200DDIMScheduler, B200PNDMScheduler  # noqa: F401
from .unet import B200UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
from .vae import B200AutoencoderKL, DecoderOutput  # noqa: F401
from .clip import B200CLIPTextModel  # noqa: F401
from .generation import B200DiFashion  # noqa: F401
from .outputs import merge_and_save_images, save_batch_outputs, save_outputs_npy  # noqa: F401
from . import checkpoint  # noqa: F401  (the reference's on-disk model / checkpoint layout)
from .history import build_history_latents, encode_all_item_latents  # noqa: F401
from .prompts import category_prompt, tokenize_categories  # noqa: F401

__all__ = ["B200UNet2DConditionModel", "UNet2DConditionOutput", "B200AttnProcessor", "Attention", "B200DDIMScheduler",
           "B200PNDMScheduler", "MutualEncoder", "B200DiFashionPipeline", "guidance_plan", "mutual_index_table",
           "shard_outfits", "shard_generation_inputs", "gather_item_rows", "B200AutoencoderKL", "DecoderOutput", "B200CLIPTextModel", "B200DiFashion",
           "save_batch_outputs", "merge_and_save_images", "save_outputs_npy", "build_history_latents", "encode_all_item_latents", "category_prompt", "tokenize_categories"]
