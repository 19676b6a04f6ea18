"""``B200DiFashion`` — the whole of ``DiFashion.fashion_generation`` (``DiFashion/models/difashion.py:277-616``) on the
B200 kernels: the stages before the loop (CLIP prompt encoding ``:339-353``, VAE encode of the white image and the given
items ``:375-376``, ``:435-437``, history lookup ``:378-386``), the denoising loop (``B200DiFashionPipeline``, ``:456-577``)
and the stages after it (VAE decode + ``VaeImageProcessor.postprocess`` ``:579-592``, result dictionary ``:598-614``).

Same argument names, same two return forms as the reference method (``return_dict=False`` -> ``(all_results,
init_latents)`` as ``inf4eval.py:736-751`` consumes it), so ``inf4eval.py``'s loop can call it unchanged.
Python exceptions are the error convention, as in diffusers; there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch

from .clip import B200CLIPTextModel
from .mutual import MutualEncoder
from .pipeline import B200DiFashionPipeline
from .unet import B200UNet2DConditionModel
from .vae import B200AutoencoderKL


@dataclass
class StableDiffusionPipelineOutput:
    images: Any
    nsfw_content_detected: Optional[List[bool]] = None


class B200DiFashion:
    """Drop-in for the inference side of ``DiFashion`` (``difashion.py:48-120``): holds ``unet``, ``vae``, ``text_encoder``,
    ``fashion_encoder`` and ``noise_scheduler`` under the reference's attribute names."""

    def __init__(self, unet: B200UNet2DConditionModel, vae: B200AutoencoderKL, text_encoder: B200CLIPTextModel,
                 fashion_encoder: Optional[MutualEncoder], noise_scheduler, *, eta: float = 0.1, use_history: bool = True,
                 use_mutual_guidance: bool = True, max_rows: int = 256, use_cuda_graph: bool = True,
                 reference_history_lookup: bool = True):
        self.unet, self.vae, self.text_encoder = unet, vae, text_encoder
        self.fashion_encoder, self.noise_scheduler = fashion_encoder, noise_scheduler
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)
        self.use_history, self.use_mutual_guidance = use_history, use_mutual_guidance
        # The reference tests `cate in history[uid]` with `cate` a 0-d TENSOR (difashion.py:380-382): tensors hash by
        # identity, so that membership test is always False and every item gets the null latent.  True (default)
        # reproduces the reference — same outputs for the same checkpoint and seed; False implements the evident intent
        # (integer category keys), an explicit opt-in because it feeds the UNet an input the reference never produces.
        self.reference_history_lookup = reference_history_lookup
        self.pipe = B200DiFashionPipeline(unet, fashion_encoder, noise_scheduler, eta_mutual=eta, use_history=use_history,
                                          use_mutual_guidance=use_mutual_guidance, max_rows=max_rows,
                                          use_cuda_graph=use_cuda_graph)
        self._prompt_cache: Dict[bytes, torch.Tensor] = {}
        self.tokenizer = None           # optional CLIPTokenizer (difashion.py:66-68); only its empty-prompt ids are used here

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, checkpoint: Optional[str] = None, *, cate_num: int = 50,
                        category_emb_size: int = 64, hid_dim: int = 256, scheduler: str = "pndm", device=None,
                        load_tokenizer: bool = True, **kw) -> "B200DiFashion":
        """What ``DiFashion.__init__`` does (difashion.py:52-101) from a local Stable-Diffusion directory: scheduler,
        text encoder, VAE and UNet from their sub-folders, ``conv_in`` widened 4 -> 8 with zero-initialised history
        channels (``:82-93``), a fresh xavier-normal ``MutualEncoder`` (``:95-101``) sized from the VAE / UNet configs;
        then, when ``checkpoint`` is given, ``load_model_hook`` (inf4eval.py:556-581).  ``scheduler``: ``"pndm"`` (the
        reference's hard-wired class) or ``"ddim"`` (BASELINE's metric) — both read ``scheduler/scheduler_config.json``."""
        from . import checkpoint as ck
        from .schedulers import B200DDIMScheduler, B200PNDMScheduler
        root = pretrained_model_name_or_path
        sched_cls = {"pndm": B200PNDMScheduler, "ddim": B200DDIMScheduler}[scheduler]
        noise_scheduler = sched_cls.from_pretrained(root, subfolder="scheduler")
        text_encoder = B200CLIPTextModel.from_pretrained(root, subfolder="text_encoder")
        vae = B200AutoencoderKL.from_pretrained(root, subfolder="vae")
        unet = ck.widen_conv_in(B200UNet2DConditionModel.from_pretrained(root, subfolder="unet"), 8)
        fashion_encoder = MutualEncoder(cate_num=cate_num, cate_emb_size=category_emb_size,
                                        latent_channels=vae.config.latent_channels, latent_size=unet.config.sample_size,
                                        hid_dim=hid_dim)
        if device is not None:
            for m in (text_encoder, vae, unet, fashion_encoder):
                m.to(device)
        self = cls(unet, vae, text_encoder, fashion_encoder, noise_scheduler, **kw)
        if load_tokenizer:
            import os
            if os.path.isdir(os.path.join(root, "tokenizer")):
                from transformers import CLIPTokenizer
                self.tokenizer = CLIPTokenizer.from_pretrained(root, subfolder="tokenizer")
        if checkpoint is not None:
            self.load_checkpoint(checkpoint)
        return self

    @classmethod
    def from_args(cls, args, logger=None, cate_num: int = 50, device=None, **kw) -> "B200DiFashion":
        """The reference constructor's signature, ``DiFashion(args, logger, cate_num, device)`` (difashion.py:52-58), reading
        the same ``args`` fields (``pretrained_model_name_or_path``, ``category_emb_size``, ``hid_dim``, ``eta``)."""
        if logger is not None:
            logger.info("load scheduler / CLIPTextModel / VAE / UNet for the B200 path...")
        for flag in ("use_history", "use_mutual_guidance"):          # the reference's ablation switches (inf4eval.py:148-160)
            if hasattr(args, flag):
                kw.setdefault(flag, bool(getattr(args, flag)))
        return cls.from_pretrained(args.pretrained_model_name_or_path, cate_num=cate_num,
                                   category_emb_size=getattr(args, "category_emb_size", 64),
                                   hid_dim=getattr(args, "hid_dim", 256), eta=getattr(args, "eta", 0.1), device=device, **kw)

    def load_checkpoint(self, input_dir: str, use_ema: bool = False, use_ema_fashion: bool = False) -> "B200DiFashion":
        """``load_model_hook`` (inf4eval.py:556-581): ``<input_dir>/unet`` and ``<input_dir>/fashion_encoder`` (diffusers
        ``save_pretrained`` directories) replace the weights AND the configs of the live modules, in place.
        ``use_ema`` / ``use_ema_fashion`` (``--use_ema``, ``--use_ema_fashion``): inference runs on the EMA copies
        (``ema_unet.copy_to(unet.parameters())``, inf4eval.py:691-697) — ``EMAModel.save_pretrained`` writes them as ordinary
        model directories ``unet_ema`` / ``fashion_encoder_ema`` (shadow parameters + a few extra config keys), so they load
        the same way."""
        loaded = B200UNet2DConditionModel.from_pretrained(input_dir, subfolder="unet_ema" if use_ema else "unet")
        if loaded.conv_in.weight.shape[1] != self.unet.conv_in.weight.shape[1]:
            raise RuntimeError(f"checkpoint UNet has {loaded.conv_in.weight.shape[1]} input channels, the model "
                               f"{self.unet.conv_in.weight.shape[1]}")
        self.unet.register_to_config(**loaded.config)        # (EMA bookkeeping keys are not model config: from_pretrained dropped them)
        self.unet.load_state_dict(loaded.state_dict())
        enc = MutualEncoder.from_pretrained(input_dir, subfolder="fashion_encoder_ema" if use_ema_fashion else "fashion_encoder")
        self.fashion_encoder.register_to_config(**enc.config)
        self.fashion_encoder.load_state_dict(enc.state_dict())
        self._prompt_cache.clear()
        self.pipe._states.clear()            # captured graphs point into the packed weights of the previous checkpoint
        return self

    def save_checkpoint(self, output_dir: str, safe_serialization: bool = False) -> None:
        """``save_model_hook`` (inf4eval.py:543-554)."""
        import os
        self.fashion_encoder.save_pretrained(os.path.join(output_dir, "fashion_encoder"), safe_serialization=safe_serialization)
        self.unet.save_pretrained(os.path.join(output_dir, "unet"), safe_serialization=safe_serialization)

    @property
    def device(self):
        return self.unet.device

    # ------------------------------------------------------------------------------------------
    def encode_prompts(self, fill_input_ids: torch.Tensor):
        """``text_encoder(fill_input_ids)[0]`` and the empty prompt (difashion.py:339-352).  Only <= 50 category prompts
        exist (data_utils.py:102-106), so distinct id rows are encoded once and cached across calls."""
        ids = fill_input_ids.to("cpu", torch.int64)
        if self.tokenizer is not None:          # tokenizer([""], padding="max_length", ...) — difashion.py:343-350
            null_ids = torch.as_tensor(self.tokenizer([""], padding="max_length", max_length=ids.shape[1], truncation=True)["input_ids"],
                                       dtype=torch.int64)
        else:
            null_ids = self.text_encoder.null_input_ids(ids.shape[1])
        uniq, inverse = torch.unique(torch.cat([ids, null_ids], 0), dim=0, return_inverse=True)
        missing = [i for i in range(uniq.shape[0]) if uniq[i].numpy().tobytes() not in self._prompt_cache]
        if missing:
            enc = self.text_encoder(uniq[missing].to(self.device))[0]
            for j, i in enumerate(missing):
                self._prompt_cache[uniq[i].numpy().tobytes()] = enc[j]
        table = torch.stack([self._prompt_cache[uniq[i].numpy().tobytes()] for i in range(uniq.shape[0])])
        rows = table[inverse.to(table.device)]
        return rows[:-1], rows[-1:]

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        """difashion.py:618-633."""
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective "
                             f"batch size of {batch_size}. Make sure the batch size matches the length of the generators.")
        if latents is None:
            gdev = generator.device if isinstance(generator, torch.Generator) else device
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents        # scaled by scheduler.init_noise_sigma (= 1.0 for DDIM / PNDM) in B200DiFashionPipeline.begin

    @torch.no_grad()
    def fashion_generation(self, uids: torch.Tensor = None, oids: torch.Tensor = None, input_ids: torch.Tensor = None,
                           olists: torch.Tensor = None, outfit_images=None, category: torch.Tensor = None, history: dict = None,
                           height: Optional[int] = None, width: Optional[int] = None, num_inference_steps: int = 50,
                           category_guidance_scale: float = 7.5, hist_guidance_scale: float = 7.5,
                           mutual_guidance_scale: float = 7.5, null_img: torch.Tensor = None, eta: float = 0.0,
                           init_latents: torch.Tensor = None, generator=None, output_type: Optional[str] = "pil",
                           return_dict: bool = True, callback=None, callback_steps: int = 1):
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("B200DiFashion needs its models on a CUDA device: there is no CPU fallback")
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        olists_c, category_c = olists.cpu(), category.cpu()
        uids_c, oids_c = uids.cpu(), oids.cpu()
        fill_idx = torch.nonzero(olists_c == 0)                                    # :332-337
        fill_num = fill_idx.shape[0]
        fill_cate = category_c[fill_idx[:, 0], fill_idx[:, 1]]
        fill_uids, fill_oids = uids_c[fill_idx[:, 0]], oids_c[fill_idx[:, 0]]
        full_cate = category_c[fill_idx[:, 0]]
        fill_input_ids = input_ids.cpu()[fill_idx[:, 0], fill_idx[:, 1]]
        category_prompts, null_prompt = self.encode_prompts(fill_input_ids)        # :339-352

        if init_latents is None:                                                   # :359-372
            latents = self.prepare_latents(fill_num, self.vae.config.latent_channels, height, width, torch.float32, dev,
                                           generator)
            init_latents = latents.clone()
        else:
            latents = init_latents.to(dev).clone()

        null_latent = self.vae.encode_latents(null_img.unsqueeze(0).to(dev))[0]     # :375-376
        hist = []                                                                  # :378-386
        for i, cate in enumerate(fill_cate.tolist()):
            uid = int(uids_c[fill_idx[i][0]])
            user_hist = (history or {}).get(uid, {})
            if self.use_history and not self.reference_history_lookup and cate in user_hist:
                hist.append(torch.as_tensor(user_hist[cate]).to(device=dev, dtype=torch.float32))
            else:
                hist.append(null_latent)
        hist_latents = torch.stack(hist)
        all_latents = self.vae.encode_latents(outfit_images.to(dev))               # :435-437

        st = self.pipe.begin(olists=olists_c, all_latents=all_latents, category_prompts=category_prompts,
                             null_prompt=null_prompt, hist_latents=hist_latents, null_latent=null_latent,
                             init_latents=latents, num_inference_steps=num_inference_steps,
                             category_guidance_scale=category_guidance_scale, hist_guidance_scale=hist_guidance_scale,
                             mutual_guidance_scale=mutual_guidance_scale, device=dev)
        for i, t in enumerate(st.timesteps):                                       # :456-577
            self.pipe.step(st, t, ddim_eta=eta, generator=generator)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, st.latents)
        latents = st.latents

        if output_type == "latent":                                                # :579-592
            image = latents.clone()
        elif output_type in ("pil", "uint8"):
            u8 = self.vae.decode_latents_uint8(latents).cpu().numpy()
            if output_type == "pil":
                from PIL import Image
                image = [Image.fromarray(a) for a in u8]
            else:
                image = list(u8)                                                   # HWC uint8 arrays (what PIL would wrap)
        elif output_type in ("pt", "np"):
            image = (self.vae.decode_latents(latents) / 2 + 0.5).clamp(0, 1)       # VaeImageProcessor.denormalize
            if output_type == "np":
                image = image.cpu().permute(0, 2, 3, 1).float().numpy()
        else:
            raise ValueError(f"unknown output_type {output_type!r} (pil | uint8 | np | pt | latent)")

        if not return_dict:                                                        # :598-614
            all_results: Dict[int, Dict[int, Dict[str, Any]]] = {}
            for i, uid in enumerate(fill_uids.tolist()):
                oid = int(fill_oids[i])
                ent = all_results.setdefault(uid, {}).setdefault(oid, dict(images=[], cates=[], full_cates=full_cate[i]))
                ent["images"].append(image[i])
                ent["cates"].append(fill_cate[i])
                ent["outfits"] = olists_c[fill_idx[i][0]]
            return all_results, init_latents
        return (StableDiffusionPipelineOutput(images=image, nsfw_content_detected=None), fill_uids, fill_oids, fill_cate,
                full_cate, init_latents)
