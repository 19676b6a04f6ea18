"""DDIM and PNDM(PLMS) schedulers with the diffusers 0.18.2 contract DiFashion consumes
(``set_timesteps / timesteps / order / init_noise_sigma / scale_model_input / step / alphas_cumprod /
add_noise / get_velocity / config``; reference ``DiFashion/models/difashion.py:64, :154-158, :241-244,
:356-357, :472, :569, :632, :659-674``).

Both schedulers reduce to ``x_prev = cx * x + sum_k ck * eps_k (+ cn * noise)`` with host-side scalar
coefficients (float64 on the host, from the fp32 ``alphas_cumprod`` table), so the whole update is ONE
streaming kernel (``dfb_cfg_step``), optionally fused with the 4-branch classifier-free-guidance combine
(``cfg_step``).  ``step`` keeps diffusers' signature — DiFashion probes it with ``inspect.signature`` for
``eta`` / ``generator`` (``difashion.py:665-673``), so the parameter names matter.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .config import FrozenConfig


def _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule) -> torch.Tensor:
    if beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    else:
        raise NotImplementedError(beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


class _SchedulerBase:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", **extra):
        if prediction_type != "epsilon":
            raise NotImplementedError("only epsilon prediction is on the DiFashion path")
        self.config = FrozenConfig(dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                        beta_schedule=beta_schedule, set_alpha_to_one=set_alpha_to_one,
                                        steps_offset=steps_offset, prediction_type=prediction_type, **extra))
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def _a(self, t: int) -> float:
        return float(self.alphas_cumprod[t]) if t >= 0 else float(self.final_alpha_cumprod)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw):
        """``PNDMScheduler.from_pretrained(dir, subfolder="scheduler")`` (difashion.py:64): reads ``scheduler_config.json``
        (the same file serves either class, as diffusers' ``from_config`` allows)."""
        from . import checkpoint as ck
        cfg = ck.read_config(ck.model_dir(path, subfolder), "scheduler_config.json")
        if cfg.get("trained_betas") is not None:
            raise NotImplementedError("trained_betas")
        if cfg.get("timestep_spacing", "leading") != "leading":
            raise NotImplementedError(f"timestep_spacing={cfg['timestep_spacing']!r} (diffusers 0.18.2 / SD configs use 'leading')")
        if cfg.get("clip_sample", False) and cls.__name__ == "B200DDIMScheduler":
            raise NotImplementedError("clip_sample=True")
        cfg.update(kw)
        return cls(**{k: v for k, v in cfg.items() if k not in ("trained_betas",)})

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        ac = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < sample.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * noise - sb * sample


class B200DDIMScheduler(_SchedulerBase):
    """diffusers ``DDIMScheduler`` (leading spacing, ``clip_sample=False``)."""

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def coefficients(self, timestep: int, eta: float = 0.0) -> Tuple[float, float, float]:
        """(cx, c_eps, c_noise) with x_prev = cx*x + c_eps*eps + c_noise*z."""
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t, a_p = self._a(t), self._a(prev_t)
        var = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * math.sqrt(max(var, 0.0))
        cx = math.sqrt(a_p / a_t)
        c_eps = math.sqrt(max(1 - a_p - std * std, 0.0)) - math.sqrt(a_p * (1 - a_t) / a_t)
        return cx, c_eps, std

    def cfg_step(self, eps_nhwc: torch.Tensor, weights: Sequence[float], timestep, sample: torch.Tensor,
                 eta: float = 0.0, generator=None, variance_noise=None, out: Optional[torch.Tensor] = None,
                 eps_nchw: bool = False):
        """Fused CFG combine + DDIM update.  eps_nhwc: fp32 NHWC [nb*N,H,W,4] straight from the UNet kernels."""
        cx, c_eps, std = self.coefficients(int(timestep), eta)
        noise = None
        if eta > 0:
            noise = variance_noise if variance_noise is not None else torch.randn(
                sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        return ops.cfg_step(eps_nhwc, weights, sample, cx, [c_eps], noise=noise, cn=std, x_out=out, eps_nchw=eps_nchw)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        prev = self.cfg_step(_nchw_f32(model_output), [1.0], timestep, sample.float().contiguous(), eta, generator,
                             variance_noise, eps_nchw=True)
        prev = prev.to(sample.dtype)
        return (prev,) if not return_dict else FrozenConfig(prev_sample=prev)


def _nchw_f32(model_output: torch.Tensor) -> torch.Tensor:
    """Public ``step`` path: the model output arrives NCHW (diffusers convention); the kernel reads it in place."""
    if not model_output.is_cuda:
        raise RuntimeError("B200 schedulers need CUDA tensors: there is no CPU fallback")
    return model_output.float().contiguous()


class B200PNDMScheduler(_SchedulerBase):
    """diffusers ``PNDMScheduler`` with ``skip_prk_steps=True`` (PLMS; the reference default,
    ``difashion.py:64``): n+1 model calls for n inference steps."""

    def __init__(self, skip_prk_steps=True, **kw):
        if not skip_prk_steps:
            raise NotImplementedError("Stable Diffusion's PNDM config uses skip_prk_steps=True")
        super().__init__(skip_prk_steps=skip_prk_steps, **kw)
        self.ets: List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample: Optional[torch.Tensor] = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        _ts = (np.arange(0, num_inference_steps) * ratio).round() + self.config.steps_offset
        plms = np.concatenate([_ts[:-1], _ts[-2:-1], _ts[-1:]])[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(plms).to(device) if device is not None else torch.from_numpy(plms)
        self.ets, self.counter, self.cur_sample = [], 0, None

    def _plan(self, timestep: int):
        """Host-side PLMS bookkeeping for this call: (t, prev_t, weights on [current eps, ets[-1], ets[-2], ets[-3]],
        append current?, use cur_sample?)."""
        t = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        prev_t = t - ratio
        append = self.counter != 1
        n_hist = len(self.ets)
        if not append:
            prev_t, t = t, t + ratio
        n_after = min(n_hist, 3) + 1 if append else n_hist
        if n_after == 1 and self.counter == 0:
            w = [1.0]
            use_cur = False
        elif n_after == 1 and self.counter == 1:
            w = [0.5, 0.5]                  # (model_output + ets[-1]) / 2 ; ets[-1] is the stored first eps
            use_cur = True
        elif n_after == 2:
            w = [1.5, -0.5]
            use_cur = False
        elif n_after == 3:
            w = [23 / 12, -16 / 12, 5 / 12]
            use_cur = False
        else:
            w = [55 / 24, -59 / 24, 37 / 24, -9 / 24]
            use_cur = False
        return t, prev_t, w, append, use_cur

    def cfg_step(self, eps_nhwc: torch.Tensor, weights: Sequence[float], timestep, sample: torch.Tensor,
                 out: Optional[torch.Tensor] = None, eps_nchw: bool = False):
        """Fused CFG combine + PLMS update; keeps the post-CFG eps history (<= 4 tensors) on device."""
        t, prev_t, w, append, use_cur = self._plan(timestep)
        a_t, a_p = self._a(t), self._a(prev_t)
        cx = math.sqrt(a_p / a_t)
        ce = -(a_p - a_t) / (a_t * math.sqrt(1 - a_p) + math.sqrt(a_t * (1 - a_t) * a_p))
        if append:
            hist = list(reversed(self.ets[-3:]))            # ets[-1], ets[-2], ets[-3] BEFORE appending the current
        else:
            hist = [self.ets[-1]]
        ck = [ce * wk for wk in w] + [0.0] * (4 - len(w))
        hist = (hist + [None] * 3)[:3]
        for k in range(3):
            if ck[k + 1] == 0.0:
                hist[k] = None
        x_src = self.cur_sample if use_cur else sample
        eps_out = torch.empty_like(sample) if append else None
        if self.counter == 0:
            self.cur_sample = sample.clone()
        prev = ops.cfg_step(eps_nhwc, weights, x_src, cx, ck, hist=hist, x_out=out, eps_out=eps_out, eps_nchw=eps_nchw)
        if append:
            self.ets = self.ets[-3:] + [eps_out]
        if use_cur:
            self.cur_sample = None
        self.counter += 1
        return prev

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        prev = self.cfg_step(_nchw_f32(model_output), [1.0], timestep, sample.float().contiguous(), eps_nchw=True)
        prev = prev.to(sample.dtype)
        return (prev,) if not return_dict else FrozenConfig(prev_sample=prev)

    def step_plms(self, model_output, timestep, sample, return_dict: bool = True):
        return self.step(model_output, timestep, sample, return_dict=return_dict)
