"""Restatement of diffusers 0.18.2 ``DDIMScheduler`` / ``PNDMScheduler`` (PLMS) in numpy + torch.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) — parity unpinned: diffusers is absent.

Reference call sites: ``DiFashion/models/difashion.py:64`` (PNDMScheduler.from_pretrained),
``:356-357`` (set_timesteps / timesteps), ``:472`` (scale_model_input), ``:569`` (step),
``:632`` (init_noise_sigma), ``:659-674`` (eta / generator probing by parameter name).
Scheduler config = Stable Diffusion's ``scheduler_config.json``.
Written step-by-step in the order diffusers computes (not in the fused ``c1*x + c2*eps``
form the CUDA kernel uses) so the two are independent derivations.
"""
from __future__ import annotations

import numpy as np
import torch


def make_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class OracleDDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, steps_offset=1, set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = make_alphas_cumprod(num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.num_inference_steps = None

    def set_timesteps(self, n: int, device=None):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None,
             variance_noise=None, return_dict: bool = False):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        prev = a_p ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + std * variance_noise
        return (prev,)


class OraclePNDMScheduler:
    """PLMS with ``skip_prk_steps=True`` (Stable Diffusion config): n+1 UNet calls for n steps."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, steps_offset=1, set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = make_alphas_cumprod(num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.ets, self.counter, self.cur_sample = [], 0, None
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n: int, device=None):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        _ts = (np.arange(0, n) * ratio).round() + self.steps_offset
        plms = np.concatenate([_ts[:-1], _ts[-2:-1], _ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets, self.counter, self.cur_sample = [], 0, None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _prev_sample(self, sample, t, prev_t, eps):
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * eps / denom

    def step(self, model_output, timestep, sample, return_dict: bool = False):
        t = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_t = t - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = t
            t = t + ratio
        if len(self.ets) == 1 and self.counter == 0:
            eps = model_output
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            eps = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            eps = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            eps = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            eps = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._prev_sample(sample, t, prev_t, eps)
        self.counter += 1
        return (prev,)
