"""Generate the small golden fixtures under tests/golden/ from the CPU oracle (fixed seeds).

The reference itself cannot be executed here (it imports diffusers 0.18.2, absent), so these vectors pin
the ORACLE (regression) and give the GPU tests seed-independent targets; they do not pin the oracle to the
reference ("parity unpinned", see oracle/__init__.py).   Usage: python tools/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.generation_oracle import make_oracle_mutual_encoder, oracle_generation  # noqa: E402
from oracle.schedulers_oracle import OracleDDIMScheduler, OraclePNDMScheduler  # noqa: E402
from oracle.unet_oracle import make_oracle_unet, tiny_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tiny_unet_case():
    cfg = tiny_config()
    unet = make_oracle_unet(cfg, seed=0)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 8, 16, 16, generator=g)
    ctx = torch.randn(2, 77, 64, generator=g)
    y = unet(x, torch.tensor(501), ctx)
    return dict(x=x, ctx=ctx, t=501, y=y)


def scheduler_case():
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    eps = [torch.randn(2, 4, 8, 8, generator=g) for _ in range(7)]
    out = {}
    for name, cls, n in (("ddim", OracleDDIMScheduler, 6), ("pndm", OraclePNDMScheduler, 7)):
        s = cls()
        s.set_timesteps(6)
        x, traj = x0.clone(), []
        for i, t in enumerate(s.timesteps[:n]):
            x = s.step(eps[i], t, x)[0]
            traj.append(x.clone())
        out[name] = dict(timesteps=s.timesteps.clone(), traj=torch.stack(traj))
    return dict(x0=x0, eps=torch.stack(eps), **out)


def generation_case():
    cfg = tiny_config()
    unet = make_oracle_unet(cfg, seed=0)
    me = make_oracle_mutual_encoder(seed=1, latent_size=16, hid_dim=64)
    g = torch.Generator().manual_seed(13)
    olists = torch.tensor([[0, 0, 0, 0], [3, 0, 7, 0]])
    n = 6
    inp = dict(olists=olists, all_latents=0.9 * torch.randn(8, 4, 16, 16, generator=g),
               category_prompts=torch.randn(n, 77, 64, generator=g), null_prompt=torch.randn(1, 77, 64, generator=g),
               hist_latents=0.9 * torch.randn(n, 4, 16, 16, generator=g), null_latent=0.9 * torch.randn(4, 16, 16, generator=g),
               init_latents=torch.randn(n, 4, 16, 16, generator=g))
    rec = []
    lat = oracle_generation(unet, me, OracleDDIMScheduler(), **inp, num_inference_steps=50, max_steps=3, record=rec)
    return dict(inputs=inp, latents=lat, eps_step0=rec[0]["noise_pred"], eps_branches_step0=rec[0]["noise_pred_branches"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.save(tiny_unet_case(), os.path.join(OUT, "tiny_unet.pt"))
    torch.save(scheduler_case(), os.path.join(OUT, "schedulers.pt"))
    torch.save(generation_case(), os.path.join(OUT, "generation_tiny.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
