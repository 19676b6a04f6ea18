"""Oracle pinning kit — run this ONCE on any machine that has ``diffusers==0.18.2`` (the reference's pin, README.md:27; CPU is
enough, ~1 minute) and commit the three files it writes under ``tests/golden/``:

    pip install diffusers==0.18.2 torch
    python tools/dump_diffusers_golden.py            # -> tests/golden/diffusers_{unet,schedulers,vae}.pt  (~6 MB)

``tests/test_oracle_cpu.py::test_oracle_matches_diffusers_golden_vectors`` then compares ``oracle/unet_oracle.py``,
``oracle/schedulers_oracle.py`` and ``oracle/vae_oracle.py`` with them at 1e-5 (state dicts are exchanged by name: the
oracle uses diffusers' key names) — which turns the "parity unpinned" of DESIGN.md §2 into "pinned to diffusers 0.18.2" for
the UNet, both schedulers and the VAE.  Until the files exist the test skips.  diffusers is not installable in the build
container (no network, not in the wheelhouse), which is why this is a kit and not a fixture.

What is dumped (small configurations with the SD-1.5 structure; seeds fixed):
  * UNet2DConditionModel, tiny config (the oracle's ``tiny_config``: block_out_channels (64, 128, 128, 128), attention_head_dim 2,
    cross_attention_dim 64, in_channels 8): state dict, sample [2, 8, 16, 16], timestep 501, encoder_hidden_states [2, 77, 64]
    -> ``.sample``; plus the same with SD-2's ``use_linear_projection=True`` and per-level heads.
  * DDIMScheduler / PNDMScheduler (``skip_prk_steps=True``) with the SD scheduler_config (scaled_linear 0.00085..0.012,
    steps_offset 1, set_alpha_to_one False): ``set_timesteps(6)`` -> timesteps and the trajectory of ``step`` over fixed eps.
  * AutoencoderKL, tiny config: state dict, ``encode(x).latent_dist.mode()``, ``decode(z).sample``.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    import diffusers
    from diffusers import AutoencoderKL, DDIMScheduler, PNDMScheduler, UNet2DConditionModel
    print("diffusers", diffusers.__version__)
    if diffusers.__version__ != "0.18.2":
        print("WARNING: the reference pins diffusers==0.18.2; vectors from another version pin the oracle to THAT version")
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # ---------------- UNet ----------------
    unet_cases = {}
    for name, extra in (("tiny", dict(attention_head_dim=2, use_linear_projection=False)),
                        ("tiny_sd2", dict(attention_head_dim=(1, 2, 2, 2), use_linear_projection=True, cross_attention_dim=96))):
        cfg = dict(sample_size=16, in_channels=8, out_channels=4, block_out_channels=(64, 128, 128, 128), cross_attention_dim=64,
                   layers_per_block=2, norm_num_groups=32,
                   down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                   up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"))
        cfg.update(extra)
        torch.manual_seed(1)
        m = UNet2DConditionModel(**cfg).eval()
        g = torch.Generator().manual_seed(7)
        x = torch.randn(2, 8, 16, 16, generator=g)
        ctx = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g)
        with torch.no_grad():
            y = m(x, 501, encoder_hidden_states=ctx).sample
            y_vec_t = m(x, torch.tensor([501, 20]), encoder_hidden_states=ctx).sample
        unet_cases[name] = dict(config={k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
                                state_dict={k: v.clone() for k, v in m.state_dict().items()}, x=x, ctx=ctx, t=501, y=y,
                                t_vec=[501, 20], y_vec_t=y_vec_t)
    torch.save(dict(diffusers=diffusers.__version__, cases=unet_cases), os.path.join(OUT, "diffusers_unet.pt"))

    # ---------------- schedulers ----------------
    sc = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1,
              set_alpha_to_one=False)
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    eps = torch.stack([torch.randn(2, 4, 8, 8, generator=g) for _ in range(52)])
    sched = {}
    for name, s, n_steps in (("ddim", DDIMScheduler(clip_sample=False, **sc), (6, 50)), ("pndm", PNDMScheduler(skip_prk_steps=True, **sc), (6, 50))):
        for n in n_steps:
            s.set_timesteps(n)
            x, traj = x0.clone(), []
            for i, t in enumerate(s.timesteps):
                x = s.step(eps[i], t, x).prev_sample
                traj.append(x.clone())
            sched[f"{name}_{n}"] = dict(timesteps=s.timesteps.clone(), traj=torch.stack(traj), init_noise_sigma=float(s.init_noise_sigma))
    torch.save(dict(diffusers=diffusers.__version__, x0=x0, eps=eps, **sched), os.path.join(OUT, "diffusers_schedulers.pt"))

    # ---------------- VAE ----------------
    vcfg = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(32, 64, 64, 64), layers_per_block=2,
                norm_num_groups=32, down_block_types=("DownEncoderBlock2D",) * 4, up_block_types=("UpDecoderBlock2D",) * 4,
                scaling_factor=0.18215, sample_size=64)
    torch.manual_seed(3)
    vae = AutoencoderKL(**vcfg).eval()
    g = torch.Generator().manual_seed(5)
    img = torch.randn(2, 3, 64, 64, generator=g).clamp(-1, 1)
    z = torch.randn(2, 4, 8, 8, generator=g)
    with torch.no_grad():
        mode = vae.encode(img).latent_dist.mode()
        dec = vae.decode(z).sample
    torch.save(dict(diffusers=diffusers.__version__, config={k: (list(v) if isinstance(v, tuple) else v) for k, v in vcfg.items()},
                    state_dict={k: v.clone() for k, v in vae.state_dict().items()}, img=img, mode=mode, z=z, dec=dec),
               os.path.join(OUT, "diffusers_vae.pt"))
    for f in sorted(os.listdir(OUT)):
        if f.startswith("diffusers_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    try:
        main()
    except ImportError as e:
        sys.exit(f"this script needs diffusers==0.18.2 ({e}); see its docstring")
