"""Diagnostic for tests/test_unet_gpu.py::test_shared_cfg_prefix_is_bitwise_the_full_computation (red on the round-1 driver
box for scales (12, 4, 5), flags (use_history=True, use_mutual_guidance=False), key (share=True, max_rows=256, graph=True)).

1. determinism + cross-arm equality of the generation arms, each arm run REPS times with a fresh pipeline;
2. eager forward_nhwc with every intermediate tapped (fine_taps), shared_tail = 0 against shared_tail = k: first differing
   intermediate, rows, element counts.
Usage: python tools/shared_prefix_diag.py [REPS]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_unet_gpu import _mk, _gen_inputs
from difashion_b200 import ops
from difashion_b200.mutual import MutualEncoder
from difashion_b200.pipeline import B200DiFashionPipeline
from difashion_b200.schedulers import B200DDIMScheduler

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
oracle, unet = _mk("tiny")
cfg = oracle.cfg
torch.manual_seed(0)
me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
olists = torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0], [0, 0, 0, 0]])
inp = _gen_inputs(cfg, olists)


def describe(a, b):
    d = (a.float() - b.float()).abs()
    nz = d > 0
    rows = nz.reshape(a.shape[0], -1).any(1).nonzero().flatten().tolist()
    return f"{int(nz.sum())}/{nz.numel()} elements differ, max {float(d.max()):.3e}, rows {rows[:12]}"


for scales, flags in (((12.0, 4.0, 5.0), (True, False)), ((12.0, 4.0, 5.0), (True, True)), ((12.0, 4.0, 1.0), (True, True))):
    print(f"== scales {scales} flags {flags}", flush=True)
    arms = ((False, 256, True), (True, 256, True), (True, 12, True), (True, 256, False), (False, 256, False))
    res = {}
    for rep in range(REPS):
        for arm in arms:
            share, max_rows, graph = arm
            pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), use_history=flags[0], use_mutual_guidance=flags[1],
                                         max_rows=max_rows, use_cuda_graph=graph, share_cfg_prefix=share)
            rec = []
            lat = pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda", category_guidance_scale=scales[0],
                                hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2], record=rec).clone()
            st = pipe._states[next(iter(pipe._states))]
            eps = [torch.cat([e.reshape(st.nb, -1, *e.shape[1:]) for e in r["eps_branches"]], 1).clone() for r in rec]
            torch.cuda.synchronize()
            res.setdefault(arm, []).append((lat, eps))
    base_lat, base_eps = res[arms[0]][0]
    for arm in arms:
        for rep, (lat, eps) in enumerate(res[arm]):
            bad = [i for i, (a, b) in enumerate(zip(eps, base_eps)) if not torch.equal(a, b)]
            if bad or not torch.equal(lat, base_lat):
                i = bad[0] if bad else -1
                msg = describe(eps[i].reshape(-1, *eps[i].shape[2:]), base_eps[i].reshape(-1, *eps[i].shape[2:])) if bad else ""
                print(f"   arm {arm} rep {rep}: DIFFERS from arm {arms[0]} rep 0 — eps steps {bad}; first: {msg}; latents: "
                      f"{describe(lat, base_lat)}", flush=True)
            else:
                print(f"   arm {arm} rep {rep}: bitwise equal", flush=True)

# ---- eager taps: shared_tail = 0 vs k on the UNet input the blend kernel makes for flags (True, False)
print("== eager fine taps, shared_tail 0 vs k", flush=True)
for n_items, nb in ((8, 3), (8, 4), (4, 3), (16, 4), (5, 3)):
    B, k = nb * n_items, n_items
    g = torch.Generator().manual_seed(11 + n_items)
    x = torch.randn(B - k, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    x = torch.cat([x, x[-k:]]).cuda()
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g).cuda()
    t = torch.full((B,), 981.0, device="cuda")
    x_in = torch.empty(B, cfg.sample_size, cfg.sample_size, cfg.in_channels, dtype=torch.bfloat16, device="cuda")
    ops.nchw_to_nhwc_bf16(x, x_in)
    c, kv = unet.set_context(ctx)
    for rep in range(3):
        t0, t1 = {}, {}
        ws0 = unet.workspace(("diag0", B), torch.device("cuda"))
        ws1 = unet.workspace(("diag1", B), torch.device("cuda"))
        full = unet.forward_nhwc(x_in, t, c, kv, ws0, taps=t0, fine_taps=True).clone()
        shared = unet.forward_nhwc(x_in, t, c, kv, ws1, taps=t1, shared_tail=k, fine_taps=True).clone()
        torch.cuda.synchronize()
        first = None
        for key in t0:
            a, b = t0[key], t1.get(key)
            if b is None:
                continue
            if a.shape != b.shape:                      # prefix tensors hold B - k rows in the shared arm
                rows = min(a.shape[0], b.shape[0])
                a, b = a[:rows], b[:rows]
            if not torch.equal(a, b):
                first = (key, describe(a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)))
                break
        print(f"   B={B} k={k} rep {rep}: eps equal {torch.equal(full, shared)}; first differing tap: {first}", flush=True)
