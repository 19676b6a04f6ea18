"""Second-stage diagnostic (round 2): WHICH intermediate differs.
(A) the two eager pipeline arms (share_cfg_prefix off / on) with every intermediate of forward_nhwc tapped at every step:
    first tap whose rows differ; for prefix tensors (B - k rows in the shared arm) also the full arm's tail rows against the
    shared arm's rows they are copied from.
(B) batch invariance with fine taps: rows [0, b) of a 48-row batch against the same rows run alone, b = 5, 7, 3, 1."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_unet_gpu import _mk, _gen_inputs
from difashion_b200.mutual import MutualEncoder
from difashion_b200.pipeline import B200DiFashionPipeline
from difashion_b200.schedulers import B200DDIMScheduler

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
    return f"{int(nz.sum())}/{nz.numel()} differ, max {float(d.max()):.3e}, rows {rows[:10]}"


orig_forward = unet.forward_nhwc.__func__


def run_arm(share, scales, flags):
    log = []

    def patched(self, x_in, t_dev, ctx, kv, ws, taps=None, shared_tail=0, fine_taps=False):
        t = {"x_in": x_in.clone()}
        out = orig_forward(self, x_in, t_dev, ctx, kv, ws, taps=t, shared_tail=shared_tail, fine_taps=True)
        t["eps"] = out.clone()
        log.append((shared_tail, t))
        return out

    unet.forward_nhwc = patched.__get__(unet)
    try:
        pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), use_history=flags[0], use_mutual_guidance=flags[1],
                                     max_rows=256, use_cuda_graph=False, share_cfg_prefix=share)
        pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda", category_guidance_scale=scales[0],
                      hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2])
        torch.cuda.synchronize()
    finally:
        del unet.forward_nhwc
    return log


for scales, flags in (((12.0, 4.0, 5.0), (True, False)), ((12.0, 4.0, 1.0), (True, True))):
    print(f"== (A) scales {scales} flags {flags}", flush=True)
    full, shared = run_arm(False, scales, flags), run_arm(True, scales, flags)
    for step, ((_, tf), (k, ts)) in enumerate(zip(full, shared)):
        B = tf["x_in"].shape[0]
        print(f"  step {step}: B={B} shared_tail={k}; x_in tail == the rows before it (full arm): "
              f"{torch.equal(tf['x_in'][B - k:], tf['x_in'][B - 2 * k:B - k])}", flush=True)
        shown = 0
        for key in tf:
            a, b = tf[key], ts.get(key)
            if b is None:
                continue
            a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
            if a.shape == b.shape:
                if not torch.equal(a, b):
                    per = a.shape[0] // B if a.shape[0] % B == 0 else 0
                    print(f"    {key} {tuple(a.shape)}: {describe(a2, b2)}" + (f" (rows per batch row: {per})" if per else ""), flush=True)
                    shown += 1
            else:       # prefix tensor: the shared arm holds (B - k) batch rows
                per = b.shape[0] // (B - k)
                head_eq = torch.equal(a2[:b2.shape[0]], b2)
                tail_eq = torch.equal(a2[b2.shape[0]:], b2[b2.shape[0] - k * per:])
                if not (head_eq and tail_eq):
                    print(f"    {key} prefix {tuple(a.shape)} vs {tuple(b.shape)}: head equal {head_eq}; full-arm tail vs shared-arm source rows: "
                          f"{describe(a2[b2.shape[0]:], b2[b2.shape[0] - k * per:])} (rows per batch row: {per})", flush=True)
                    shown += 1
            if shown >= 6:
                break
        if shown:
            break

print("== (B) batch invariance with fine taps", flush=True)
g = torch.Generator().manual_seed(7)
Bfull = 48
x = torch.randn(Bfull, cfg.sample_size, cfg.sample_size, cfg.in_channels, generator=g).bfloat16().cuda()
ctx = torch.randn(Bfull, 77, cfg.cross_attention_dim, generator=g).cuda()
t = torch.full((Bfull,), 981.0, device="cuda")


def run(b):
    ws = unet.workspace(("diag2", b), torch.device("cuda"))
    c, kv = unet.set_context(ctx[:b].contiguous())
    taps = {}
    out = unet.forward_nhwc(x[:b].contiguous(), t[:b], c, kv, ws, taps=taps, fine_taps=True).clone()
    torch.cuda.synchronize()
    taps["eps"] = out
    return taps


full = run(Bfull)
for b in (5, 7, 3, 1, 6, 40):
    part = run(b)
    shown = 0
    for key in full:
        a, p = full[key], part[key]
        per = p.shape[0] // b if p.shape[0] % b == 0 else None
        if per is None:
            continue
        a2, p2 = a.reshape(a.shape[0], -1)[:p.shape[0]], p.reshape(p.shape[0], -1)
        if a2.shape != p2.shape:
            continue
        if not torch.equal(a2, p2):
            print(f"  b={b}: {key} {tuple(p.shape)}: {describe(a2, p2)} (rows per batch row: {per})", flush=True)
            shown += 1
            if shown >= 4:
                break
    if not shown:
        print(f"  b={b}: all taps bitwise equal", flush=True)
