"""GPU parity of the full UNet forward and of the generation loop against the CPU oracle
(oracle/unet_oracle.py, oracle/generation_oracle.py) on identical random-init weights / inputs.
Tolerances are BASELINE.json's: per-step noise-prediction rel-L2 <= 1e-2 (bf16 tensor-core operands,
fp32 accumulate / residual stream), final-latent cosine >= 0.999."""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _mk(cfg_kw, seed=0):
    from oracle.unet_oracle import UNetConfig, make_oracle_unet, tiny_config
    from difashion_b200.unet import B200UNet2DConditionModel
    # "sd2*": the reference's DEFAULT base model, stabilityai/stable-diffusion-2-base (train.py:44, inf4eval.py:65):
    # linear proj_in / proj_out, per-level head counts (5, 10, 20, 20) -> head dim 64, cross-attention dim 1024
    ocfg = {"tiny": tiny_config,
            "tiny_sd2": lambda: tiny_config(use_linear_projection=True, attention_head_dim=(1, 2, 2, 2), cross_attention_dim=96),
            "sd2": lambda: UNetConfig(use_linear_projection=True, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024),
            "full": UNetConfig}[cfg_kw]()
    oracle = make_oracle_unet(ocfg, seed=seed)
    kw = dict(sample_size=ocfg.sample_size, in_channels=ocfg.in_channels, out_channels=ocfg.out_channels,
              block_out_channels=tuple(ocfg.block_out_channels), cross_attention_dim=ocfg.cross_attention_dim,
              attention_head_dim=ocfg.attention_head_dim, use_linear_projection=ocfg.use_linear_projection)
    unet = B200UNet2DConditionModel(**kw)
    missing = unet.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, unet.cuda()


def _block_report(taps_o, taps_g):
    out = []
    for k, v in taps_o.items():
        if k in taps_g:
            out.append(f"{k}: {rel_l2(taps_g[k].permute(0, 3, 1, 2).cpu(), v):.2e}")
    return " | ".join(out)


@pytest.mark.parametrize("which,B,S", [("tiny", 4, 77), ("tiny", 3, 85), ("full", 2, 77), ("tiny_sd2", 3, 77), ("sd2", 1, 77)])
def test_unet_forward_matches_oracle(which, B, S):
    oracle, unet = _mk(which)
    cfg = oracle.cfg
    g = torch.Generator().manual_seed(123)
    x = torch.randn(B, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    ctx = torch.randn(B, S, cfg.cross_attention_dim, generator=g)
    t = torch.tensor(981)
    taps_o = {}
    ref = oracle(x, t, ctx, taps=taps_o)
    # block-level taps through the NHWC entry point
    from difashion_b200 import ops
    ws = unet.workspace(("test", B), torch.device("cuda"))
    x_in = torch.empty(B, cfg.sample_size, cfg.sample_size, cfg.in_channels, dtype=torch.bfloat16, device="cuda")
    ops.nchw_to_nhwc_bf16(x.cuda(), x_in)
    taps_g = {}
    ctx_dev = ctx.cuda()
    unet.forward_nhwc(x_in, torch.full((B,), 981.0, device="cuda"), *unet.set_context(ctx_dev), ws, taps=taps_g)
    torch.cuda.synchronize()
    report = _block_report(taps_o, taps_g)
    # public diffusers-style call
    got = unet(x.cuda(), t.cuda(), encoder_hidden_states=ctx_dev, return_dict=False)[0]
    out2 = unet(x.cuda(), 981, ctx_dev).sample
    torch.cuda.synchronize()
    e = rel_l2(got.cpu(), ref)
    print(f"\n[{which} B={B}] eps rel-L2 {e:.3e}; per-block: {report}")
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert e <= 1e-2, report
    assert torch.equal(got, out2)
    for i in range(B):
        assert rel_l2(got[i].cpu(), ref[i]) <= 1e-2


def _gen_inputs(cfg, olists, S=77, seed=123):
    g = torch.Generator().manual_seed(seed)
    bsz, olen = olists.shape
    n = int((olists == 0).sum())
    s = cfg.sample_size
    return dict(
        olists=olists,
        all_latents=0.9 * torch.randn(bsz * olen, 4, s, s, generator=g),
        category_prompts=torch.randn(n, S, cfg.cross_attention_dim, generator=g),
        null_prompt=torch.randn(1, S, cfg.cross_attention_dim, generator=g),
        hist_latents=0.9 * torch.randn(n, 4, s, s, generator=g),
        null_latent=0.9 * torch.randn(4, s, s, generator=g),
        init_latents=torch.randn(n, 4, s, s, generator=g),
    )


def _run_both(which, olists, steps, sched_name, scales=(12.0, 4.0, 5.0), total_steps=None, use_graph=True, flags=(True, True)):
    from oracle.generation_oracle import make_oracle_mutual_encoder, oracle_generation
    from oracle.schedulers_oracle import OracleDDIMScheduler, OraclePNDMScheduler
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler, B200PNDMScheduler
    oracle, unet = _mk(which)
    cfg = oracle.cfg
    hid = 64 if which == "tiny" else 256
    ome = make_oracle_mutual_encoder(seed=1, latent_size=cfg.sample_size, hid_dim=hid)
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=hid)
    me.load_state_dict(ome.state_dict())
    inp = _gen_inputs(cfg, olists)
    total = total_steps or steps
    osched = OracleDDIMScheduler() if sched_name == "ddim" else OraclePNDMScheduler()
    rec_o = []
    lat_o = oracle_generation(oracle, ome, osched, **inp, num_inference_steps=total, category_guidance_scale=scales[0],
                              hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2], use_history=flags[0],
                              use_mutual_guidance=flags[1], record=rec_o, max_steps=steps)
    sched = B200DDIMScheduler() if sched_name == "ddim" else B200PNDMScheduler()
    pipe = B200DiFashionPipeline(unet, me.cuda(), sched, eta_mutual=0.1, use_history=flags[0], use_mutual_guidance=flags[1],
                                 use_cuda_graph=use_graph)
    rec_g = []
    dev_inp = {k: (v.cuda() if k != "olists" else v) for k, v in inp.items()}
    lat_g = pipe.generate(**dev_inp, num_inference_steps=total, category_guidance_scale=scales[0],
                          hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2], record=rec_g, max_steps=steps)
    torch.cuda.synchronize()
    return lat_o, lat_g.cpu(), rec_o, rec_g


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("sched,use_graph", [("ddim", True), ("ddim", False), ("pndm", True)])
def test_generation_gor_tiny_50_steps(sched, use_graph):
    """GOR: 2 outfits x 4 generated items, 50 steps, 4-branch CFG: final-latent cosine >= 0.999."""
    olists = torch.zeros(2, 4, dtype=torch.long)
    lat_o, lat_g, rec_o, rec_g = _run_both("tiny", olists, 51 if sched == "pndm" else 50, sched, total_steps=50, use_graph=use_graph)
    # first step: identical inputs -> per-branch eps parity
    e0 = rel_l2(rec_g[0]["eps_branches"][0].permute(0, 3, 1, 2).cpu(), rec_o[0]["noise_pred_branches"])
    c = _cos(lat_g, lat_o)
    print(f"\n[{sched} graph={use_graph}] step-0 eps rel-L2 {e0:.3e}; final cosine {c:.6f}; final rel-L2 {rel_l2(lat_g, lat_o):.3e}")
    assert e0 <= 1e-2
    assert c >= 0.999


@pytest.mark.parametrize("scales,flags", [((12.0, 4.0, 5.0), (True, True)), ((12.0, 1.0, 5.0), (True, True)),
                                          ((12.0, 4.0, 1.0), (True, True)), ((12.0, 1.0, 1.0), (True, True)),
                                          ((1.0, 4.0, 1.0), (True, True)), ((1.0, 1.0, 5.0), (True, True)),
                                          ((1.0, 1.0, 1.0), (True, True)), ((12.0, 4.0, 5.0), (False, True)),
                                          ((12.0, 4.0, 5.0), (True, False)), ((1.0, 4.0, 5.0), (True, True))])
def test_generation_fitb_and_degenerate_cfg_variants(scales, flags):
    """FITB-style outfits (given items + blanks) through every CFG branch layout of difashion.py:533-566."""
    olists = torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0]])
    lat_o, lat_g, rec_o, rec_g = _run_both("tiny", olists, 3, "ddim", scales=scales, flags=flags)
    e0 = rel_l2(rec_g[0]["eps_branches"][0].permute(0, 3, 1, 2).cpu(), rec_o[0]["noise_pred_branches"])
    el = rel_l2(lat_g, lat_o)
    print(f"\n[scales={scales} flags={flags}] step-0 eps rel-L2 {e0:.3e}; latents rel-L2 after 3 steps {el:.3e}")
    assert e0 <= 1e-2
    assert _cos(lat_g, lat_o) >= 0.999


def test_generation_full_size_two_steps():
    """SD-1.5-shaped UNet, one FITB outfit (1 blank, 3 given items) -> 4 UNet rows, 2 DDIM steps of 50."""
    olists = torch.tensor([[11, 0, 7, 9]])
    lat_o, lat_g, rec_o, rec_g = _run_both("full", olists, 2, "ddim", total_steps=50)
    for i in range(2):
        if i == 0:
            e = rel_l2(rec_g[0]["eps_branches"][0].permute(0, 3, 1, 2).cpu(), rec_o[0]["noise_pred_branches"])
            print(f"\n[full] step-0 per-branch eps rel-L2 {e:.3e}")
            assert e <= 1e-2
    print(f"[full] latents rel-L2 after 2 steps {rel_l2(lat_g, lat_o):.3e} cosine {_cos(lat_g, lat_o):.6f}")
    assert _cos(lat_g, lat_o) >= 0.999


def test_against_committed_golden_fixtures():
    """Kernels vs the committed oracle fixtures (tests/golden, made by tools/make_golden.py)."""
    import os
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler, B200PNDMScheduler
    from oracle.generation_oracle import make_oracle_mutual_encoder
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    oracle, unet = _mk("tiny")
    gold = torch.load(os.path.join(gold_dir, "tiny_unet.pt"))
    got = unet(gold["x"].cuda(), gold["t"], gold["ctx"].cuda()).sample
    assert rel_l2(got.cpu(), gold["y"]) <= 1e-2
    # scheduler trajectories through the public step() API (fp32 streaming kernel: tight tolerance)
    sg = torch.load(os.path.join(gold_dir, "schedulers.pt"))
    for name, cls in (("ddim", B200DDIMScheduler), ("pndm", B200PNDMScheduler)):
        s = cls()
        s.set_timesteps(6)
        assert torch.equal(s.timesteps.cpu(), sg[name]["timesteps"])
        x = sg["x0"].cuda()
        for i in range(sg[name]["traj"].shape[0]):
            x = s.step(sg["eps"][i].cuda(), s.timesteps[i], x, return_dict=False)[0]
            assert rel_l2(x.cpu(), sg[name]["traj"][i]) < 1e-5, (name, i)
    # 3 generation steps (mixed GOR + FITB outfits)
    gg = torch.load(os.path.join(gold_dir, "generation_tiny.pt"))
    ome = make_oracle_mutual_encoder(seed=1, latent_size=16, hid_dim=64)
    me = MutualEncoder(latent_size=16, hid_dim=64)
    me.load_state_dict(ome.state_dict())
    pipe = B200DiFashionPipeline(unet, me.cuda(), B200DDIMScheduler())
    inp = {k: (v.cuda() if k != "olists" else v) for k, v in gg["inputs"].items()}
    rec = []
    lat = pipe.generate(**inp, num_inference_steps=50, max_steps=3, record=rec)
    assert rel_l2(rec[0]["eps_branches"][0].permute(0, 3, 1, 2).cpu(), gg["eps_branches_step0"]) <= 1e-2
    assert _cos(lat.cpu(), gg["latents"]) >= 0.999


def test_bitwise_reproducible():
    """All reductions run in a fixed order: two runs of the same step give identical bits."""
    oracle, unet = _mk("tiny")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 8, 16, 16, generator=g).cuda()
    ctx = torch.randn(4, 77, 64, generator=g).cuda()
    a = unet(x, 500, ctx).sample.clone()
    b = unet(x, 500, ctx).sample
    assert torch.equal(a, b)


def test_captured_graphs_follow_weight_changes():
    """inf4eval.py's loop over checkpoints: generate, load other weights into the SAME model, generate again.  The captured
    graphs hold pointers into the packed weights of the first checkpoint: the pipeline has to notice the re-pack and capture
    again (it used to replay the stale graph).  Second generation == a fresh model built from the second weights, bit for bit."""
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler
    oracle_a, unet = _mk("tiny", seed=0)
    oracle_b, unet_b = _mk("tiny", seed=5)
    cfg = oracle_a.cfg
    torch.manual_seed(3)
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
    inp = _gen_inputs(cfg, torch.tensor([[3, 0, 7, 9], [0, 0, 0, 0]]))
    pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler())
    first = pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda").clone()
    unet.load_state_dict(oracle_b.state_dict())
    second = pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda").clone()
    fresh = B200DiFashionPipeline(unet_b, me, B200DDIMScheduler()).generate(**inp, num_inference_steps=50, max_steps=3, device="cuda")
    torch.cuda.synchronize()
    assert not torch.equal(first, second)
    assert torch.equal(second, fresh)
    # eta is a kernel scalar inside the graph: a pipeline with another eta on the same UNet must not reuse anything stale
    pipe.eta_mutual = 0.3
    third = pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda").clone()
    fresh3 = B200DiFashionPipeline(unet_b, me, B200DDIMScheduler(), eta_mutual=0.3).generate(**inp, num_inference_steps=50, max_steps=3, device="cuda")
    assert torch.equal(third, fresh3) and not torch.equal(third, second)


def test_attn_processor_on_foreign_attention_module():
    """B200AttnProcessor implements the diffusers processor protocol for any module exposing
    to_q/to_k/to_v/to_out/heads/scale (here: the oracle's Attention, standing in for diffusers')."""
    from difashion_b200.attention import B200AttnProcessor
    from oracle.unet_oracle import Attention as OracleAttention
    torch.manual_seed(0)
    for cross in (None, 96):
        attn = OracleAttention(320, cross, 8, 40).eval()
        x = torch.randn(2, 256, 320)
        ctx = torch.randn(2, 77, cross) if cross else None
        ref = attn(x, ctx)
        proc = B200AttnProcessor()
        attn_cuda = attn.cuda()
        got = proc(attn_cuda, x.cuda(), encoder_hidden_states=None if ctx is None else ctx.cuda())
        assert got.shape == ref.shape and got.dtype == torch.float32
        assert rel_l2(got.cpu(), ref) <= 1e-2


def test_multi_stream_chunks_bitwise_equal_single_stream():
    """streams=2: the row chunks of a step run concurrently on two streams inside one graph (own workspaces);
    kernels are deterministic and rows independent, so the latents are bitwise those of the sequential path."""
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler
    oracle, unet = _mk("tiny")
    cfg = oracle.cfg
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
    olists = torch.zeros(5, 4, dtype=torch.long)                       # 20 items -> 80 rows; max_rows 16 -> 5 chunks
    inp = _gen_inputs(cfg, olists)
    outs = []
    for streams in (1, 2, 3):
        pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), max_rows=16, streams=streams)
        outs.append(pipe.generate(**inp, num_inference_steps=50, max_steps=4, device="cuda").clone())
        assert (streams > 1) == bool(pipe._states[next(iter(pipe._states))].multi)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("scales,flags,shares", [((12.0, 4.0, 5.0), (True, True), True), ((12.0, 1.0, 5.0), (True, True), True),
                                                 ((12.0, 4.0, 1.0), (True, True), True), ((12.0, 1.0, 1.0), (True, True), True),
                                                 ((1.0, 4.0, 1.0), (True, True), False), ((1.0, 1.0, 5.0), (True, True), False),
                                                 ((12.0, 4.0, 5.0), (False, True), True), ((12.0, 4.0, 5.0), (True, False), True)])
def test_shared_cfg_prefix_is_bitwise_the_full_computation(scales, flags, shares):
    """The last two CFG branches get the same UNet input and differ only in the prompt (difashion.py:388-431, :494-512):
    computing the UNet ahead of its first cross-attention once for both (forward_nhwc(shared_tail=...)) must give the very
    bits the full computation gives — per-branch eps of every step and the latents — for every branch layout, chunked or not,
    eager or captured."""
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler
    oracle, unet = _mk("tiny")
    cfg = oracle.cfg
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
    olists = torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0], [0, 0, 0, 0]])           # 8 blanks
    inp = _gen_inputs(cfg, olists)
    runs = {}
    for share, max_rows, graph in ((False, 256, True), (True, 256, True), (True, 12, True), (True, 256, False)):
        pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), use_history=flags[0], use_mutual_guidance=flags[1],
                                     max_rows=max_rows, use_cuda_graph=graph, share_cfg_prefix=share)
        rec = []
        lat = pipe.generate(**inp, num_inference_steps=50, max_steps=3, device="cuda", category_guidance_scale=scales[0],
                            hist_guidance_scale=scales[1], mutual_guidance_scale=scales[2], record=rec).clone()
        st = pipe._states[next(iter(pipe._states))]
        assert st.shared_tail == (share and shares)
        runs[(share, max_rows, graph)] = (lat, [torch.cat([e.reshape(st.nb, -1, *e.shape[1:]) for e in r["eps_branches"]], 1) for r in rec])
    base_lat, base_eps = runs[(False, 256, True)]
    for key, (lat, eps) in runs.items():
        assert torch.equal(lat, base_lat), key
        for a, b in zip(eps, base_eps):
            assert torch.equal(a, b), key


def test_shared_tail_rows_are_not_read():
    """forward_nhwc(shared_tail=k) promises not to read x_in[B-k:]: poison them and compare with the full computation."""
    from difashion_b200 import ops
    oracle, unet = _mk("tiny")
    cfg = oracle.cfg
    g = torch.Generator().manual_seed(11)
    k = 3
    x = torch.randn(2 * k, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g)
    x = torch.cat([x, x[k:]]).cuda()                                     # rows [2k, 3k) repeat rows [k, 2k)
    ctx = torch.randn(3 * k, 77, cfg.cross_attention_dim, generator=g).cuda()
    B = 3 * k
    t = torch.full((B,), 421.0, device="cuda")
    x_in = torch.empty(B, cfg.sample_size, cfg.sample_size, cfg.in_channels, dtype=torch.bfloat16, device="cuda")
    ops.nchw_to_nhwc_bf16(x, x_in)
    ws = unet.workspace(("tail", B), torch.device("cuda"))
    c, kv = unet.set_context(ctx)
    full = unet.forward_nhwc(x_in, t, c, kv, ws).clone()
    x_in[2 * k:] = float("nan")
    shared = unet.forward_nhwc(x_in, t, c, kv, ws, shared_tail=k).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(shared).all() and torch.equal(shared, full)
    assert not torch.equal(full[k:2 * k], full[2 * k:])                  # the two branches do differ (different prompts)


def test_plms_row_chunks_keep_their_own_history():
    """PNDM/PLMS (the reference's scheduler, difashion.py:64) carries four past noise predictions per sample: a batch split
    into row chunks must give the bits of the unsplit batch (7 model calls: warm-up, the repeated second timestep, 4th order)."""
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200PNDMScheduler
    oracle, unet = _mk("tiny")
    cfg = oracle.cfg
    me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
    olists = torch.zeros(2, 4, dtype=torch.long)                       # 8 items -> 32 rows
    inp = _gen_inputs(cfg, olists)
    outs = []
    for max_rows in (256, 16, 20):                                     # 1 chunk / 2 chunks of 4 items / chunks of 5 and 3
        pipe = B200DiFashionPipeline(unet, me, B200PNDMScheduler(), max_rows=max_rows)
        outs.append(pipe.generate(**inp, num_inference_steps=50, max_steps=7, device="cuda").clone())
        assert len(pipe._states[next(iter(pipe._states))].chunks) == {256: 1, 16: 2, 20: 2}[max_rows]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
