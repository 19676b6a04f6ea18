"""A row's result must not depend on the batch it is computed in: the row chunks of a generation (``max_rows``) and the
outfit shards of a multi-GPU run then reproduce the unsplit batch bit for bit.  Two position-dependent roundings have been
found with these checks and tools/*_diag*.py on B200: the standalone GroupNorm statistics pass sized its blocks from the
batch (round 1), and the GEMM's generic epilogue folded the time-embedding row bias into the column bias only where a
32-row chunk lay inside one batch row — the last chunk of an odd batch of 4x4 / 2x2 images (round 2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_unet_rows_are_independent_of_the_batch_size():
    from tests.test_unet_gpu import _mk
    oracle, unet = _mk("tiny")
    cfg = oracle.cfg
    g = torch.Generator().manual_seed(7)
    B = 48
    x = torch.randn(B, cfg.sample_size, cfg.sample_size, cfg.in_channels, generator=g).bfloat16().cuda()
    ctx = torch.randn(B, 77, cfg.cross_attention_dim, generator=g).cuda()
    t = torch.full((B,), 981.0, device="cuda")

    def run(b):
        ws = unet.workspace(("inv", b), torch.device("cuda"))
        c, kv = unet.set_context(ctx[:b].contiguous())
        return unet.forward_nhwc(x[:b].contiguous(), t[:b], c, kv, ws).clone()

    full = run(B)
    for b in (40, 16, 5, 7, 1):                    # odd batches: a partial last 32-row epilogue chunk at the 4x4 / 2x2 levels
        part = run(b)
        torch.cuda.synchronize()
        assert torch.equal(full[:b], part), f"rows [0, {b}) differ between a batch of {B} and a batch of {b}"


def test_standalone_groupnorm_statistics_do_not_depend_on_the_batch():
    """dfb_groupnorm (the statistics pass used where the producing GEMM emits no partials: images with H*W % 32 != 0,
    the fp32 verification path): image i of a 48-image launch == image i of a 7-image launch, bit for bit."""
    from difashion_b200 import ops
    g = torch.Generator().manual_seed(3)
    for hw_side, c in ((4, 128), (2, 128), (16, 64), (24, 320)):
        x = torch.randn(48, hw_side, hw_side, c, generator=g).cuda()
        gamma, beta = torch.randn(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
        outs = []
        for b in (48, 7):
            ws = torch.empty(ops.groupnorm_ws_floats(b, 32), dtype=torch.float32, device="cuda")
            o = torch.empty(b, hw_side, hw_side, c, dtype=torch.bfloat16, device="cuda")
            ops.groupnorm(x[:b].contiguous(), None, gamma, beta, groups=32, eps=1e-5, silu=True, stats_ws=ws, out=o)
            outs.append(o)
        torch.cuda.synchronize()
        assert torch.equal(outs[0][:7], outs[1]), (hw_side, c)
