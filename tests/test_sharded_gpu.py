"""Two ranks of ``B200DiFashionPipeline.generate_sharded`` against the single-process ``generate`` on the same job: the
sharded latents must equal the single-rank latents BIT FOR BIT (whole outfits are independent; every kernel is
batch-invariant), in global item order, on every rank — uneven shards included.  Both ranks share the one GPU of the test box
and gather over gloo (NCCL refuses two ranks on one device); on a multi-GPU box set DFB_TEST_NCCL=1 to use one GPU per rank
and the NCCL all-gather the product uses there."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _job(olists, seed=5):
    from tests.test_unet_gpu import _gen_inputs
    from oracle.unet_oracle import tiny_config
    return _gen_inputs(tiny_config(), olists, seed=seed)


def _pipe(sched_name):
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler, B200PNDMScheduler
    from tests.test_unet_gpu import _mk
    oracle, unet = _mk("tiny")
    torch.manual_seed(11)                                   # identical MutualEncoder on every rank
    me = MutualEncoder(latent_size=oracle.cfg.sample_size, hid_dim=64).cuda()
    return B200DiFashionPipeline(unet, me, B200DDIMScheduler() if sched_name == "ddim" else B200PNDMScheduler())


OLISTS = (torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0], [0, 0, 0, 0], [0, 1, 1, 1]]),      # 5 outfits, 9 blanks: shards of 3 + 2 outfits
          torch.zeros(3, 4, dtype=torch.long))                                                       # GOR, 2 + 1 outfits


def _worker(rank, world, port, sched_name, nccl, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    pipe = _pipe(sched_name)
    outs = []
    for olists in OLISTS:
        lat = pipe.generate_sharded(**_job(olists), num_inference_steps=50, max_steps=4, device=dev)
        outs.append(lat.cpu().numpy())          # by value: tensors would travel as shared-memory handles the exiting worker unlinks
    q.put((rank, outs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sched_name", ["ddim", "pndm"])
def test_two_rank_generate_sharded_equals_single_rank_bitwise(sched_name):
    nccl = os.environ.get("DFB_TEST_NCCL") == "1" and torch.cuda.device_count() >= 2
    pipe = _pipe(sched_name)
    want = [pipe.generate(**_job(o), num_inference_steps=50, max_steps=4, device="cuda").clone().cpu() for o in OLISTS]
    del pipe
    torch.cuda.synchronize()
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sched_name, nccl, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        for got, ref in zip(res[r], want):
            assert got.shape == tuple(ref.shape) and torch.equal(torch.from_numpy(got), ref), f"rank {r}: sharded latents differ from the single-rank generation"
