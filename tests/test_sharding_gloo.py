"""world_size-2 CPU (gloo) test of the multi-GPU host logic: whole-outfit sharding + final-latent gather.
The denoising loop itself needs no collective; only the gather of finished latents does."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_outfits, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from difashion_b200.pipeline import mutual_index_table, shard_outfits
    mine = shard_outfits(n_outfits, rank, world)
    # every rank fabricates the "finished latents" of its own outfits: value = global item id
    olists = torch.zeros(n_outfits, 4, dtype=torch.long)
    local = olists[mine.start:mine.stop]
    tab = mutual_index_table(local)                      # indices are LOCAL to the rank's shard: no cross-rank refs
    n_local = len(mine) * 4
    assert tab.shape == (n_local, 3) and int(tab.min()) >= -n_local and int(tab.max()) < 0
    lat = torch.stack([torch.full((4, 2, 2), float(o * 4 + i)) for o in mine for i in range(4)]) if n_local else torch.zeros(0, 4, 2, 2)
    counts = [len(shard_outfits(n_outfits, r, world)) * 4 for r in range(world)]
    pad = max(counts)
    buf = torch.zeros(pad, 4, 2, 2)
    buf[:n_local] = lat
    out = [torch.zeros(pad, 4, 2, 2) for _ in range(world)]
    dist.all_gather(out, buf)
    full = torch.cat([o[:c] for o, c in zip(out, counts)])
    q.put((rank, list(mine), full[:, 0, 0, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_outfit_sharding_and_gather_world2():
    for n_outfits in (5, 16):
        port = _free_port()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_outfits, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        shards = sorted(res)
        covered = shards[0][1] + shards[1][1]
        assert covered == list(range(n_outfits))                       # every outfit exactly once, never split
        for _, _, gathered in shards:
            assert gathered == [float(i) for i in range(n_outfits * 4)]  # gather restores global item order


def _worker_api(rank, world, port, q):
    """shard_generation_inputs + gather_item_rows (the host side of B200DiFashionPipeline.generate_sharded) on FITB-style
    outfits with uneven blank counts; the 'generation' is a stand-in that tags every item row with its global item id."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from difashion_b200.pipeline import gather_item_rows, mutual_index_table, shard_generation_inputs
    res = []
    for olists in (torch.tensor([[3, 0, 7, 9], [0, 5, 0, 2], [4, 4, 4, 0], [0, 0, 0, 0], [0, 1, 1, 1]]),
                   torch.tensor([[0, 0, 0, 0]]),                                   # fewer outfits than ranks: rank 1 idles
                   torch.zeros(6, 4, dtype=torch.long)):
        bsz, olen = olists.shape
        n = int((olists == 0).sum())
        glob = dict(olists=olists, all_latents=torch.arange(bsz * olen).float().view(-1, 1, 1, 1).expand(-1, 4, 2, 2),
                    category_prompts=torch.arange(n).float().view(-1, 1, 1).expand(-1, 3, 5), null_prompt=torch.zeros(1, 3, 5),
                    hist_latents=torch.arange(n).float().view(-1, 1, 1, 1).expand(-1, 4, 2, 2), null_latent=torch.zeros(4, 2, 2),
                    init_latents=torch.arange(n).float().view(-1, 1, 1, 1).expand(-1, 4, 2, 2))
        local, (i0, i1), counts = shard_generation_inputs(glob, rank, world)
        assert sum(counts) == n and counts[rank] == i1 - i0 == int((local["olists"] == 0).sum())
        assert local["all_latents"].shape[0] == local["olists"].shape[0] * olen
        for k in ("category_prompts", "hist_latents", "init_latents"):
            assert local[k].shape[0] == i1 - i0 and (i1 == i0 or float(local[k].flatten()[0]) == float(i0))
        if i1 > i0:
            tab = mutual_index_table(local["olists"])                   # indices stay inside the rank's own shard
            assert int(tab.max()) < local["all_latents"].shape[0] and int(tab.min()) >= -(i1 - i0)
        lat = local["init_latents"] * 2.0 + 1.0 if i1 > i0 else None      # stand-in for generate()
        full = gather_item_rows(lat, counts, (4, 2, 2), torch.device("cpu"))
        res.append(full[:, 0, 0, 0].tolist())
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_generate_sharded_host_logic_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_api, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for n, a, b in zip((9, 4, 24), res[0], res[1]):
        assert a == b == [2.0 * i + 1.0 for i in range(n)]              # global item order on every rank


def test_shard_outfits_properties():
    from difashion_b200.pipeline import shard_outfits
    for n in (0, 1, 7, 8, 9, 128, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_outfits(n, r, w) for r in range(w)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
