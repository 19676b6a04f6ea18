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


def test_shard_outfits_properties():
    from difashion_b200.pipeline import shard_outfits
    for n in (0, 1, 7, 8, 9, 128, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_outfits(n, r, w) for r in range(w)]
            assert sum(len(p) for p in parts) == n
            assert [i for p in parts for i in p] == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
