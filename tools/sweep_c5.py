"""BASELINE configs[4] — throughput sweep: ONE job of O outfits (1..512) x DDIM steps {20, 50}, sharded over the N GPUs of the
box (strong scaling), all points inside one process group so the start-up cost is paid once.  Per point: `value` =
device-resident outfits/s (K timed steps, CUDA events, max over ranks) and, for jobs of <= E2E_MAX outfits, `e2e` through
B200DiFashionPipeline.generate_sharded with pinned host inputs (H2D, 50 or 20 steps, NCCL gather, D2H).  Also reports kernel
launches per step and, with --whole-step-graph, what capturing the entire step (mutual + all chunks + CFG/DDIM) in one graph
buys at the small-batch end (the reference's own batches are 4 GOR / 15 FITB outfits, inf4eval.py:521-524).
Usage (N ranks):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
                  tools/sweep_c5.py --out gpurun_out/c5_sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outfits", default="1,4,15,16,64,128,512")
    ap.add_argument("--ddim-steps", default="20,50")
    ap.add_argument("--timed-steps", type=int, default=10)
    ap.add_argument("--e2e-max", type=int, default=128, help="largest job that also gets an end-to-end generation")
    ap.add_argument("--skv", type=int, default=77)
    ap.add_argument("--task", default="GOR")
    ap.add_argument("--out", default="gpurun_out/c5_sweep.jsonl")
    a = ap.parse_args()
    import torch.distributed as dist
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from difashion_b200 import ops
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline, shard_generation_inputs
    from difashion_b200.schedulers import B200DDIMScheduler
    from difashion_b200.unet import B200UNet2DConditionModel
    torch.manual_seed(0)
    unet, me = B200UNet2DConditionModel(), MutualEncoder()
    unet.pack(dev)
    pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), eta_mutual=0.1, max_rows=256)
    rows_per_outfit = 16 if a.task == "GOR" else 4
    lines = []
    for total in [int(v) for v in a.outfits.split(",")]:
        glob = bench.synthetic_inputs(total, a.skv, seed=123, task=a.task)
        inp, (i0, i1), counts = shard_generation_inputs(glob, rank, world)
        mine = inp["olists"].shape[0]
        for steps in [int(v) for v in a.ddim_steps.split(",")]:
            st, t_ms = None, 0.0
            if mine:
                st = pipe.begin(**inp, num_inference_steps=steps, device=dev)
                for w in range(3):
                    pipe.step(st, st.timesteps[w])
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = min(a.timed_steps, steps - 3)
            e0.record()
            for k in range(K if mine else 0):
                pipe.step(st, st.timesteps[3 + k])
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
            line = dict(config="C5", task=a.task, outfits_total=total, ddim_steps=steps, n_gpus=world, unit="outfits/s",
                        value=total * (K / steps) / (t_ms / 1e3), ms_per_step=t_ms / K, timed_steps=K,
                        unet_rows_per_step_rank0=mine * rows_per_outfit, launches_per_step=pipe.last_step_launches if mine else 0,
                        idle_ranks=sum(1 for c in counts if c == 0))
            if total <= a.e2e_max:
                n_items = sum(counts)
                out_host = torch.empty(n_items, 4, 64, 64).pin_memory()

                def gen():
                    if world > 1:
                        pipe.generate_sharded(**glob, num_inference_steps=steps, device=dev, out=out_host)
                    else:
                        pipe.generate(**inp, num_inference_steps=steps, device=dev, out=out_host)
                    torch.cuda.synchronize()
                gen()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                gen()
                dt = torch.tensor([time.perf_counter() - t0], device=dev)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                line["e2e"] = dict(value=total / float(dt.item()), unit="outfits/s", seconds_per_generation=float(dt.item()),
                                   finite=bool(torch.isfinite(out_host).all()))
            lines.append(line)
            if rank == 0:
                print(json.dumps(line), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
