"""One full-size denoising step (16 outfits = 256 UNet rows) between cudaProfilerStart/Stop, for

  ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --csv --log-file gpurun_out/step_metrics.csv python tools/step_traffic.py

and, with ``--summarise <csv> <out.json>``, the per-kernel summary of that launch list (time share and DRAM
traffic per kernel family; the ``traffic`` figure bench.py reports for the tcgen05 GEMM/conv kernel)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from bench import synthetic_inputs
    from difashion_b200.mutual import MutualEncoder
    from difashion_b200.pipeline import B200DiFashionPipeline
    from difashion_b200.schedulers import B200DDIMScheduler
    from difashion_b200.unet import B200UNet2DConditionModel
    outfits = int(os.environ.get("OUTFITS", "16"))
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    unet, me = B200UNet2DConditionModel(), MutualEncoder()
    unet.pack(dev)
    pipe = B200DiFashionPipeline(unet, me, B200DDIMScheduler(), eta_mutual=0.1, use_cuda_graph=False)
    st = pipe.begin(**synthetic_inputs(outfits), num_inference_steps=50, device=dev)
    pipe.step(st, st.timesteps[0])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pipe.step(st, st.timesteps[1])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("launches in the profiled step:", pipe.last_step_launches)


def family(name: str) -> str:
    for key, fam in (("gemm_tcgen05", "gemm_tcgen05_kernel"), ("attn_fwd_sa8", "attn_fwd_sa8_kernel (S=4096 self-attention)"),
                     ("attn_fwd_sa", "attn_fwd_sa_kernel (other self-attention + redo pass)"), ("attn_short_kv", "attn_short_kv_kernel (cross-attention)"),
                     ("attn_fwd_db", "attn_fwd_db_kernel"), ("attn_fwd", "attn_fwd_kernel"), ("cast_rows", "cast"),
                     ("groupnorm_apply", "groupnorm_apply"), ("groupnorm_finalize", "groupnorm_finalize"), ("layernorm", "layernorm"),
                     ("cfg_step", "cfg_step"), ("mutual", "mutual_*"), ("upsample", "upsample2x"), ("space_to_depth", "space_to_depth")):
        if key in name:
            return fam
    return "other"


def summarise(path: str, out: str):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ci = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        key = (r[ci["ID"]], r[ci["Kernel Name"]])
        d = per.setdefault(key, {})
        val = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(unit, 1.0)
        d[r[ci["Metric Name"]]] = val * scale
    fam = {}
    for (_, name), d in per.items():
        f = fam.setdefault(family(name), dict(launches=0, time_s=0.0, dram_read=0.0, dram_write=0.0))
        f["launches"] += 1
        f["time_s"] += d.get("gpu__time_duration.sum", 0.0)
        f["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        f["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(f["time_s"] for f in fam.values())
    for f in fam.values():
        f["share_of_step"] = f["time_s"] / total if total else None
        f["dram_bytes_per_launch"] = (f["dram_read"] + f["dram_write"]) / max(f["launches"], 1)
        f["dram_gbs"] = (f["dram_read"] + f["dram_write"]) / f["time_s"] / 1e9 if f["time_s"] else None
    res = dict(source=os.path.basename(path), note="ncu per-launch times are serialised and cold-cache: compare shares, not absolutes",
               step_time_s=total, families=fam)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["time_s"]):
        print(f"{k:24s} launches={f['launches']:4d} time={f['time_s'] * 1e3:8.3f} ms share={f['share_of_step']:.3f} "
              f"dram={(f['dram_read'] + f['dram_write']) / 1e9:8.2f} GB ({f['dram_gbs']:.0f} GB/s)")


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2], sys.argv[3])
    else:
        run()
