"""Diagnostic: PLMS / DDIM generation with the batch split into row chunks vs unsplit (max abs difference per item)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_unet_gpu import _gen_inputs, _mk
from difashion_b200.mutual import MutualEncoder
from difashion_b200.pipeline import B200DiFashionPipeline
from difashion_b200.schedulers import B200DDIMScheduler, B200PNDMScheduler

oracle, unet = _mk("tiny")
cfg = oracle.cfg
me = MutualEncoder(latent_size=cfg.sample_size, hid_dim=64).cuda()
inp = _gen_inputs(cfg, torch.zeros(3, 4, dtype=torch.long))
for name, cls in (("ddim", B200DDIMScheduler), ("pndm", B200PNDMScheduler)):
    for steps in (1, 2, 3, 7):
        outs = []
        for max_rows, graph in ((256, True), (256, True), (256, False), (16, True), (16, False), (20, True)):
            pipe = B200DiFashionPipeline(unet, me, cls(), max_rows=max_rows, use_cuda_graph=graph)
            outs.append(pipe.generate(**inp, num_inference_steps=50, max_steps=steps, device="cuda").clone())
        d = [float((o - outs[0]).abs().max()) for o in outs]
        per_item = (outs[3] - outs[0]).abs().amax(dim=(1, 2, 3)).tolist()
        print(name, "steps", steps, "max|diff| vs run0 [same, eager, 16, 16 eager, 20]:", ["%.2e" % v for v in d[1:]],
              "per item (16):", ["%.1e" % v for v in per_item], flush=True)
