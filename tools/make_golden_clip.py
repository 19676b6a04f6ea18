"""Generate tests/golden/clip_tiny.pt by importing the REAL ``transformers.CLIPTextModel`` (the third-party module the
reference's text stage lives in, ``DiFashion/models/difashion.py:71-73``, ``:339-352``): a small random-init config, its
state dict, input ids (incl. the empty prompt) and the model's ``last_hidden_state``.  This pins oracle/clip_oracle.py —
and through it ``B200CLIPTextModel`` — to the reference's own implementation.   Usage: python tools/make_golden_clip.py
(transformers 5.5.0 in this image; the reference pins 4.32.1 — same CLIP text arithmetic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# clip_tiny.pt: SD-1.5's activation (quick_gelu); clip_tiny_gelu.pt: SD-2-base's (exact erf "gelu", head dim 64 as in
# OpenCLIP ViT-H's text tower) — the reference's default base model (train.py:44)
VARIANTS = {
    "clip_tiny.pt": dict(vocab_size=300, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         max_position_embeddings=77, layer_norm_eps=1e-5),
    "clip_tiny_gelu.pt": dict(vocab_size=200, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                              num_attention_heads=2, max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act="gelu"),
}


def main():
    for name, cfg in VARIANTS.items():
        make(os.path.join(GOLD, name), cfg)


def make(OUT, CFG):
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(0)
    hf = CLIPTextModel(CLIPTextConfig(**dict(dict(hidden_act="quick_gelu", projection_dim=64, pad_token_id=1, bos_token_id=0,
                                                  eos_token_id=2), **CFG))).eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in hf.parameters():                       # non-trivial norms / biases (HF inits biases to zero)
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.mul_(3.0)
    ids = torch.randint(0, CFG["vocab_size"], (4, 77), generator=g)
    ids[:, 0] = CFG["vocab_size"] - 2                   # BOS-like
    ids[1, 6:] = CFG["vocab_size"] - 1                  # short prompt padded with an EOS-like id
    ids[3, 1:] = CFG["vocab_size"] - 1                  # the empty prompt's shape: BOS, EOS, EOS, ...
    with torch.no_grad():
        y = hf(ids)[0]
    sd = {k: v.clone() for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    torch.save(dict(config=CFG, state_dict=sd, input_ids=ids, last_hidden_state=y,
                    transformers_version=transformers.__version__), OUT)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e3:.0f} kB, output std {float(y.std()):.3f}")


if __name__ == "__main__":
    main()
