"""``MutualEncoder`` of DiFashion (``DiFashion/models/difashion.py:21-46``) on the tcgen05 GEMM kernel.

``tanh(W2 · leaky_relu(W1 · x + b1, 0.01) + b2)`` (Dropout(0.1) is inactive in eval / inference).  Same
attribute and state-dict names as the reference (``category_embedding`` — unused there too — and
``mlp.0`` / ``mlp.3``), xavier-normal init as ``difashion.py:731-746``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .config import FrozenConfig


class MutualEncoder(nn.Module):
    def __init__(self, cate_num: int = 50, cate_emb_size: int = 64, latent_channels: int = 4, latent_size: int = 64,
                 hid_dim: int = 256):
        super().__init__()
        # @register_to_config of the reference ctor (difashion.py:25-26): what save_pretrained writes to config.json
        self._config = FrozenConfig(dict(cate_num=cate_num, cate_emb_size=cate_emb_size, latent_channels=latent_channels,
                                         latent_size=latent_size, hid_dim=hid_dim))
        self.category_embedding = nn.Embedding(cate_num, cate_emb_size)  # useless embedding (difashion.py:28)
        self.latent_channels, self.latent_size, self.hid_dim = latent_channels, latent_size, hid_dim
        d = latent_channels * latent_size * latent_size
        self.mlp = nn.Sequential(nn.Linear(d, hid_dim), nn.LeakyReLU(), nn.Dropout(0.1), nn.Linear(hid_dim, d),
                                 nn.Tanh())
        for m in self.modules():                      # xavier_normal_initialization, difashion.py:731-746
            if isinstance(m, nn.Embedding):
                nn.init.xavier_normal_(m.weight.data)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    nn.init.constant_(m.bias.data, 0)
        for p in self.parameters():
            p.requires_grad_(False)
        self._pk = None

    @property
    def d(self) -> int:
        return self.latent_channels * self.latent_size * self.latent_size

    @property
    def config(self) -> FrozenConfig:
        return self._config

    def register_to_config(self, **kw):
        """``model.fashion_encoder.register_to_config(**load_model.config)`` (inf4eval.py:579)."""
        d = dict(self._config)
        d.update({k: v for k, v in kw.items() if not k.startswith("_")})
        self._config = FrozenConfig(d)

    def save_pretrained(self, save_directory: str, safe_serialization: bool = False, **kw):
        """diffusers ``ModelMixin.save_pretrained`` layout, as ``save_model_hook`` writes ``<ckpt>/fashion_encoder``
        (inf4eval.py:550)."""
        from . import checkpoint as ck
        ck.write_config(save_directory, dict(self._config), "MutualEncoder")
        ck.write_state_dict(save_directory, self.state_dict(), ck.DIFFUSERS_STEM, safe_serialization)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw):
        """``MutualEncoder.from_pretrained(input_dir, subfolder="fashion_encoder")`` (inf4eval.py:578)."""
        from . import checkpoint as ck
        d = ck.model_dir(path, subfolder)
        cfg = ck.read_config(d)
        m = cls(**{k: cfg[k] for k in ("cate_num", "cate_emb_size", "latent_channels", "latent_size", "hid_dim") if k in cfg})
        m.load_state_dict(ck.read_state_dict(d), strict=True)
        return m

    def pack(self, device, dtype: torch.dtype = torch.bfloat16):
        key = (str(device), str(dtype), tuple((p.data_ptr(), p._version) for p in self.mlp.parameters()))
        if self._pk is None or self._pk["key"] != key:
            l1, l2 = self.mlp[0], self.mlp[3]
            f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
            self._pk = dict(key=key, w1=ops.pack_linear(l1.weight.to(device), dtype), b1=f(l1.bias),
                            w2=ops.pack_linear(l2.weight.to(device), dtype), b2=f(l2.bias))
        return self._pk

    def encode_bf16(self, x_bf16: torch.Tensor, out: torch.Tensor, hid: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: bf16 (or fp32: verification path) [N, d] (the neighbour sum) -> out fp32 [N, d] in (-1, 1)."""
        pk = self.pack(x_bf16.device, x_bf16.dtype)
        n = x_bf16.shape[0]
        if hid is None:
            hid = torch.empty(n, self.hid_dim, dtype=x_bf16.dtype, device=x_bf16.device)
        ops.gemm([x_bf16], pk["w1"], self.hid_dim, out=hid, bias=pk["b1"], act=ops.ACT_LEAKY_RELU)
        ops.gemm([hid], pk["w2"], self.d, out=out, bias=pk["b2"], act=ops.ACT_TANH)
        return out

    @torch.no_grad()
    def forward(self, mutual_emb: torch.Tensor) -> torch.Tensor:
        if not mutual_emb.is_cuda:
            raise RuntimeError("MutualEncoder (B200) needs CUDA tensors: there is no CPU fallback")
        bsz = mutual_emb.shape[0]
        x = mutual_emb.reshape(bsz, 1, -1).contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        xb = torch.empty(bsz, 1, self.d, dtype=torch.bfloat16, device=x.device)
        ops.pad_cast_rows(x, xb)
        out = torch.empty(bsz, self.d, dtype=torch.float32, device=x.device)
        self.encode_bf16(xb.view(bsz, self.d), out)
        return out.view(bsz, self.latent_channels, self.latent_size, self.latent_size).to(mutual_emb.dtype)
