"""``B200CLIPTextModel`` — transformers' ``CLIPTextModel`` (the SD-1.5 text encoder) on the hand-written sm_100a
kernels: the stage right before the denoising loop,

    category_prompts = self.text_encoder(fill_input_ids)[0]        (DiFashion/models/difashion.py:339-341)
    null_prompt      = self.text_encoder(null_input_ids)[0]        (:343-352; training: :224)

Same conventions as ``unet.py`` / ``vae.py``: the transformers module tree and state-dict keys
(``text_model.embeddings.token_embedding.weight`` ...), plain ``nn`` modules as parameter containers, an fp32
residual stream with bf16 tensor-core operands (or fp32 operands on the verification path), no PyTorch arithmetic,
no CPU fallback.  Per layer: LayerNorm -> fused q/k/v projection (one GEMM, biases in the epilogue) -> causal flash
attention (``dfb_attention`` with ``causal = 1``: 77 keys are one tile) -> out projection + residual -> LayerNorm ->
fc1 + quick-GELU epilogue -> fc2 + residual; token + position embedding is one gather kernel.

Only ~50 distinct category prompts + the empty prompt exist (``data_utils.py:102-106``), so ``encode_table`` turns the
whole vocabulary of prompts into a ``[P, 77, 768]`` table once; the loop then indexes it.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .config import FrozenConfig
from .unet import Workspace, _f32

SD15_TEXT_ENCODER_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                                layer_norm_eps=1e-5, bos_token_id=49406, eos_token_id=49407, pad_token_id=49407)


@dataclass
class CLIPTextOutput:
    """``BaseModelOutputWithPooling`` subset: ``out[0]`` / ``out.last_hidden_state`` is what the reference reads."""
    last_hidden_state: torch.Tensor
    pooler_output: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.token_embedding = nn.Embedding(cfg["vocab_size"], cfg["hidden_size"])
        self.position_embedding = nn.Embedding(cfg["max_position_embeddings"], cfg["hidden_size"])


class _SelfAttn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)


class _MLP(nn.Module):
    def __init__(self, d, inner):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, inner), nn.Linear(inner, d)


class _Layer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["hidden_size"]
        self.self_attn = _SelfAttn(d)
        self.layer_norm1 = nn.LayerNorm(d, eps=cfg["layer_norm_eps"])
        self.mlp = _MLP(d, cfg["intermediate_size"])
        self.layer_norm2 = nn.LayerNorm(d, eps=cfg["layer_norm_eps"])


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(cfg) for _ in range(cfg["num_hidden_layers"])])


class _TextTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self.final_layer_norm = nn.LayerNorm(cfg["hidden_size"], eps=cfg["layer_norm_eps"])


class B200CLIPTextModel(nn.Module):
    def __init__(self, **config):
        super().__init__()
        cfg = dict(SD15_TEXT_ENCODER_CONFIG)
        cfg.update(config)
        if cfg["hidden_act"] not in ("quick_gelu", "gelu"):
            raise NotImplementedError("hidden_act must be 'quick_gelu' (SD-1.5's text encoder) or 'gelu' (SD-2's)")
        if cfg["hidden_size"] % cfg["num_attention_heads"] or (cfg["hidden_size"] // cfg["num_attention_heads"]) % 16:
            raise NotImplementedError("head dim must be a multiple of 16")
        self._config = FrozenConfig(cfg)
        self.text_model = _TextTransformer(cfg)
        for p in self.parameters():
            p.requires_grad_(False)
        self._op_dtype = torch.bfloat16
        self._pack: Optional[Dict[str, Any]] = None
        self._pack_key = None
        self._ws: Dict[Any, Workspace] = {}

    @property
    def config(self) -> FrozenConfig:
        return self._config

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def set_precision(self, precision: str):
        """``"bf16"`` (tensor cores) or ``"fp32"`` (verification path on the CUDA cores); see the UNet."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._op_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        return self

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw):
        """``CLIPTextModel.from_pretrained(dir, subfolder="text_encoder")`` (difashion.py:70-72): transformers layout,
        ``config.json`` + ``model.safetensors`` or ``pytorch_model.bin``."""
        from . import checkpoint as ck
        d = ck.model_dir(path, subfolder)
        cfg = ck.read_config(d)
        m = cls(**{k: v for k, v in cfg.items() if k in SD15_TEXT_ENCODER_CONFIG})
        m.load_transformers_state_dict(ck.read_state_dict(d, ck.TRANSFORMERS_STEMS))
        return m

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **kw):
        from . import checkpoint as ck
        ck.write_config(save_directory, dict(self._config), "CLIPTextModel")
        ck.write_state_dict(save_directory, self.state_dict(), "model" if safe_serialization else "pytorch_model",
                            safe_serialization)

    def load_transformers_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Load a ``CLIPTextModel`` state dict (the ``position_ids`` buffer of older checkpoints is skipped)."""
        return self.load_state_dict({k: v for k, v in sd.items() if not k.endswith("position_ids")}, strict=True)

    # ---------------------------------------------------------------- packing
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def pack(self, device=None):
        device = torch.device(device) if device is not None else self.device
        if device.type != "cuda":
            raise RuntimeError("B200CLIPTextModel needs a CUDA device: there is no CPU fallback")
        dt = self._op_dtype
        key = (str(device), str(dt), self._weights_key())
        if self._pack is not None and self._pack_key == key:
            return self._pack
        tm = self.text_model
        P: Dict[str, Any] = dict(tok=_f32(tm.embeddings.token_embedding.weight, device),
                                 pos=_f32(tm.embeddings.position_embedding.weight, device), layers=[])
        for ly in tm.encoder.layers:
            a = ly.self_attn
            wqkv = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], dim=0).detach().to(device)
            bqkv = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], dim=0)
            P["layers"].append(dict(
                ln1=(_f32(ly.layer_norm1.weight, device), _f32(ly.layer_norm1.bias, device), ly.layer_norm1.eps),
                ln2=(_f32(ly.layer_norm2.weight, device), _f32(ly.layer_norm2.bias, device), ly.layer_norm2.eps),
                wqkv=ops.pack_linear(wqkv, dt), bqkv=_f32(bqkv, device),
                wo=ops.pack_linear(a.out_proj.weight.to(device), dt), bo=_f32(a.out_proj.bias, device),
                w1=ops.pack_linear(ly.mlp.fc1.weight.to(device), dt), b1=_f32(ly.mlp.fc1.bias, device),
                w2=ops.pack_linear(ly.mlp.fc2.weight.to(device), dt), b2=_f32(ly.mlp.fc2.bias, device)))
        fl = tm.final_layer_norm
        P["final"] = (_f32(fl.weight, device), _f32(fl.bias, device), fl.eps)
        self._pack, self._pack_key = P, key
        return P

    # ---------------------------------------------------------------- kernel sequencing
    def _encode_chunk(self, ids: torch.Tensor, out: torch.Tensor, ws: Workspace):
        """ids int32 [B, S] (device) -> out fp32 [B, S, D]."""
        P = self.pack(ids.device)
        cfg, dt = self.config, self._op_dtype
        B, S = ids.shape
        D, heads, inner = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        hd, M = D // heads, B * S
        h = ws.get("h", (M, D), torch.float32)
        ops.embed_tokens(ids, P["tok"], P["pos"], h)
        xn = ws.get("xn", (M, D), dt)
        qkv = ws.get("qkv", (B, S, 3 * D), dt)
        att = ws.get("att", (B, S, D), dt)
        mid = ws.get("mid", (M, inner), dt)
        for L in P["layers"]:
            ops.layernorm(h, L["ln1"][0], L["ln1"][1], xn, eps=L["ln1"][2])
            ops.gemm([xn], L["wqkv"], 3 * D, out=qkv.view(M, 3 * D), bias=L["bqkv"])
            ops.attention(qkv, qkv, qkv, att, heads=heads, dp=hd, scale=float(hd) ** -0.5, q_col0=0, k_col0=D, v_col0=2 * D,
                          causal=True)
            ops.gemm([att.view(M, D)], L["wo"], D, out=h, bias=L["bo"], residual=h)
            ops.layernorm(h, L["ln2"][0], L["ln2"][1], xn, eps=L["ln2"][2])
            ops.gemm([xn], L["w1"], inner, out=mid, bias=L["b1"], act=ops.ACT_GELU if cfg.hidden_act == "gelu" else ops.ACT_QUICK_GELU)
            ops.gemm([mid], L["w2"], D, out=h, bias=L["b2"], residual=h)
        ops.layernorm(h, P["final"][0], P["final"][1], out.view(M, D), eps=P["final"][2])
        return out

    def workspace(self, key, device) -> Workspace:
        ws = self._ws.get(key)
        if ws is None or ws.device != device:
            ws = Workspace(device)
            self._ws[key] = ws
        return ws

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask=None, position_ids=None, output_attentions=None,
                output_hidden_states=None, return_dict: Optional[bool] = None, max_batch: int = 256):
        """``CLIPTextModel.forward``: input_ids [B, S<=max_position_embeddings] (int64/int32) ->
        ``CLIPTextOutput`` (``out[0]`` = last_hidden_state fp32 [B, S, D]); ``(last_hidden_state,)`` when
        ``return_dict=False``.  The reference passes ids only; masks / custom positions are not implemented."""
        if attention_mask is not None or position_ids is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("B200CLIPTextModel implements the reference's call: text_encoder(input_ids)")
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [batch, sequence]")
        cfg = self.config
        B, S = input_ids.shape
        if S > cfg.max_position_embeddings:
            raise ValueError(f"sequence length {S} exceeds max_position_embeddings {cfg.max_position_embeddings}")
        if input_ids.device.type != "cuda" and self.device.type != "cuda":
            raise RuntimeError("B200CLIPTextModel needs a CUDA device: there is no CPU fallback")
        dev = input_ids.device if input_ids.is_cuda else self.device
        if bool(((input_ids < 0) | (input_ids >= cfg.vocab_size)).any()):
            raise IndexError("input_ids out of the vocabulary")
        ids = input_ids.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(B, S, cfg.hidden_size, dtype=torch.float32, device=dev)
        ws = self.workspace(("enc", min(B, max_batch), S, str(self._op_dtype)), dev)
        for b0 in range(0, B, max_batch):
            self._encode_chunk(ids[b0:b0 + max_batch], out[b0:b0 + max_batch], ws)
        if return_dict is False:
            return (out,)
        return CLIPTextOutput(last_hidden_state=out)

    def null_input_ids(self, max_length: Optional[int] = None, pad_token_id: Optional[int] = None) -> torch.Tensor:
        """``tokenizer([""], padding="max_length", max_length=...)`` of the CLIPTokenizer (difashion.py:343-350): BOS, EOS, then
        padding.  The ids are the CLIP vocabulary's last two entries (``<|startoftext|>`` = vocab_size - 2 = 49406,
        ``<|endoftext|>`` = vocab_size - 1 = 49407), NOT ``config.bos_token_id / eos_token_id``: the ``text_encoder/config.json``
        Stable Diffusion ships carries transformers' generic defaults (bos 0, eos 2, pad 1), which are not CLIP tokens.
        ``pad_token_id``: SD-1.5's tokenizer pads with EOS (the default here); SD-2's pads with id 0 — pass it, or give
        ``B200DiFashion`` the tokenizer directory, which is what the reference does."""
        cfg = self.config
        n = max_length or cfg.max_position_embeddings
        bos, eos = cfg.vocab_size - 2, cfg.vocab_size - 1
        ids = torch.full((1, n), eos if pad_token_id is None else int(pad_token_id), dtype=torch.long)
        ids[0, 0] = bos
        if n > 1:
            ids[0, 1] = eos
        return ids

    @torch.no_grad()
    def encode_table(self, prompt_ids: torch.Tensor) -> torch.Tensor:
        """All distinct prompts at once: ids [P, S] -> [P, S, D] (the reference re-encodes the same <=51 prompts for
        every batch, difashion.py:339-353; the table is step- and batch-invariant)."""
        return self.forward(prompt_ids)[0]
