"""The denoising loop of ``DiFashion.fashion_generation`` on the B200 kernels.

Mirrors ``DiFashion/models/difashion.py:277-616`` for the part that is the hot path (the loop body
``:456-577`` and its static set-up ``:309-325``, ``:388-451``); VAE / CLIP stages are the caller's.
"""
from __future__ import annotations

import torch

INT32_MIN = -(2 ** 31)


def mutual_index_table(olists: torch.Tensor) -> torch.Tensor:
    """Static gather table for the mutual condition (difashion.py:439-451 + :477-487).

    Row n (n-th generated item in ``torch.nonzero(olists == 0)`` order) lists the sources of the other
    ``olen-1`` slots of its outfit: ``j >= 0`` -> row j of ``all_latents`` (a given item),
    ``j < 0`` -> row ``-j-1`` of the generated latents.  int32 ``[N, olen-1]`` (CPU)."""
    olists = olists.cpu()
    bsz, olen = olists.shape
    gen = olists == 0
    gen_row = torch.full((bsz, olen), -1, dtype=torch.int64)
    gen_row[gen] = torch.arange(int(gen.sum()))
    rows = []
    for o in range(bsz):
        for i in range(olen):
            if not gen[o, i]:
                continue
            row = []
            for s in range(olen):
                if s == i:
                    continue
                row.append(-int(gen_row[o, s]) - 1 if gen[o, s] else o * olen + s)
            rows.append(row)
    return torch.tensor(rows, dtype=torch.int32).reshape(len(rows), olen - 1)


# --------------------------------------------------------------------------------------------------
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

from . import ops
from .mutual import MutualEncoder
from .schedulers import B200DDIMScheduler, B200PNDMScheduler
from .unet import B200UNet2DConditionModel, Workspace


def guidance_plan(use_history: bool, use_mutual: bool, s_cate: float, s_hist: float, s_mutual: float):
    """CFG branch layout of ``fashion_generation`` (difashion.py:309-325, :388-431, :494-512, :525-566).

    Returns (ctx_is_cat[b], use_m[b], use_h[b], weights[b]) per branch b, so that
    ``eps = sum_b weights[b] * eps_b`` equals the reference's nested guidance formula."""
    do_h = bool(use_history and s_hist > 1.0)
    do_m = bool(use_mutual and s_mutual > 1.0)
    do_c = bool(s_cate > 1.0)
    if do_h and do_m and do_c:
        return [1, 1, 1, 0], [1, 1, 0, 0], [1, 0, 0, 0], [s_hist, s_mutual - s_hist, s_cate - s_mutual, 1.0 - s_cate]
    if do_c:
        if do_h:
            return [1, 1, 0], [1, 1, 1], [1, 0, 0], [s_hist, s_cate - s_hist, 1.0 - s_cate]
        if do_m:
            return [1, 1, 0], [1, 0, 0], [1, 1, 1], [s_mutual, s_cate - s_mutual, 1.0 - s_cate]
        return [1, 0], [1, 1], [1, 1], [s_cate, 1.0 - s_cate]
    if do_h:
        # (the mutual section checks do_m first, difashion.py:506-508: with both history and mutual guidance on and no
        # category guidance the second branch drops the mutual condition too, and the hist scale combines them, :555-560)
        return [1, 1], ([1, 0] if do_m else [1, 1]), [1, 0], [s_hist, 1.0 - s_hist]
    if do_m:
        return [1, 1], [1, 0], [1, 1], [s_mutual, 1.0 - s_mutual]
    return [1], [1], [1], [1.0]


@dataclass
class _Chunk:
    n0: int
    n1: int
    ctx: torch.Tensor                          # fp32/any [nb*(n1-n0), S, D] static buffer (refilled per generation)
    ctx_bf16: torch.Tensor
    kv_store: dict = field(default_factory=dict)
    graph: Optional[torch.cuda.CUDAGraph] = None
    eps: Optional[torch.Tensor] = None
    ws: Optional[Workspace] = None             # per-stream workspace when the chunks of a step run concurrently
    launches: int = 0


class _State:
    """Static device buffers + captured graphs for one problem shape (reused across ``generate`` calls)."""
    pass


class B200DiFashionPipeline:
    """Denoising loop of ``DiFashion.fashion_generation`` (difashion.py:456-577) on the B200 kernels.

    One step = [mutual gather-sum -> MutualEncoder MLP] for all items, then per row chunk
    [blend/concat/branch-expand -> UNet] (one CUDA graph per chunk) + one fused CFG-combine/scheduler kernel.
    Whole outfits are the unit of work: all items of an outfit must be in the same call (they couple
    through the mutual condition); different outfits never interact — which multi-GPU sharding relies on."""

    def __init__(self, unet: B200UNet2DConditionModel, mutual_encoder: Optional[MutualEncoder], scheduler,
                 eta_mutual: float = 0.1, use_history: bool = True, use_mutual_guidance: bool = True,
                 max_rows: int = 256, use_cuda_graph: bool = True, streams: int = 1,
                 share_cfg_prefix: Optional[bool] = None):
        self.unet, self.mutual_encoder, self.scheduler = unet, mutual_encoder, scheduler
        self.eta_mutual = float(eta_mutual)
        self.use_history, self.use_mutual_guidance = use_history, use_mutual_guidance
        self.max_rows = int(max_rows)
        self.use_cuda_graph = use_cuda_graph
        # streams > 1: the row chunks of one step run on that many CUDA streams inside ONE graph (fork / join), each
        # stream with its own workspace, so that the HBM-bound kernels of one chunk (GroupNorm / LayerNorm / layout
        # kernels: no shared memory, they co-reside with a persistent GEMM CTA) overlap the tensor-bound kernels of
        # another.  Needs use_cuda_graph and >= 2 chunks (max_rows < rows of the batch).
        self.streams = max(1, int(streams))
        # The last two CFG branches of every multi-branch plan that guides on the category get the SAME UNet input and
        # differ only in the prompt (branches [.., (0m,0h,c), (0m,0h,0c)], difashion.py:388-431 / :494-512): the UNet part
        # ahead of the first cross-attention is computed once for both (``forward_nhwc(shared_tail=...)``; bit-identical).
        # DFB_SHARE_PREFIX=0 (or share_cfg_prefix=False) computes every branch in full.
        if share_cfg_prefix is None:
            share_cfg_prefix = os.environ.get("DFB_SHARE_PREFIX", "1") != "0"
        self.share_cfg_prefix = bool(share_cfg_prefix)
        self._states = {}
        self.last_step_launches = 0          # kernels launched per denoising step (counted at capture / eager run)

    # ---------------------------------------------------------------------------------------------
    def _chunk_forward(self, ch: _Chunk, st) -> torch.Tensor:
        """blend -> UNet over the rows of one chunk, on the current stream."""
        n = ch.n1 - ch.n0
        x = st.latents[ch.n0:ch.n1]
        m = st.m[ch.n0:ch.n1] if st.m is not None else None
        hist = st.hist[ch.n0:ch.n1] if st.hist is not None else None
        ws = ch.ws if ch.ws is not None else st.ws
        x_in = ws.get("pipe_x_in", (st.nb * n, st.size, st.size, 8), self.unet._op_dtype)
        ops.mutual_blend(x, m, hist, st.null, self.eta_mutual, st.use_m, st.use_h, x_in)
        return self.unet.forward_nhwc(x_in, st.t_dev[: st.nb * n], ch.ctx_bf16, ch.kv_store, ws,
                                      shared_tail=n if st.shared_tail else 0)

    def _mutual(self, st):
        if st.m is None:
            return
        ops.mutual_gather_sum(st.all_latents, st.latents, st.idx, st.msum)
        self.mutual_encoder.encode_bf16(st.msum, st.m.view(st.n, -1), st.mhid)

    def _state(self, dev, n, n_given, olen, size, nb, S, D, plan) -> "_State":
        key = (str(dev), n, n_given, olen, size, nb, S, D, tuple(map(tuple, plan[:3])), self.use_history,
               self.use_mutual_guidance, self.max_rows, str(self.unet._op_dtype), self.streams, self.share_cfg_prefix,
               self.eta_mutual)                 # (eta is a kernel scalar inside the captured graph)
        st = self._states.get(key)
        if st is not None:
            return st
        st = _State()
        hw = size * size
        st.n, st.size, st.hw, st.nb = n, size, hw, nb
        st.ctx_cat, st.use_m, st.use_h = plan[0], plan[1], plan[2]
        # do the last two branches see the same (mutual, history) input?  (without a MutualEncoder / history the flag is moot)
        same = lambda flags, active: (not active) or flags[-1] == flags[-2]
        st.shared_tail = bool(self.share_cfg_prefix and nb >= 2 and same(plan[1], self.use_mutual_guidance)
                              and same(plan[2], self.use_history))
        f = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        st.latents = f(n, 4, size, size)
        st.null = f(4, size, size)
        st.hist = f(n, 4, size, size) if self.use_history else None
        st.all_latents = f(max(n_given, 1), 4, size, size)
        st.m = None
        if self.use_mutual_guidance:
            if self.mutual_encoder is None:
                raise ValueError("use_mutual_guidance=True needs a MutualEncoder")
            st.idx = torch.empty(n, max(olen - 1, 1), dtype=torch.int32, device=dev)
            st.msum = torch.empty(n, 4 * hw, dtype=self.unet._op_dtype, device=dev)
            st.mhid = torch.empty(n, self.mutual_encoder.hid_dim, dtype=self.unet._op_dtype, device=dev)
            st.m = f(n, 4, size, size)
        per = max(1, self.max_rows // nb)
        st.t_dev = torch.zeros(nb * min(n, per), dtype=torch.float32, device=dev)
        st.ws = self.unet.workspace(("pipe", nb, min(n, per), size, str(self.unet._op_dtype)), dev)
        st.chunks = []
        for n0 in range(0, n, per):
            n1 = min(n, n0 + per)
            rows = nb * (n1 - n0)
            st.chunks.append(_Chunk(n0, n1, torch.empty(rows, S, D, dtype=torch.float32, device=dev),
                                    torch.empty(rows, S, D, dtype=self.unet._op_dtype, device=dev)))
        st.multi = self.streams > 1 and self.use_cuda_graph and len(st.chunks) > 1
        if st.multi:
            st.side = [torch.cuda.Stream(device=dev) for _ in range(self.streams)]
            for i, ch in enumerate(st.chunks):
                ch.ws = self.unet.workspace(("pipe", nb, min(n, per), size, str(self.unet._op_dtype), "stream", i % self.streams), dev)
        st.step_graph = None
        st.warm = False
        st.graph_sig = None
        self._states[key] = st
        return st

    @torch.no_grad()
    def begin(self, *, olists: torch.Tensor, all_latents: Optional[torch.Tensor], category_prompts: torch.Tensor,
              null_prompt: torch.Tensor, hist_latents: Optional[torch.Tensor], null_latent: torch.Tensor,
              init_latents: torch.Tensor, num_inference_steps: int = 50, category_guidance_scale: float = 12.0,
              hist_guidance_scale: float = 4.0, mutual_guidance_scale: float = 5.0, device=None) -> "_State":
        """Stage one generation: copy the inputs (host or device) into the static device buffers, project the
        step-invariant text K/V, reset the scheduler.  Returns the state to pass to ``step``."""
        dev = torch.device(device) if device is not None else (init_latents.device if init_latents.is_cuda
                                                               else torch.device("cuda", torch.cuda.current_device()))
        if dev.type != "cuda":
            raise RuntimeError("B200DiFashionPipeline needs a CUDA device: there is no CPU fallback")
        n, size = init_latents.shape[0], init_latents.shape[-1]
        bsz, olen = olists.shape
        if int((olists == 0).sum()) != n:
            raise ValueError("init_latents must have one row per blank (olists == 0) slot")
        plan = guidance_plan(self.use_history, self.use_mutual_guidance, category_guidance_scale, hist_guidance_scale,
                             mutual_guidance_scale)
        ctx_cat, use_m, use_h, weights = plan
        nb = len(weights)
        S, D = category_prompts.shape[1], category_prompts.shape[2]
        n_given = 0 if all_latents is None else all_latents.shape[0]
        st = self._state(dev, n, n_given, olen, size, nb, S, D, plan)
        # A captured graph holds raw pointers into the UNet's packed weights: when the weights were replaced since the
        # capture (load_checkpoint between two generations — inf4eval.py's loop over checkpoints), the pack is re-made
        # here and the graphs of this state are dropped, to be captured again on the first step.
        sig = self.unet.graph_signature(dev)
        if st.graph_sig != sig:
            st.graph_sig, st.step_graph = sig, None
            for ch in st.chunks:
                ch.graph = None
        st.weights = weights
        sched = self.scheduler
        sched.set_timesteps(num_inference_steps)
        st.timesteps = [int(t) for t in sched.timesteps]
        # PLMS keeps the last four noise predictions of its sample (difashion.py:569 -> PNDMScheduler.step_plms): with more
        # than one row chunk every chunk gets its own scheduler state (same config and timesteps)
        st.chunk_scheds = None
        if not isinstance(sched, B200DDIMScheduler) and len(st.chunks) > 1:
            st.chunk_scheds = []
            for _ in st.chunks:
                c = type(sched)(**dict(sched.config))
                c.set_timesteps(num_inference_steps)
                st.chunk_scheds.append(c)

        st.latents.copy_(init_latents, non_blocking=True)
        if float(sched.init_noise_sigma) != 1.0:
            st.latents.mul_(float(sched.init_noise_sigma))
        st.null.copy_(null_latent, non_blocking=True)
        if st.hist is not None:
            if hist_latents is None:
                st.hist.copy_(st.null.unsqueeze(0).expand_as(st.hist))
            else:
                st.hist.copy_(hist_latents, non_blocking=True)
        if all_latents is not None:
            st.all_latents.copy_(all_latents, non_blocking=True)
        if st.m is not None:
            st.idx.copy_(mutual_index_table(olists), non_blocking=True)
        for ch in st.chunks:
            k = ch.n1 - ch.n0
            for b, c in enumerate(ctx_cat):
                dst = ch.ctx[b * k:(b + 1) * k]
                if c:
                    dst.copy_(category_prompts[ch.n0:ch.n1], non_blocking=True)
                else:
                    dst.copy_(null_prompt.expand(k, -1, -1), non_blocking=True)
            ch.ctx_bf16.copy_(ch.ctx)
            self.unet.project_context(ch.ctx_bf16, ch.kv_store)     # step-invariant text K/V: once per generation
        return st

    @torch.no_grad()
    def step(self, st: "_State", t: int, ddim_eta: float = 0.0, generator=None, record: Optional[list] = None):
        """One denoising step (difashion.py:456-577 loop body) for every item of the staged batch."""
        sched = self.scheduler
        latents = st.latents
        st.t_dev.fill_(float(t))
        c0 = ops.launch_count()
        self._mutual(st)
        step_launches = ops.launch_count() - c0
        eps_all = []
        if st.multi:
            # one graph for the UNet of every chunk: fork onto the side streams, join back (see __init__)
            if st.step_graph is None:
                for ch in st.chunks:
                    ch.eps = torch.empty_like(self._chunk_forward(ch, st))      # warm-up; the chunk's own eps buffer
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cur = torch.cuda.current_stream()
                    for s_ in st.side:
                        s_.wait_stream(cur)
                    for i, ch in enumerate(st.chunks):
                        c1 = ops.launch_count()
                        with torch.cuda.stream(st.side[i % self.streams]):
                            ch.eps.copy_(self._chunk_forward(ch, st))    # the result is a workspace buffer shared by the stream's chunks
                        ch.launches = ops.launch_count() - c1
                    for s_ in st.side:
                        cur.wait_stream(s_)
                st.step_graph = g
            st.step_graph.replay()
        for ci, ch in enumerate(st.chunks):
            if st.multi:
                eps = ch.eps
                step_launches += ch.launches
            elif self.use_cuda_graph and ch.graph is None:
                # warm-up (allocates workspaces, sets kernel attributes), then capture this chunk's step
                self._chunk_forward(ch, st)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                c1 = ops.launch_count()
                with torch.cuda.graph(g):
                    ch.eps = self._chunk_forward(ch, st)
                ch.graph, ch.launches = g, ops.launch_count() - c1
            if st.multi:
                pass
            elif ch.graph is not None:
                ch.graph.replay()
                eps = ch.eps
                step_launches += ch.launches
            else:
                c1 = ops.launch_count()
                eps = self._chunk_forward(ch, st)
                step_launches += ops.launch_count() - c1
            x = latents[ch.n0:ch.n1]
            if record is not None:
                eps_all.append(eps.clone())
            if isinstance(sched, B200DDIMScheduler):
                sched.cfg_step(eps, st.weights, t, x, eta=ddim_eta, generator=generator, out=x)
            else:
                (sched if st.chunk_scheds is None else st.chunk_scheds[ci]).cfg_step(eps, st.weights, t, x, out=x)
            step_launches += 1
        self.last_step_launches = step_launches
        if record is not None:
            record.append(dict(t=t, eps_branches=[e for e in eps_all], latents=latents.clone()))

    @torch.no_grad()
    def generate(self, *, num_inference_steps: int = 50, ddim_eta: float = 0.0, generator=None,
                 max_steps: Optional[int] = None, record: Optional[list] = None, out: Optional[torch.Tensor] = None,
                 **inputs) -> torch.Tensor:
        """Arguments mirror what ``fashion_generation`` has at hand when the loop starts (difashion.py:332-453):
        olists [bsz, olen] (0 = slot to generate), all_latents [bsz*olen,4,h,w] (VAE latents of the given
        items), category_prompts [N,S,D] / null_prompt [1,S,D] (CLIP hidden states), hist_latents [N,4,h,w]
        (null_latent where the user has no history), null_latent [4,h,w], init_latents [N,4,h,w], the three
        guidance scales.  Inputs may live on the host (pinned or not): they are copied into static device
        buffers.  Returns the final latents [N,4,h,w] fp32 on the device, or copies them into ``out``."""
        st = self.begin(num_inference_steps=num_inference_steps, **inputs)
        for i, t in enumerate(st.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            self.step(st, t, ddim_eta=ddim_eta, generator=generator, record=record)
        if out is not None:
            out.copy_(st.latents, non_blocking=True)
            return out
        return st.latents


    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate_sharded(self, *, olists: torch.Tensor, all_latents: Optional[torch.Tensor], category_prompts: torch.Tensor,
                         null_prompt: torch.Tensor, hist_latents: Optional[torch.Tensor], null_latent: torch.Tensor,
                         init_latents: torch.Tensor, group=None, out: Optional[torch.Tensor] = None, device=None,
                         **kw) -> torch.Tensor:
        """``generate`` for one job spread over the ranks of a ``torch.distributed`` group (one process per GPU).

        Every rank passes the SAME global inputs (host tensors are fine: only the rank's shard is copied to its GPU).  Whole
        outfits are dealt out in contiguous blocks (``shard_outfits``; the mutual condition couples only the items of one
        outfit, so the denoising loop needs no collective), every rank runs ``generate`` on its block, and ONE padded
        ``all_gather`` of the finished latents (NCCL over NVLink; uneven blocks are padded to the largest) returns the
        global ``[N, 4, h, w]`` tensor, in ``torch.nonzero(olists == 0)`` order, on every rank — bit-identical to what a
        single rank computes for the same inputs.  Ranks left without an outfit (fewer outfits than ranks) only take part
        in the gather.  Replaces the single-process loop body of ``inf4eval.py:688-760`` for a multi-GPU box."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("generate_sharded needs an initialised torch.distributed process group (torchrun); "
                               "use generate() in a single process")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        local, (i0, i1), counts = shard_generation_inputs(
            dict(olists=olists, all_latents=all_latents, category_prompts=category_prompts, null_prompt=null_prompt,
                 hist_latents=hist_latents, null_latent=null_latent, init_latents=init_latents), rank, world)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lat = self.generate(**local, device=dev, **kw) if i1 > i0 else None
        full = gather_item_rows(lat, counts, tuple(init_latents.shape[1:]), dev, group)
        if out is not None:
            out.copy_(full, non_blocking=True)
            return out
        return full


def shard_generation_inputs(inputs: dict, rank: int, world_size: int):
    """The slice of ``generate``'s inputs that rank ``rank`` of ``world_size`` owns: whole outfits ``shard_outfits`` deals it,
    the rows of ``all_latents`` of those outfits, and the item rows (blank slots, ``nonzero(olists == 0)`` order — outfit-major,
    hence contiguous per block) of the per-item tensors.  Returns ``(local_inputs, (item_start, item_stop), items_per_rank)``.
    Pure host logic (CPU-tested with gloo, world_size 2)."""
    olists = inputs["olists"].cpu()
    bsz, olen = olists.shape
    blanks = (olists == 0).sum(1)
    cum = [0]
    for b in blanks.tolist():
        cum.append(cum[-1] + int(b))
    counts = []
    for r in range(world_size):
        sh = shard_outfits(bsz, r, world_size)
        counts.append(cum[sh.stop] - cum[sh.start])
    mine = shard_outfits(bsz, rank, world_size)
    i0, i1 = cum[mine.start], cum[mine.stop]
    local = dict(inputs)
    local["olists"] = olists[mine.start:mine.stop]
    if inputs.get("all_latents") is not None:
        local["all_latents"] = inputs["all_latents"][mine.start * olen:mine.stop * olen]
    for k in ("category_prompts", "hist_latents", "init_latents"):
        if inputs.get(k) is not None:
            local[k] = inputs[k][i0:i1]
    return local, (i0, i1), counts


def gather_item_rows(local: Optional[torch.Tensor], counts: Sequence[int], row_shape, device, group=None) -> torch.Tensor:
    """One padded all-gather of per-item rows (finished latents): rank r contributes ``counts[r]`` rows; every rank gets the
    ``[sum(counts), *row_shape]`` tensor in rank order.  The only collective of a sharded generation."""
    import torch.distributed as dist
    world = len(counts)
    rank = dist.get_rank(group)
    pad = max(max(counts), 1)
    buf = torch.zeros(pad, *row_shape, dtype=torch.float32, device=device)
    if counts[rank]:
        buf[:counts[rank]].copy_(local)
    gathered = torch.empty(world * pad, *row_shape, dtype=torch.float32, device=device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(gathered, buf, group=group)
    else:                                   # gloo (CPU tests; GPU tests with both ranks on one device): staged through the host
        host = [torch.empty(pad, *row_shape, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(host, buf.cpu(), group=group)
        gathered.copy_(torch.cat(host, 0))
    if all(c == pad for c in counts):
        return gathered
    return torch.cat([gathered[r * pad:r * pad + c] for r, c in enumerate(counts)], 0)


def shard_outfits(n_outfits: int, rank: int, world_size: int):
    """Whole-outfit sharding: contiguous block of outfits for ``rank``; no outfit is ever split.
    Returns ``range`` of outfit indices."""
    base, rem = divmod(n_outfits, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
