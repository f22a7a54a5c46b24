"""Multi-GPU execution of the denoise path on one NVSwitch box (one process per GPU, torch.distributed).

The reference has no multi-GPU inference at all; what exists here follows SURVEY.md 8(e):

* **Data parallel over prompts** (configs C3 / C5): every image's trajectory is independent, both CFG halves of a
  prompt stay on the same rank, weights are replicated => *no collective on the data path*; only an optional final
  gather of the finished latents.
* **Ulysses sequence parallel** (config C4, single 2048^2 image): tokens are sharded L/P per rank everywhere except
  self-attention, which is head-sharded over the full sequence; two all-to-alls per block over NCCL/NVLink
  (``DiT.enable_sequence_parallel`` in ``model.py``; the fused peer-memory exchange lives in ``peer.py`` and the
  ``flite_gemm_qkv_p2p`` / ``flite_attention_varlen_p2p`` kernels).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of ``n_items`` units: the first ``n_items % world`` ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def dp_denoise(denoise_fn: Callable, latents: torch.Tensor, negative_embeds: torch.Tensor,
               prompt_embeds: torch.Tensor, mask: Optional[torch.Tensor], *, gather: bool = True,
               group=None, **kwargs) -> torch.Tensor:
    """Shard the prompt batch over the ranks of ``group`` and run ``denoise_fn`` on the local shard.

    ``latents`` (B, C, h, w), ``negative_embeds`` / ``prompt_embeds`` (B, Lc, ci), ``mask`` (2B, Lc) ordered
    ``[negative rows ; positive rows]`` like ``flite_b200.denoise``.  Every rank passes the *full* batch (cheap: the
    inputs are a few MB) and computes only ``shard_range(B, rank, world)``.  With ``gather`` the finished latents of
    all ranks are all-gathered (the only collective, after the last step); otherwise the local shard is returned.
    """
    if not dist.is_available() or not dist.is_initialized():
        return denoise_fn(latents, negative_embeds, prompt_embeds, mask, **kwargs)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = latents.shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        m = None if mask is None else torch.cat([mask[lo:hi], mask[B + lo:B + hi]])
        local = denoise_fn(latents[lo:hi], negative_embeds[lo:hi], prompt_embeds[lo:hi], m, **kwargs)
    else:
        local = latents[:0]
    if not gather:
        return local
    # a rank with an empty shard does not know the dtype denoise_fn returns on the others (bf16 latents or the fp32
    # accumulator): agree on it first (max over ranks of a dtype code, 0 = "no result here")
    codes = [None, torch.bfloat16, torch.float16, torch.float32, torch.float64]
    code = torch.tensor([codes.index(local.dtype) if (hi > lo and local.dtype in codes) else 0], device=latents.device)
    dist.all_reduce(code, op=dist.ReduceOp.MAX, group=group)
    if int(code.item()) > 0:
        local = local.to(codes[int(code.item())])
    # ragged all-gather: pad every shard to the largest one
    cap = (B + world - 1) // world
    pad = torch.zeros((cap,) + tuple(latents.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(world):
        a, b = shard_range(B, r, world)
        out.append(parts[r][: b - a])
    return torch.cat(out, 0)


def make_groups(cfg_ranks: int = 1, sp_ranks: int = 1):
    """Split the world into data-parallel replicas of (cfg_ranks x sp_ranks) GPUs.

    Rank layout inside a replica: ``r = cfg_idx * sp_ranks + sp_idx`` so that the ranks of a sequence-parallel group
    are adjacent.  Returns ``(sp_group, cfg_group, replica_index, n_replicas)``; a group of size 1 is ``None``.
    Every process must call this with the same arguments (``dist.new_group`` is collective).  The natural 8-GPU
    layout for one 2048^2 image with 12 heads is ``cfg_ranks=2, sp_ranks=4``.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    per = cfg_ranks * sp_ranks
    if world % per:
        raise ValueError(f"world size {world} is not a multiple of cfg_ranks*sp_ranks = {per}")
    sp_group = cfg_group = None
    for rep in range(world // per):
        base = rep * per
        for c in range(cfg_ranks):
            ranks = [base + c * sp_ranks + s for s in range(sp_ranks)]
            g = dist.new_group(ranks) if sp_ranks > 1 else None
            if rank in ranks:
                sp_group = g
        for s_ in range(sp_ranks):
            ranks = [base + c * sp_ranks + s_ for c in range(cfg_ranks)]
            g = dist.new_group(ranks) if cfg_ranks > 1 else None
            if rank in ranks:
                cfg_group = g
    return sp_group, cfg_group, rank // per, world // per
