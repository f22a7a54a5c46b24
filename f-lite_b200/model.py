"""Drop-in ``DiT`` for the F Lite denoise path, running on hand-written sm_100a kernels.

Mirrors ``/root/reference/f_lite/model.py``:

* same constructor keywords and defaults (model.py:419-433), ``.config.<kw>`` access (diffusers
  ``register_to_config`` semantics), and the *same state-dict keys and shapes*  so
  ``new.load_state_dict(ref.state_dict())`` is the weight hand-off;
* ``forward(x, context, context_attn_mask, timesteps)`` (model.py:525-526); the legacy 3-argument call
  ``forward(x, context, timesteps)`` that ``FLitePipeline.__call__`` still makes (pipeline.py:271,293)
  is accepted too (mask := all ones);
* inference only (``torch.no_grad``), bf16 only -- there is no CPU or PyTorch fallback: every op is a
  kernel from ``libflite_b200.so``.

What is fused relative to the reference graph: RMSNorm+adaLN modulate (K5/K6),
QKV bias + RoPE + QK-norm in the GEMM epilogue (K1/K7/K9), gated residual adds in the GEMM epilogue
(K2/K8), SiLU*up in the gate/up GEMM epilogue (K3), no pack/unpack copies or host syncs (K14), per-sample
modulation never materialised per token (K6), RoPE table per (h, w) cached (K15), context projection /
K,V hoisted out of the step loop (K13).
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Optional

import torch
from torch import nn

from . import _lib, ops
from ._lib import (ATTN_XRES, EPI_GATED_RES, EPI_QKV_ROPE, EPI_STORE, EPI_SWIGLU, GEMM_AUTO, FliteError)

N_REGISTER = 16  # model.py:446,535,540


class _Config(dict):
    """dict with attribute access, like diffusers' FrozenDict behind ``register_to_config``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _NormWeight(nn.Module):
    """Holds ``weight`` only (LigerRMSNorm / RMSNorm state-dict layout, model.py:92-99,238)."""

    def __init__(self, dim: int, has_weight: bool = True):
        super().__init__()
        if has_weight:
            self.weight = nn.Parameter(torch.ones(dim))
        else:
            self.weight = None


class Attention(nn.Module):
    """Parameter container with the reference's names (model.py:133-158)."""

    def __init__(self, dim: int, num_heads: int, qkv_bias: bool, is_self_attn: bool):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.is_self_attn = is_self_attn
        if is_self_attn:
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        else:
            self.q = nn.Linear(dim, dim, bias=qkv_bias)
            self.context_kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=False)


class _SwiGLU(nn.Module):
    """LigerSwiGLUMLP parameter layout (model.py:262-267,312-314)."""

    def __init__(self, hidden: int, inter: int):
        super().__init__()
        self.gate_proj = nn.Linear(hidden, inter, bias=False)
        self.up_proj = nn.Linear(hidden, inter, bias=False)
        self.down_proj = nn.Linear(inter, hidden, bias=False)


class DiTBlock(nn.Module):
    """model.py:226-267."""

    def __init__(self, hidden_size, num_heads, do_cross_attn, mlp_ratio, qkv_bias):
        super().__init__()
        self.norm1 = _NormWeight(hidden_size)
        self.self_attn = Attention(hidden_size, num_heads, qkv_bias, True)
        if do_cross_attn:
            self.norm2 = _NormWeight(hidden_size)
            self.cross_attn = Attention(hidden_size, num_heads, qkv_bias, False)
        else:
            self.norm2 = None
            self.cross_attn = None
        self.norm3 = _NormWeight(hidden_size)
        self.mlp = _SwiGLU(hidden_size, int(hidden_size * mlp_ratio))


class PatchEmbed(nn.Module):
    """model.py:318-322."""

    def __init__(self, patch_size, in_channels, embed_dim):
        super().__init__()
        self.patch_proj = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.patch_size = patch_size


def _rope_table(head_dim: int, h: int, w: int, base: float, n_reg: int, round_bf16: bool):
    """cos/sin [n_reg + h*w, head_dim/2], values as the reference module holds them (model.py:334-386;
    the buffers are bf16 after ``model.to(bf16)``).  Built with the same CPU torch
    ops as the reference constructor so the table is bit-identical; returned in fp32 (``round_bf16`` only
    rounds the values), the model stores them as bf16 on the device."""
    dim = head_dim // 2
    inv = torch.tensor([1.0 / (base ** (i / dim)) for i in range(0, dim, 2)], dtype=torch.float32)
    fh = torch.outer(torch.arange(h, dtype=torch.float32), inv).unsqueeze(1).repeat(1, w, 1)
    fw = torch.outer(torch.arange(w, dtype=torch.float32), inv).unsqueeze(0).repeat(h, 1, 1)
    f = torch.cat([fh, fw], 2).reshape(h * w, dim)
    cos, sin = f.cos(), f.sin()
    if round_bf16:
        cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    cos = torch.cat([torch.ones(n_reg, dim), cos], 0)
    sin = torch.cat([torch.zeros(n_reg, dim), sin], 0)
    return cos.contiguous(), sin.contiguous()


class DiT(nn.Module):
    def __init__(
        self,
        in_channels=4,
        patch_size=2,
        hidden_size=1152,
        depth=28,
        num_heads=16,
        mlp_ratio=4.0,
        cross_attn_input_size=128,
        train_bias_and_rms=True,
        use_rope=True,
        gradient_checkpoint=False,
        dynamic_softmax_temperature=False,
        rope_base=10000,
    ):
        super().__init__()
        self.config = _Config(
            in_channels=in_channels, patch_size=patch_size, hidden_size=hidden_size, depth=depth,
            num_heads=num_heads, mlp_ratio=mlp_ratio, cross_attn_input_size=cross_attn_input_size,
            train_bias_and_rms=train_bias_and_rms, use_rope=use_rope, gradient_checkpoint=gradient_checkpoint,
            dynamic_softmax_temperature=dynamic_softmax_temperature, rope_base=rope_base,
        )
        if hidden_size % num_heads or hidden_size // num_heads != 256:
            raise FliteError("the sm_100a attention kernel is built for head_dim 256 "
                             "(F Lite uses num_heads = hidden_size // 256, f_lite/train.py:690)")
        if not use_rope:
            raise FliteError("use_rope=False (learned positional embedding, model.py:444,546) is not on the "
                             "F Lite hot path and is not implemented")
        self.context_proj = nn.Linear(cross_attn_input_size, hidden_size)
        self.context_norm = _NormWeight(hidden_size)
        self.patch_embed = PatchEmbed(patch_size, in_channels, hidden_size)
        self.register_tokens = nn.Parameter(torch.randn(1, N_REGISTER, hidden_size))
        self.time_embed = nn.Sequential(
            nn.Linear(hidden_size, 4 * hidden_size), nn.SiLU(), nn.Linear(4 * hidden_size, hidden_size))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 9 * hidden_size, bias=True))
        self.adaLN_modulation[-1].weight.data.zero_()
        self.adaLN_modulation[-1].bias.data.zero_()
        self.blocks = nn.ModuleList([
            DiTBlock(hidden_size, num_heads, (idx % 4 == 0 or idx < 8), mlp_ratio, train_bias_and_rms)
            for idx in range(depth)
        ])
        self.final_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))
        self.final_norm = _NormWeight(hidden_size, has_weight=train_bias_and_rms)
        self.final_proj = nn.Linear(hidden_size, patch_size * patch_size * in_channels)
        for t in (self.final_modulation[-1].weight, self.final_modulation[-1].bias, self.final_proj.weight,
                  self.final_proj.bias):
            nn.init.zeros_(t)
        # host-side caches (not part of the state dict)
        self._rope_cache = {}
        self._ws = {}
        self._gu_cache = {}
        self._kvcat_cache = None
        self._ctx_cache = None
        self._freqs = None
        self.hoist_context = True
        # Self-attention schedule (every image sequence has the same length, so the persistent kernel applies):
        #   "rr" (default)  flite_attention_streamk with whole 256-query units handed out round-robin to ONE wave of
        #                   clusters: no unit is ever split, so the result is bit-identical to the one-cluster-per-unit launch
        #                   and batch-invariant (which is what keeps the multi-GPU layouts bit-identical to the 1-GPU path);
        #                   barrier init / TMEM allocation happen once and the next unit's Q / K loads start under the
        #                   previous unit's tail.  Measured (profiles/r2x_attn_probe.json, r2x_ab_step_c2.json): +3.4 % on the
        #                   C2 launch, +4.7 % at C5, 0.7 ms off the C2 step.
        #   "0"             flite_attention_varlen with FLITE_ATTN_AUTO: the same persistent kernel in its ragged-length
        #                   mode (what the cross-attention calls use); FLITE_TUNE_ATTN_VARIANT = 5 selects the round-1 path,
        #                   one cluster per unit.
        #   "1"             stream-K shares: a unit split between two clusters is merged in fp32, so the result depends on
        #                   where the shares fall (batch size / head count); loses 8 % at C2 (the clusters no longer walk the
        #                   same key tiles in lock step: 1.0 GB of DRAM reads per launch instead of 0.15).
        #   "hybrid"        whole rounds in lock step, stream-K shares over the last 1..2 units per cluster: same caveat
        #                   (not batch-invariant), but it removes the last-round quantisation without the L2 cost:
        #                   -6 % on a sequence-parallel rank's 2 x 3 heads x 16400 tokens (5.27 -> 6 rounds), -3 % at C4.
        #   "auto"          "hybrid" for long sequences whose last round is >= 10 % empty, else "rr".
        self.attn_streamk = os.environ.get("FLITE_ATTN_STREAMK", "rr")
        self.gemm_variant = GEMM_AUTO
        self.sp_group = None      # Ulysses sequence-parallel process group (see enable_sequence_parallel)
        self.sp_fused = False     # exchanges fused into the kernels over NVLink peer memory instead of NCCL
        self._sp_sym = None

    # ------------------------------------------------------------------ helpers
    @property
    def dtype(self):
        return self.context_proj.weight.dtype

    @property
    def device(self):
        return self.context_proj.weight.device

    def enable_sequence_parallel(self, group, fused: bool = False):
        """Ulysses sequence parallelism over ``group``: every rank holds
        L/P tokens of each sequence; self-attention is computed head-sharded over the full sequence with two
        exchanges per block; everything else is token-local.  ``None`` disables it.

        ``fused=False``: the exchanges are NCCL ``all_to_all_single`` calls.  ``fused=True``: the QKV GEMM epilogue
        and the attention epilogue store straight into the destination rank's buffer over NVLink peer memory
        (``peer.SymmetricBuffer``; CUDA IPC, all ranks on one node) and only a flag handshake separates the kernels,
        so the transfer overlaps the math tile by tile and the permute/copy kernels disappear."""
        if group is not None:
            import torch.distributed as dist
            P = dist.get_world_size(group)
            if self.config.num_heads % P:
                raise FliteError(f"sequence parallel degree {P} must divide num_heads {self.config.num_heads}")
        if self._sp_sym is not None:
            self._sp_sym[1].close()
            self._sp_sym = None
        self.sp_group = group
        self.sp_fused = bool(fused) and group is not None

    def _sym(self, B, L, Lq, dq, d, dev):
        """Symmetric peer buffer: [receive q|k|v of my heads for the full sequences | returned attention rows]."""
        key = (B, L, Lq, dq, d, str(dev))
        if self._sp_sym is None or self._sp_sym[0] != key:
            from .peer import SymmetricBuffer
            if self._sp_sym is not None:
                self._sp_sym[1].close()
            recv_elems = B * L * 3 * dq
            sym = SymmetricBuffer(self.sp_group, 2 * (recv_elems + B * Lq * d), dev)
            recv = sym.local[:recv_elems].view(B * L, 3 * dq)
            ao = sym.local[recv_elems:recv_elems + B * Lq * d].view(B * Lq, d)
            self._sp_sym = (key, sym, recv, ao, sym.table(0), sym.table(2 * recv_elems))
        return self._sp_sym[1:]

    def _use_streamk(self, B, heads, L):
        """Picks the schedule of the persistent self-attention kernel (FLITE_TUNE_ATTN_SK_MODE) for this call; False =
        go through flite_attention_varlen (FLITE_ATTN_AUTO) instead."""
        mode = self.attn_streamk
        if mode == "auto":   # long sequences whose whole-unit rounds waste >= 10 % of the last one: hybrid, else round-robin
            mode = "hybrid" if self._use_streamk_auto(B, heads, L) else "rr"
        sk = {"rr": 1, "hybrid": 2, "1": 0, 1: 0, True: 0}.get(mode)
        if sk is None:
            return False
        _lib.load().flite_set_tuning(15, sk)
        return True

    def _use_streamk_auto(self, B, heads, L):
        units = B * heads * ((L + 255) // 256)
        slots = max(1, torch.cuda.get_device_properties(self.device).multi_processor_count // 2)
        waves = units / slots
        return L >= 8192 and units >= slots and math.ceil(waves) / waves >= 1.10

    def _check_ready(self):
        w = self.context_proj.weight
        if not w.is_cuda:
            raise FliteError("flite_b200.DiT runs on CUDA (sm_100a) only; call .to('cuda') -- there is no CPU path")
        if w.dtype != torch.bfloat16:
            raise FliteError("flite_b200.DiT computes in bf16; call .to(torch.bfloat16)")

    def _rope(self, h, w, device):
        key = (h, w, str(device))
        if key not in self._rope_cache:
            hd = self.config.hidden_size // self.config.num_heads
            cos, sin = _rope_table(hd, h, w, self.config.rope_base, N_REGISTER, round_bf16=True)
            self._rope_cache[key] = (cos.to(device, torch.bfloat16), sin.to(device, torch.bfloat16))
        return self._rope_cache[key]

    def _rope_slice(self, h, w, device, l0, n):
        key = (h, w, str(device), l0, n)
        if key not in self._rope_cache:
            cos, sin = self._rope(h, w, device)
            self._rope_cache[key] = (cos[l0:l0 + n].contiguous(), sin[l0:l0 + n].contiguous())
        return self._rope_cache[key]

    @staticmethod
    def _ver(t):
        """Cache-key component for a tensor: (address, version counter).  Inference tensors (created under
        ``torch.inference_mode()``, the usual way to load weights / run the text encoder) have no version counter:
        their version reads as None, so in-place updates to them are invisible to the caches -- call
        ``invalidate_caches()`` after such an update (``load_state_dict`` / ``.to()`` do it themselves)."""
        if t is None:
            return None
        return (t.data_ptr(), None if t.is_inference() else t._version, tuple(t.shape), t.dtype)

    def invalidate_caches(self):
        """Drop every derived tensor (interleaved gate|up weights, concatenated context_kv weights, hoisted context
        K/V).  Called by ``load_state_dict`` and ``_apply`` (``.to()`` / ``.cuda()`` / ``.bfloat16()``); call it by hand
        after an in-place weight edit made under ``torch.inference_mode()`` (e.g. a LoRA merge)."""
        self._gu_cache = {}
        self._kvcat_cache = None
        self._ctx_cache = None

    def release_workspaces(self):
        """Free the activation workspaces, RoPE tables and the sequence-parallel peer buffer (collective over the
        sequence-parallel group when one is set)."""
        self._ws = {}
        self._rope_cache = {}
        if self._sp_sym is not None:
            self._sp_sym[1].close()
            self._sp_sym = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_caches()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_gu_cache"):
            self.invalidate_caches()
        return out

    def _gate_up(self, i, blk):
        g, u = blk.mlp.gate_proj.weight, blk.mlp.up_proj.weight
        key = (self._ver(g), self._ver(u))
        hit = self._gu_cache.get(i)
        if hit is None or hit[0] != key:
            hit = (key, ops.interleave_gate_up(g.detach(), u.detach()))
            self._gu_cache[i] = hit
        return hit[1]

    def _context_kv_cat(self, cross):
        """[K rows of every cross block ; V rows of every cross block] (+ bias), cached until a parameter changes."""
        lins = [self.blocks[i].cross_attn.context_kv for i in cross]
        key = tuple((self._ver(l.weight), self._ver(l.bias)) for l in lins)
        hit = self._kvcat_cache
        if hit is None or hit[0] != key:
            d = self.config.hidden_size
            w = torch.cat([l.weight.detach()[:d] for l in lins] + [l.weight.detach()[d:] for l in lins]).contiguous()
            b = None
            if lins[0].bias is not None:
                b = torch.cat([l.bias.detach()[:d] for l in lins] + [l.bias.detach()[d:] for l in lins]).contiguous()
            hit = (key, w, b)
            self._kvcat_cache = hit
        return hit[1], hit[2]

    def _buf(self, name, shape, device):
        """Workspace ``name`` viewed as ``shape``: one flat allocation per name that only ever grows to the largest
        request (a new batch size / resolution re-uses or replaces it instead of pinning another full set)."""
        n = 1
        for s_ in shape:
            n *= int(s_)
        flat = self._ws.get(name)
        if flat is None or flat.device != device or flat.numel() < n:
            flat = torch.empty(max(n, 1), dtype=torch.bfloat16, device=device)
            self._ws[name] = flat
        return flat[:n].view(shape)

    # ------------------------------------------------------------------ context (t-independent, K13)
    def prepare_context(self, context: torch.Tensor, mask: Optional[torch.Tensor]):
        """context_proj -> context_norm -> varlen pack -> per cross block K (normed) / V.
        model.py:527-530,190-197.  Independent of the timestep, so computed once per prompt batch."""
        cfg = self.config
        d = cfg.hidden_size
        B, Lc, ci = context.shape
        ctx2 = context.reshape(B * Lc, ci)
        if ctx2.dtype != torch.bfloat16:
            ctx2 = ctx2.to(torch.bfloat16)
        c = ops.gemm(ctx2, self.context_proj.weight, self.context_proj.bias, variant=self.gemm_variant)
        c = ops.rmsnorm_modulate(c, self.context_norm.weight, 1)
        if mask is None:
            mask_f = torch.ones((B, Lc), dtype=torch.float32, device=context.device)
        else:
            mask_f = mask.to(torch.float32)
        packed, cu_k = ops.pack_context(c.view(B, Lc, d), mask_f)
        # context_kv of ALL cross-attention blocks in one GEMM (model.py:190-195 runs one Linear(d, 2d) per block): the
        # weights are concatenated once as [K rows of every block ; V rows of every block], so that the QK-norm columns
        # are the leading ones and block j's K / V are column slices j*d and (X + j)*d of one [Tc, 2*X*d] buffer.
        cross = [i for i, blk in enumerate(self.blocks) if blk.cross_attn is not None]
        kvs = {}
        if cross:
            X = len(cross)
            wcat, bcat = self._context_kv_cat(cross)
            kv_all = ops.gemm(packed, wcat, bcat, epilogue=EPI_QKV_ROPE, qk_cols=X * d, variant=self.gemm_variant)
            for j, i in enumerate(cross):
                kvs[i] = (kv_all[:, j * d:(j + 1) * d], kv_all[:, (X + j) * d:(X + j + 1) * d])
        return SimpleNamespace(kvs=kvs, cu_k=cu_k, B=B, Lc=Lc)

    def _context(self, context, mask):
        # inference tensors carry no version counter, so an in-place change of the embeddings could not be seen: do not
        # hoist for them (0.25 % of a step's FLOPs); callers that want the hoist anyway use prepare_context() themselves
        if not self.hoist_context or context.is_inference() or (mask is not None and mask.is_inference()):
            return self.prepare_context(context, mask)
        cross = [b.cross_attn.context_kv for b in self.blocks if b.cross_attn is not None]
        key = (self._ver(context), self._ver(mask),
               tuple(self._ver(p) for p in (self.context_proj.weight, self.context_proj.bias, self.context_norm.weight)),
               tuple((self._ver(l.weight), self._ver(l.bias)) for l in cross))
        if self._ctx_cache is None or self._ctx_cache[0] != key:
            # keep references to the keyed tensors so their storage cannot be recycled under the cache
            self._ctx_cache = (key, self.prepare_context(context, mask), context, mask)
        return self._ctx_cache[1]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, context, context_attn_mask=None, timesteps=None):
        if timesteps is None:  # legacy call form dit(x, context, t)  (pipeline.py:271,293)
            if context_attn_mask is None:
                raise TypeError("forward() missing timesteps")
            timesteps, context_attn_mask = context_attn_mask, None
        self._check_ready()
        cfg = self.config
        d, p, nh = cfg.hidden_size, cfg.patch_size, cfg.num_heads
        dev = x.device
        B, C, H, W = x.shape
        hp, wp = H // p, W // p
        L = N_REGISTER + hp * wp
        v = self.gemm_variant

        ctx = self._context(context, context_attn_mask)
        if ctx.B != B:
            raise FliteError(f"context batch {ctx.B} != latent batch {B}")

        # --- sequence-parallel layout: this rank owns tokens [l0, l0 + Lq) of every sequence
        sp = self.sp_group
        if sp is not None:
            import torch.distributed as dist
            P, rk = dist.get_world_size(sp), dist.get_rank(sp)
            if L % P:
                raise FliteError(f"sequence length {L} is not divisible by the sequence-parallel degree {P}")
            Lq, l0 = L // P, rk * (L // P)
        else:
            P, rk, Lq, l0 = 1, 0, L, 0
        T = B * Lq

        # --- tokens: patchify + register tokens (model.py:533-535)
        xin = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        xs = self._buf("x", (T, d), dev)
        ops.patch_embed(xin.contiguous(), self.patch_embed.patch_proj.weight, self.patch_embed.patch_proj.bias,
                        self.register_tokens, p, out=xs, tok_offset=l0, tok_count=Lq)
        cos, sin = self._rope_slice(hp, wp, dev, l0, Lq)
        cu_x = self._ws.get(("cu_x", B, Lq))
        if cu_x is None or cu_x.device != dev:
            cu_x = (torch.arange(0, B + 1, dtype=torch.int32) * Lq).to(dev)
            self._ws[("cu_x", B, Lq)] = cu_x
        cu_full = cu_x
        if sp is not None:
            cu_full = self._ws.get(("cu_x", B, L))
            if cu_full is None or cu_full.device != dev:
                cu_full = (torch.arange(0, B + 1, dtype=torch.int32) * L).to(dev)
                self._ws[("cu_x", B, L)] = cu_full

        # --- timestep path (model.py:551-556,578): sinusoid -> MLP -> SiLU -> adaLN / final modulation
        if self._freqs is None or self._freqs.device != dev:
            half = d // 2
            self._freqs = torch.exp(
                -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
        if timesteps.dtype == torch.bfloat16:
            t32, tflag = timesteps.float(), 1
        elif timesteps.dtype == torch.float32:
            t32, tflag = timesteps, 0
        else:  # any other dtype: do `timesteps * 1000` in the caller's dtype like model.py:551
            t32, tflag = (timesteps * 1000).float(), 2
        temb = ops.timestep_embed(t32.to(dev).contiguous(), tflag, self._freqs, d)
        te0, te2 = self.time_embed[0], self.time_embed[2]
        h1 = ops.gemm(temb, te0.weight, te0.bias, act=1, variant=v)
        st = ops.gemm(h1, te2.weight, te2.bias, act=1, variant=v)           # silu(t_emb)
        ada = self.adaLN_modulation[1]
        mod = ops.gemm(st, ada.weight, ada.bias, variant=v)                  # [B, 9d]
        fm = self.final_modulation[1]
        fmod = ops.gemm(st, fm.weight, fm.bias, variant=v)                   # [B, 2d]
        (shift_sa, scale_sa, gate_sa, shift_ca, scale_ca, gate_ca, shift_mlp, scale_mlp, gate_mlp) = (
            mod[:, k * d:(k + 1) * d] for k in range(9))

        nbuf = self._buf("n", (T, d), dev)
        abuf = self._buf("attn", (T, d), dev)
        qc = self._buf("qc", (T, d), dev)
        inter = self.blocks[0].mlp.gate_proj.weight.shape[0] if len(self.blocks) else 0
        hmid = self._buf("hmid", (T, inter), dev)
        scale = (d // nh) ** -0.5
        if sp is None:
            qkv = self._buf("qkv", (T, 3 * d), dev)
        else:
            hq, dq = nh // P, d // P                                 # heads / width of this rank's head group
        if sp is not None and self.sp_fused:
            sym, p2p_recv, p2p_ao, recv_tab, ao_tab = self._sym(B, L, Lq, dq, d, dev)
            stream = torch.cuda.current_stream().cuda_stream
            # Host-paced skew between the ranks (one of them saving an image, a GC pause, first-call lazy loading) is
            # absorbed HERE by NCCL, which tolerates minutes: one 4-byte all-reduce orders the ranks' streams before the
            # first peer store of this forward, so the device-side flag spins below only ever cover kernel-level skew.
            tok = self._ws.get("sp_token")
            if tok is None or tok.device != dev:
                tok = torch.zeros(1, dtype=torch.int32, device=dev)
                self._ws["sp_token"] = tok
            dist.all_reduce(tok, group=sp)
        elif sp is not None:
            a2a_send = self._buf("a2a_send", (B, P * Lq, 3 * dq), dev)   # [sample][dest rank][local token][q|k|v]
            a2a_recv = self._buf("a2a_recv", (B, L, 3 * dq), dev)        # [sample][full sequence][q|k|v of my heads]
            ao_full = self._buf("ao_full", (B, L, dq), dev)
            ao_recv = self._buf("ao_recv", (B, P, Lq * dq), dev)         # [sample][source rank][local token, head group]

        # Cross-attention goes through flite_attention_varlen with FLITE_ATTN_AUTO = the persistent kernel with per-sequence
        # key lengths (72 -> 48 us per launch at C2, bit-identical to the per-unit kernel).  The round-1 resident-K/V kernel
        # (FLITE_ATTN_XRES, <= 256 context tokens) stays selectable through FLITE_TUNE_ATTN_VARIANT_SHORT_K = 9; it is
        # slower than the default now (75 us).
        x_variant = ATTN_XRES if (ctx.Lc <= 256 and ops.get_tuning(12) == ATTN_XRES) else 0
        for i, blk in enumerate(self.blocks):
            # ---- self-attention (model.py:283-289)
            ops.rmsnorm_modulate(xs, blk.norm1.weight, 1, scale_sa, shift_sa, rows_per_sample=Lq, out=nbuf)
            sa = blk.self_attn
            if sp is None:
                ops.gemm(nbuf, sa.qkv.weight, sa.qkv.bias, epilogue=EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin,
                         qk_cols=2 * d, rows_per_sample=Lq, variant=v, out=qkv)
                if self._use_streamk(B, nh, L):   # every image sequence has L tokens: one persistent stream-K wave
                    ops.attention_streamk(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu_x, cu_x, nh, L, L, scale, out=abuf)
                else:
                    ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu_x, cu_x, nh, L, scale, out=abuf)
            elif self.sp_fused:
                # Ulysses over peer memory: the QKV epilogue stores each head into its owner's receive buffer, the
                # attention epilogue stores each query row into the token owner's buffer; flags order the kernels.
                ops.gemm_qkv_p2p(nbuf, sa.qkv.weight, sa.qkv.bias, cos, sin, Lq, P, hq, rk, L, recv_tab, variant=v)
                sym.exchange_done(stream)
                if self._use_streamk(B, hq, L):
                    ops.attention_streamk_p2p(p2p_recv[:, :dq], p2p_recv[:, dq:2 * dq], p2p_recv[:, 2 * dq:], cu_full,
                                              cu_full, hq, L, L, scale, ao_tab, P, Lq, rk * hq, d)
                else:
                    ops.attention_varlen_p2p(p2p_recv[:, :dq], p2p_recv[:, dq:2 * dq], p2p_recv[:, 2 * dq:], cu_full,
                                             cu_full, hq, L, scale, ao_tab, P, Lq, rk * hq, d)
                sym.exchange_done(stream)
                ops.gemm(p2p_ao, sa.proj.weight, None, epilogue=EPI_GATED_RES, resid=xs, gate=gate_sa,
                         rows_per_sample=Lq, variant=v, out=xs)
            else:
                # Ulysses: the QKV epilogue scatters heads into the all-to-all send layout; after the exchange this
                # rank holds q|k|v of its hq heads for the FULL sequence; the second exchange returns the outputs.
                ops.gemm(nbuf, sa.qkv.weight, sa.qkv.bias, epilogue=EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin,
                         qk_cols=2 * d, rows_per_sample=Lq, variant=v, out=a2a_send.view(B * P * Lq, 3 * dq),
                         sp_ranks=P, sp_heads_per_rank=hq)
                for b in range(B):
                    dist.all_to_all_single(a2a_recv[b], a2a_send[b], group=sp)
                r2 = a2a_recv.view(B * L, 3 * dq)
                if self._use_streamk(B, hq, L):
                    ops.attention_streamk(r2[:, :dq], r2[:, dq:2 * dq], r2[:, 2 * dq:], cu_full, cu_full, hq, L, L, scale,
                                          out=ao_full.view(B * L, dq))
                else:
                    ops.attention_varlen(r2[:, :dq], r2[:, dq:2 * dq], r2[:, 2 * dq:], cu_full, cu_full, hq, L, scale,
                                         out=ao_full.view(B * L, dq))
                for b in range(B):
                    dist.all_to_all_single(ao_recv[b], ao_full[b], group=sp)
                for b in range(B):   # [source rank][token][dq] -> [token][source rank * dq] = head-major columns
                    ops.permute_021(ao_recv[b].view(P, Lq, dq), out=abuf[b * Lq:(b + 1) * Lq].view(Lq, P, dq))
            if not (sp is not None and self.sp_fused):
                ops.gemm(abuf, sa.proj.weight, None, epilogue=EPI_GATED_RES, resid=xs, gate=gate_sa,
                         rows_per_sample=Lq, variant=v, out=xs)
            # ---- cross-attention (model.py:291-297): token-local, context K/V replicated
            if blk.cross_attn is not None:
                ca = blk.cross_attn
                ops.rmsnorm_modulate(xs, blk.norm2.weight, 1, scale_ca, shift_ca, rows_per_sample=Lq, out=nbuf)
                ops.gemm(nbuf, ca.q.weight, ca.q.bias, epilogue=EPI_QKV_ROPE, qk_cols=d, rows_per_sample=Lq,
                         variant=v, out=qc)
                ck, cv = ctx.kvs[i]
                ops.attention_varlen(qc, ck, cv, cu_x, ctx.cu_k, nh, Lq, scale, out=abuf, variant=x_variant)
                ops.gemm(abuf, ca.proj.weight, None, epilogue=EPI_GATED_RES, resid=xs, gate=gate_ca,
                         rows_per_sample=Lq, variant=v, out=xs)
            # ---- SwiGLU MLP (model.py:299-301)
            ops.rmsnorm_modulate(xs, blk.norm3.weight, 1, scale_mlp, shift_mlp, rows_per_sample=Lq, out=nbuf)
            ops.gemm(nbuf, self._gate_up(i, blk), None, epilogue=EPI_SWIGLU, variant=v, out=hmid)
            ops.gemm(hmid, blk.mlp.down_proj.weight, None, epilogue=EPI_GATED_RES, resid=xs, gate=gate_mlp,
                     rows_per_sample=Lq, variant=v, out=xs)

        # ---- final head (model.py:577-590)
        fshift, fscale = fmod[:, :d], fmod[:, d:]
        fw = self.final_norm.weight
        ops.rmsnorm_modulate(xs, fw, 2 if fw is not None else 0, fscale, fshift, rows_per_sample=Lq, out=nbuf)
        o = ops.gemm(nbuf, self.final_proj.weight, self.final_proj.bias, variant=v)
        if sp is not None:
            # gather every rank's token slice: [rank][sample][Lq*64] -> [sample][rank][Lq*64] = [B*L, 64]
            no = o.shape[1]
            gathered = torch.empty((P, B, Lq * no), dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(gathered.view(P * B * Lq, no), o, group=sp)
            o = ops.permute_021(gathered).view(B * L, no)
        out = ops.unpatchify(o, B, C, H, W, p, N_REGISTER)
        if sp is not None and self.sp_fused:
            # fail closed: a timed-out peer wait leaves a half-exchanged buffer behind; callers of forward() that never
            # look at flite_watchdog_status (denoise() does, once per trajectory) must not get plausible numbers from it
            ops.poison_on_abort(out)
        return out if x.dtype == torch.bfloat16 else out.to(x.dtype)
