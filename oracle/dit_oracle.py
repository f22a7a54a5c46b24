"""TEST INFRASTRUCTURE -- torch restatement of the reference DiT forward.

A functional (state-dict driven), dtype-generic restatement of
``/root/reference/f_lite/model.py`` that issues the *same torch ops in the same order* as
the reference, so that running it in bf16 reproduces the reference's bf16 rounding points
 and running it in fp32 on the CPU gives the fp32 oracle.  The three
third-party kernels the reference calls are restated from their published algorithms:

* ``LigerRMSNorm`` -- liger_kernel 0.8.0 ``ops/rms_norm.py`` forward kernel, "llama"
  casting mode, eps 1e-6, offset 0: x->fp32, rstd, x*rstd, cast to input dtype, then *weight
  in the input dtype.
* ``LigerSwiGLUMLP`` -- liger_kernel 0.8.0 ``ops/swiglu.py``: ``silu(a.fp32).cast(b.dtype) * b``.
* ``flash_attn_interface.flash_attn_varlen_func`` (FlashAttention-3, version unpinned by the
  reference, not installed here): per sequence ``softmax(q k^T * scale) v``, non-causal,
  fp32 softmax / accumulation, output in the input dtype (call site f_lite/model.py:203-210).

Pinned against the real module by ``oracle/make_golden.py`` -> ``tests/golden/*.pt``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .synth import cross_attn_blocks


# ----------------------------------------------------------------------------------------
# third-party kernel restatements
# ----------------------------------------------------------------------------------------
def liger_rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """liger_kernel/ops/rms_norm.py fwd kernel, casting_mode=llama (f_lite/model.py:238)."""
    xf = x.float()
    rstd = torch.rsqrt((xf * xf).sum(-1, keepdim=True) / x.shape[-1] + eps)
    return (xf * rstd).to(x.dtype) * weight


def rms_norm_ref(x: torch.Tensor, weight=None, eps: float = 1e-6) -> torch.Tensor:
    """f_lite/model.py:92-108 (RMSNorm used for QK-norm and final_norm)."""
    xd = x.dtype
    xf = x.float()
    norm = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if weight is not None:
        return (xf * norm * weight).to(xd)
    return (xf * norm).to(xd)


def liger_swiglu(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """liger_kernel/ops/swiglu.py:28-31."""
    af = a.float()
    return (af * torch.sigmoid(af)).to(b.dtype) * b


def flash_attn_varlen(q, k, v, cu_q, cu_k, scale):
    """Restated flash_attn_varlen_func: q (Tq,H,D), k/v (Tk,H,D), non-causal, per sequence."""
    out = torch.empty_like(q)
    cq = cu_q.tolist()
    ck = cu_k.tolist()
    for b in range(len(cq) - 1):
        qs = q[cq[b]:cq[b + 1]].float().transpose(0, 1)      # H, Lq, D
        ks = k[ck[b]:ck[b + 1]].float().transpose(0, 1)
        vs = v[ck[b]:ck[b + 1]].float().transpose(0, 1)
        s = torch.matmul(qs, ks.transpose(1, 2)) * scale
        p = torch.softmax(s, dim=-1)
        o = torch.matmul(p, vs)
        out[cq[b]:cq[b + 1]] = o.transpose(0, 1).to(q.dtype)
    return out


# ----------------------------------------------------------------------------------------
# model pieces (each cites the reference lines it follows)
# ----------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """f_lite/model.py:20-28."""
    half = dim // 2
    freqs = torch.exp(
        -math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half
    ).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def pack_varlen(hidden: torch.Tensor, mask=None):
    """f_lite/model.py:31-64 (prepare_flash_attention_inputs)."""
    b, l, d = hidden.shape
    if mask is None:
        mask = torch.ones((b, l), device=hidden.device)
    seqlens = mask.sum(dim=-1, dtype=torch.int32)
    cu = torch.cat([torch.zeros(1, dtype=torch.int32, device=hidden.device),
                    seqlens.cumsum(0, dtype=torch.int32)])
    idx = torch.nonzero(mask.reshape(-1), as_tuple=True)[0]
    return torch.index_select(hidden.reshape(-1, d), 0, idx), cu, l, idx


def unpack_varlen(flat, idx, b, l, d):
    """f_lite/model.py:67-89 (unprepare_flash_attention_outputs)."""
    out = torch.zeros((b * l, d), dtype=flat.dtype, device=flat.device)
    out.index_copy_(0, idx, flat)
    return out.view(b, l, d)


def rope_tables(head_dim: int, h: int, w: int, base: float, device, dtype, n_reg: int = 16):
    """f_lite/model.py:334-386 (TwoDimRotary with dim = head_dim/2, register rows cos=1 sin=0).

    ``dtype`` is the dtype the module's non-persistent buffers have: ``model.to(bf16)`` casts
    them to bf16, so a bf16 model rotates with bf16-rounded tables.
    """
    dim = head_dim // 2
    inv = torch.tensor([1.0 / (base ** (i / dim)) for i in range(0, dim, 2)], dtype=torch.float32)
    th = torch.arange(h, dtype=torch.float32)
    tw = torch.arange(w, dtype=torch.float32)
    fh = torch.outer(th, inv).unsqueeze(1).repeat(1, w, 1)
    fw = torch.outer(tw, inv).unsqueeze(0).repeat(h, 1, 1)
    f = torch.cat([fh, fw], 2)
    cos = f.cos().to(dtype).reshape(h * w, -1)
    sin = f.sin().to(dtype).reshape(h * w, -1)
    # f_lite/model.py:371-384: torch.ones/zeros are fp32 -> cat promotes to fp32
    cos = torch.cat([torch.ones(n_reg, cos.shape[1]), cos], 0)
    sin = torch.cat([torch.zeros(n_reg, sin.shape[1]), sin], 0)
    return cos[None].to(device), sin[None].to(device)


def apply_rotary_emb(x, cos, sin):
    """f_lite/model.py:403-414."""
    od = x.dtype
    x = x.float()
    cos = cos.float()
    sin = sin.float()
    d = x.shape[2] // 2
    x1, x2 = x[..., :d], x[..., d:]
    y1 = x1 * cos + x2 * sin
    y2 = x1 * (-sin) + x2 * cos
    return torch.cat([y1, y2], 2).to(od)


def attention(sd, pre, x, cu_x, num_heads, self_attn, rope=None, context=None, cu_ctx=None,
              attn_fn=flash_attn_varlen):
    """f_lite/model.py:160-213 (Attention.forward)."""
    d = x.shape[-1]
    hd = d // num_heads
    scale = hd ** -0.5
    if self_attn:
        qkv = F.linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
        qkv = qkv.view(-1, 3, num_heads, hd).permute(1, 2, 0, 3)   # "l (k h d) -> k h l d"
        q, k, v = qkv.unbind(0)
        if rope is not None:
            q = apply_rotary_emb(q, rope[0], rope[1])
            k = apply_rotary_emb(k, rope[0], rope[1])
        q = rms_norm_ref(q)
        k = rms_norm_ref(k)
        q, k, v = (t.permute(1, 0, 2) for t in (q, k, v))           # "h l d -> l h d"
        cu_q = cu_k = cu_x
    else:
        q = F.linear(x, sd[pre + "q.weight"], sd.get(pre + "q.bias")).view(-1, num_heads, hd)
        kv = F.linear(context, sd[pre + "context_kv.weight"], sd.get(pre + "context_kv.bias"))
        kv = kv.view(-1, 2, num_heads, hd).permute(1, 0, 2, 3)      # "l (k h d) -> k l h d"
        k, v = kv.unbind(0)
        q = rms_norm_ref(q)
        k = rms_norm_ref(k)
        cu_q, cu_k = cu_x, cu_ctx
    o = attn_fn(q.contiguous(), k.contiguous(), v.contiguous(), cu_q, cu_k, scale)
    o = o.reshape(-1, d)
    return F.linear(o, sd[pre + "proj.weight"])


def dit_block(sd, i, x, cu_x, ctx, cu_ctx, mod, rope, num_heads, has_cross, attn_fn):
    """f_lite/model.py:270-303 (DiTBlock.forward)."""
    (shift_sa, scale_sa, gate_sa, shift_ca, scale_ca, gate_ca, shift_mlp, scale_mlp, gate_mlp) = mod
    b = f"blocks.{i}."
    n = liger_rms_norm(x, sd[b + "norm1.weight"])
    n = n * (1 + scale_sa) + shift_sa
    a = attention(sd, b + "self_attn.", n, cu_x, num_heads, True, rope=rope, attn_fn=attn_fn)
    x = x + a * gate_sa
    if has_cross:
        n = liger_rms_norm(x, sd[b + "norm2.weight"])
        n = n * (1 + scale_ca) + shift_ca
        x = x + attention(sd, b + "cross_attn.", n, cu_x, num_heads, False,
                          context=ctx, cu_ctx=cu_ctx, attn_fn=attn_fn) * gate_ca
    n = liger_rms_norm(x, sd[b + "norm3.weight"])
    n = n * (1 + scale_mlp) + shift_mlp
    g = F.linear(n, sd[b + "mlp.gate_proj.weight"])
    u = F.linear(n, sd[b + "mlp.up_proj.weight"])
    y = F.linear(liger_swiglu(g, u), sd[b + "mlp.down_proj.weight"])
    x = x + y * gate_mlp
    return x


@torch.no_grad()
def dit_forward(sd: dict, cfg: dict, x, context, context_attn_mask, timesteps,
                rope_dtype=None, attn_fn=flash_attn_varlen, return_hidden: bool = False):
    """f_lite/model.py:525-591 (DiT.forward). ``sd`` uses the reference state-dict keys."""
    d = cfg["hidden_size"]
    p = cfg["patch_size"]
    nh = cfg["num_heads"]
    dtype = x.dtype
    if rope_dtype is None:
        rope_dtype = dtype
    ctx = F.linear(context, sd["context_proj.weight"], sd["context_proj.bias"])
    ctx = liger_rms_norm(ctx, sd["context_norm.weight"])
    ctx_flat, cu_ctx, _, _ = pack_varlen(ctx, context_attn_mask)

    b, c, h, w = x.shape
    t = F.conv2d(x, sd["patch_embed.patch_proj.weight"], sd["patch_embed.patch_proj.bias"], stride=p)
    t = t.flatten(2).transpose(1, 2)                                # "b c h w -> b (h w) c"
    t = torch.cat([sd["register_tokens"].repeat(b, 1, 1), t], 1)
    cos, sin = rope_tables(d // nh, h // p, w // p, cfg.get("rope_base", 10000), x.device, rope_dtype)
    cos = cos.repeat(1, b, 1)
    sin = sin.repeat(1, b, 1)
    x_flat, cu_x, lmax, x_idx = pack_varlen(t)

    t_emb = timestep_embedding(timesteps * 1000, d).to(x.device, dtype=dtype)
    t_emb = F.linear(F.silu(F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                     sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    L = 16 + h // p * w // p
    mod = F.linear(F.silu(t_emb), sd["adaLN_modulation.1.weight"], sd["adaLN_modulation.1.bias"])
    mod = mod.repeat_interleave(L, dim=0).chunk(9, dim=1)

    xs = set(cross_attn_blocks(cfg["depth"]))
    for i in range(cfg["depth"]):
        x_flat = dit_block(sd, i, x_flat, cu_x, ctx_flat, cu_ctx, mod, (cos, sin), nh, i in xs, attn_fn)
    hidden = x_flat
    t = unpack_varlen(x_flat, x_idx, b, lmax, d)[:, 16:, :]
    fmod = F.linear(F.silu(t_emb), sd["final_modulation.1.weight"], sd["final_modulation.1.bias"])
    fshift, fscale = fmod.chunk(2, dim=1)
    t = rms_norm_ref(t, sd.get("final_norm.weight"))
    t = t * (1 + fscale[:, None, :]) + fshift[:, None, :]
    t = F.linear(t, sd["final_proj.weight"], sd["final_proj.bias"])
    hh, ww = h // p, w // p
    ci = cfg["in_channels"]
    # "b (h w) (p1 p2 c) -> b c (h p1) (w p2)"
    out = t.view(b, hh, ww, p, p, ci).permute(0, 5, 1, 3, 2, 4).reshape(b, ci, hh * p, ww * p)
    if return_hidden:
        return out, hidden
    return out
