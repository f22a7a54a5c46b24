"""Deterministic, platform-independent synthetic tensors (TEST INFRASTRUCTURE).

The parity harness needs identical weights / inputs in the build container (where the
reference module can be imported) and on the GPU box (where it cannot), without
committing hundreds of MB of tensors.  ``torch.randn`` streams differ between CPU and
CUDA generators, so values are derived from a splitmix64 hash of the flat element index
instead: the same (seed, shape) gives bit-identical fp32 values on any device.

Default-init conventions follow the reference constructor (f_lite/model.py:419-479):
nn.Linear / Conv2d default init is U(-1/sqrt(fan_in), 1/sqrt(fan_in)); the tensors the
reference zero-initialises (adaLN_modulation[-1], final_modulation[-1], final_proj;
f_lite/model.py:455-456,476-479) are re-drawn with std 0.02 because with zeros every gate
is 0 and the model output is identically 0, which would make parity vacuous.
"""
from __future__ import annotations

import math
import zlib

import torch

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    """Python int -> two's complement int64 value."""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """Logical shift right on int64 tensors."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def hash_u01(numel: int, seed: int, device="cpu") -> torch.Tensor:
    """splitmix64(index + seed * golden) -> float32 uniform in [0, 1) with 24 bits."""
    out = torch.empty(numel, dtype=torch.float32, device=device)
    chunk = 1 << 24
    base = _s64(seed * 0x9E3779B97F4A7C15)
    for s in range(0, numel, chunk):
        e = min(numel, s + chunk)
        z = torch.arange(s, e, dtype=torch.int64, device=device) * _GOLD + base
        z = (z ^ _lsr(z, 30)) * _C1
        z = (z ^ _lsr(z, 27)) * _C2
        z = z ^ _lsr(z, 31)
        out[s:e] = _lsr(z, 40).to(torch.float32) * (1.0 / (1 << 24))
    return out


def uniform(shape, seed: int, bound: float, device="cpu") -> torch.Tensor:
    """U(-bound, bound), fp32."""
    n = int(math.prod(shape))
    return ((hash_u01(n, seed, device) * 2.0 - 1.0) * bound).reshape(shape)


def normal_like(shape, seed: int, std: float, device="cpu") -> torch.Tensor:
    """Zero-mean, given std (Irwin-Hall of 4 uniforms: bell-shaped, bounded, portable)."""
    n = int(math.prod(shape))
    acc = torch.zeros(n, dtype=torch.float32, device=device)
    for j in range(4):
        acc += hash_u01(n, seed * 4 + j + 0x5151, device)
    # sum of 4 U(0,1): mean 2, var 4/12
    return ((acc - 2.0) * (std / math.sqrt(4.0 / 12.0))).reshape(shape)


def _key_seed(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x01000193)) & 0x7FFFFFFF


def cross_attn_blocks(depth: int):
    """f_lite/model.py:464 -- cross attention in blocks idx % 4 == 0 or idx < 8."""
    return [i for i in range(depth) if (i % 4 == 0 or i < 8)]


def param_shapes(cfg: dict) -> dict:
    """State-dict keys and shapes of the reference DiT."""
    d = cfg["hidden_size"]
    c = cfg["in_channels"]
    p = cfg["patch_size"]
    ci = cfg["cross_attn_input_size"]
    inter = int(d * cfg.get("mlp_ratio", 4.0))
    bias = cfg.get("train_bias_and_rms", True)
    shapes = {
        "context_proj.weight": (d, ci), "context_proj.bias": (d,),
        "context_norm.weight": (d,),
        "patch_embed.patch_proj.weight": (d, c, p, p), "patch_embed.patch_proj.bias": (d,),
        "register_tokens": (1, 16, d),
        "time_embed.0.weight": (4 * d, d), "time_embed.0.bias": (4 * d,),
        "time_embed.2.weight": (d, 4 * d), "time_embed.2.bias": (d,),
        "adaLN_modulation.1.weight": (9 * d, d), "adaLN_modulation.1.bias": (9 * d,),
    }
    xs = set(cross_attn_blocks(cfg["depth"]))
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        shapes[b + "norm1.weight"] = (d,)
        shapes[b + "self_attn.qkv.weight"] = (3 * d, d)
        if bias:
            shapes[b + "self_attn.qkv.bias"] = (3 * d,)
        shapes[b + "self_attn.proj.weight"] = (d, d)
        if i in xs:
            shapes[b + "norm2.weight"] = (d,)
            shapes[b + "cross_attn.q.weight"] = (d, d)
            shapes[b + "cross_attn.context_kv.weight"] = (2 * d, d)
            if bias:
                shapes[b + "cross_attn.q.bias"] = (d,)
                shapes[b + "cross_attn.context_kv.bias"] = (2 * d,)
            shapes[b + "cross_attn.proj.weight"] = (d, d)
        shapes[b + "norm3.weight"] = (d,)
        shapes[b + "mlp.gate_proj.weight"] = (inter, d)
        shapes[b + "mlp.up_proj.weight"] = (inter, d)
        shapes[b + "mlp.down_proj.weight"] = (d, inter)
    shapes["final_modulation.1.weight"] = (2 * d, d)
    shapes["final_modulation.1.bias"] = (2 * d,)
    if bias:
        shapes["final_norm.weight"] = (d,)
    shapes["final_proj.weight"] = (p * p * c, d)
    shapes["final_proj.bias"] = (p * p * c,)
    return shapes


_ZERO_INIT = ("adaLN_modulation.1.", "final_modulation.1.", "final_proj.")


def make_state_dict(cfg: dict, seed: int = 0, device="cpu", dtype=torch.float32) -> dict:
    """Seeded 'de-zeroed default init' for every reference state-dict key."""
    sd = {}
    for name, shape in param_shapes(cfg).items():
        s = _key_seed(name, seed)
        if name == "register_tokens":
            t = normal_like(shape, s, 0.02, device)            # f_lite/model.py:509
        elif name.startswith(_ZERO_INIT):
            t = normal_like(shape, s, 0.02, device)            # de-zeroed (D10)
        elif name.endswith("norm.weight") or ".norm" in name:
            t = 1.0 + uniform(shape, s, 0.1, device)           # ones + jitter so the weight matters
        elif name.endswith(".weight"):
            fan_in = int(math.prod(shape[1:]))
            t = uniform(shape, s, 1.0 / math.sqrt(fan_in), device)
        else:  # bias of a default-init Linear/Conv: bound 1/sqrt(fan_in of its weight)
            wshape = param_shapes(cfg)[name[:-4] + "weight"]
            fan_in = int(math.prod(wshape[1:]))
            t = uniform(shape, s, 1.0 / math.sqrt(fan_in), device)
        sd[name] = t.to(dtype)
    return sd


def make_inputs(cfg: dict, batch: int, height: int, width: int, ctx_len: int,
                valid_len=None, seed: int = 1234, device="cpu", dtype=torch.float32):
    """Latents, CFG-ordered context [negative(zeros) ; positive], mask, as the samplers build
    them (f_lite/pipeline.py:160-161,264-268; f_lite/train.py:561-562).

    Returns x (batch,C,h,w), context (2*batch, ctx_len, ctx_in), mask (2*batch, ctx_len).
    ``valid_len`` (list of per-prompt valid token counts) exercises the varlen path: the
    positive rows get a prefix mask, the negative rows stay all-ones like train.py:562.
    """
    c = cfg["in_channels"]
    ci = cfg["cross_attn_input_size"]
    x = normal_like((batch, c, height // 8, width // 8), _key_seed("latents", seed), 1.0, device)
    pos = normal_like((batch, ctx_len, ci), _key_seed("context", seed), 1.0, device)
    neg = torch.zeros_like(pos)
    mask_pos = torch.ones(batch, ctx_len, device=device)
    if valid_len is not None:
        for b, n in enumerate(valid_len):
            mask_pos[b, n:] = 0
    mask = torch.cat([torch.ones_like(mask_pos), mask_pos], 0)
    context = torch.cat([neg, pos], 0)
    return x.to(dtype), context.to(dtype), mask.to(dtype)


TINY = dict(in_channels=16, patch_size=2, hidden_size=512, depth=4, num_heads=2, mlp_ratio=4.0,
            cross_attn_input_size=4096, train_bias_and_rms=True, use_rope=True,
            gradient_checkpoint=False, dynamic_softmax_temperature=False, rope_base=10000)
# "10B architecture" as instantiated by the mounted model.py: 6.84 B params.
ARCH_10B = dict(TINY, hidden_size=3072, depth=40, num_heads=12)
# ASSUMED "7B architecture" (SURVEY.md D7): not defined anywhere in the reference.
ARCH_7B = dict(TINY, hidden_size=3072, depth=28, num_heads=12)
