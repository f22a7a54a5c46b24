"""``torch.library.custom_op`` registration of the hot-path entry points ("loaded with ctypes, wrapped in
torch.library.custom_op for clean dispatch").

The model itself calls ``ops.*`` directly (one ctypes call per launch, nothing in between); these registrations exist so
that code living in a ``torch`` graph world -- ``torch.compile``d callers, ``torch.export``, FakeTensor shape propagation --
can treat the kernels as opaque operators: each op has a fake (meta) implementation that only produces shapes, and
declares which arguments it mutates.  CUDA only: like the rest of the package there is no CPU implementation.

    torch.ops.flite_b200.linear(a, w, bias)                                  # nn.Linear             model.py:436,448-454
    torch.ops.flite_b200.rmsnorm_modulate(x, weight, scale, shift, rows)     # LigerRMSNorm+modulate  model.py:283-284
    torch.ops.flite_b200.attention_varlen(q, k, v, cu_q, cu_k, H, max_q, s)  # flash_attn_varlen_func model.py:203-210
    torch.ops.flite_b200.cfg_euler_(acc, uncond, cond, g, dt, lat_out)       # pipeline.py:290,296-297 (in place)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

_NS = "flite_b200"


@torch.library.custom_op(f"{_NS}::linear", mutates_args=(), device_types="cuda")
def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return ops.gemm(a, w, bias)


@linear.register_fake
def _(a, w, bias=None):
    return a.new_empty((a.shape[0], w.shape[0]))


@torch.library.custom_op(f"{_NS}::rmsnorm_modulate", mutates_args=(), device_types="cuda")
def rmsnorm_modulate(x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                     rows_per_sample: int) -> torch.Tensor:
    return ops.rmsnorm_modulate(x, weight, 1, scale, shift, rows_per_sample=rows_per_sample)


@rmsnorm_modulate.register_fake
def _(x, weight, scale, shift, rows_per_sample):
    return torch.empty_like(x)


@torch.library.custom_op(f"{_NS}::attention_varlen", mutates_args=(), device_types="cuda")
def attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor,
                     num_heads: int, max_q: int, softmax_scale: float) -> torch.Tensor:
    return ops.attention_varlen(q, k, v, cu_q, cu_k, num_heads, max_q, softmax_scale)


@attention_varlen.register_fake
def _(q, k, v, cu_q, cu_k, num_heads, max_q, softmax_scale):
    return q.new_empty((q.shape[0], num_heads * 256))


@torch.library.custom_op(f"{_NS}::cfg_euler_", mutates_args=("acc", "lat_out"), device_types="cuda")
def cfg_euler_(acc: torch.Tensor, v_uncond: torch.Tensor, v_cond: torch.Tensor, guidance: float, dt: float,
               lat_out: torch.Tensor) -> None:
    ops.cfg_euler(acc, v_uncond, v_cond, guidance, dt, lat_out, do_cfg=True)


REGISTERED = ("linear", "rmsnorm_modulate", "attention_varlen", "cfg_euler_")
