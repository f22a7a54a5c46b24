"""torch-tensor front end of the C ABI (``include/flite_b200.h``).

PyTorch is used for device memory and streams only; every function here launches a hand-written
sm_100a kernel from ``libflite_b200.so`` on torch's current CUDA stream.  No fallbacks.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_GATED_RES, EPI_QKV_ROPE, EPI_STORE, EPI_SWIGLU, GEMM_AUTO)

BF16 = torch.bfloat16

# number of libflite_b200 kernel launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = [0]
# optional hook called as PROFILE_HOOK(name, phase) with phase "begin"/"end" around selected ops (bench.py
# uses it to bracket the dominant GEMM with CUDA events on the launching stream)
PROFILE_HOOK = None


# Per-op device timing (tools/mgpu_check.py --trace, tools/profile_step.py): when TRACE is a list every public op
# below appends (label, start_event, end_event) recorded on the launching stream; trace_report() aggregates them.
TRACE = None
# FLITE_DEBUG_SYNC=1 (or ops.DEBUG_SYNC = True around a region): synchronise after every op and raise naming the op
# after which a device fault surfaced -- how the layout-dependent TMA fault of DESIGN.md section 7 was located.  Not
# usable under CUDA-graph capture (synchronising is illegal there).
DEBUG_SYNC = bool(int(os.environ.get("FLITE_DEBUG_SYNC", "0")))


def _traced(label_fn):
    def deco(fn):
        def wrapper(*args, **kwargs):
            tr = TRACE
            if DEBUG_SYNC:
                r = fn(*args, **kwargs)
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    raise _lib.FliteError(f"device fault surfaced after {label_fn(*args, **kwargs)}: {e}") from e
                return r
            if tr is None:
                return fn(*args, **kwargs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kwargs)
            e1.record()
            tr.append((label_fn(*args, **kwargs), e0, e1))
            return r
        wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
        wrapper.__wrapped__ = fn
        return wrapper
    return deco


def trace_report():
    """{label: (launches, total_ms)} of the ops recorded since TRACE was set; synchronises."""
    torch.cuda.synchronize()
    agg = {}
    for label, e0, e1 in TRACE or []:
        n, ms = agg.get(label, (0, 0.0))
        agg[label] = (n + 1, ms + e0.elapsed_time(e1))
    return agg


_EPI_NAMES = {EPI_STORE: "store", EPI_GATED_RES: "gated_res", EPI_SWIGLU: "swiglu", EPI_QKV_ROPE: "qkv_rope"}


def get_tuning(key: int) -> int:
    return _lib.load().flite_get_tuning(key)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=BF16, rows_contig=True):
    if not t.is_cuda:
        raise _lib.FliteError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise _lib.FliteError(f"{name}: expected {dtype}, got {t.dtype}")
    if rows_contig and t.dim() >= 1 and t.stride(-1) != 1:
        raise _lib.FliteError(f"{name}: innermost dimension must be contiguous")


@_traced(lambda acc, *a, **k: f"cfg_euler {acc.numel()}")
def cfg_euler(acc: torch.Tensor, v_uncond: Optional[torch.Tensor], v_cond: torch.Tensor, guidance: float,
              dt: float, lat_out: torch.Tensor, do_cfg: bool = True) -> None:
    """acc += dt * (u + g (c - u)); lat_out = bf16(acc).  f_lite/pipeline.py:290,296-297."""
    lib = _lib.load()
    _chk(v_cond, "v_cond")
    _chk(lat_out, "lat_out")
    if do_cfg:
        _chk(v_uncond, "v_uncond")
    if acc.dtype not in (BF16, torch.float32):
        raise _lib.FliteError("acc must be bf16 or fp32")
    for t in (acc, v_cond, lat_out) + ((v_uncond,) if do_cfg else ()):
        if not t.is_contiguous():
            raise _lib.FliteError("cfg_euler operands must be contiguous")
    _lib.check(lib.flite_cfg_euler(acc.data_ptr(), int(acc.dtype == torch.float32), _ptr(v_uncond) if do_cfg else None,
                                   v_cond.data_ptr(), float(guidance), float(dt), int(do_cfg), lat_out.data_ptr(),
                                   acc.numel(), _stream()), "cfg_euler")
    LAUNCHES[0] += 1


_APG_WS = {}


@_traced(lambda acc, *a, **k: f"apg_euler {acc.numel()}")
def apg_euler(acc: torch.Tensor, v_uncond: torch.Tensor, v_cond: torch.Tensor, guidance: float, dt: float,
              orthogonal_threshold: float, lat_out: torch.Tensor) -> None:
    """Augmented Parallel Guidance combine + Euler update in three stream-ordered launches, no host sync.
    f_lite/pipeline.py:276-287,296-297."""
    lib = _lib.load()
    for t, n in ((v_uncond, "v_uncond"), (v_cond, "v_cond"), (lat_out, "lat_out")):
        _chk(t, n)
    if acc.dtype not in (BF16, torch.float32):
        raise _lib.FliteError("acc must be bf16 or fp32")
    for t in (acc, v_uncond, v_cond, lat_out):
        if not t.is_contiguous():
            raise _lib.FliteError("apg_euler operands must be contiguous")
    ws = _APG_WS.get(acc.device)
    if ws is None:
        ws = torch.empty(lib.flite_apg_workspace_bytes() // 8, dtype=torch.float64, device=acc.device)
        _APG_WS[acc.device] = ws
    _lib.check(lib.flite_apg_euler(acc.data_ptr(), int(acc.dtype == torch.float32), v_uncond.data_ptr(),
                                   v_cond.data_ptr(), float(guidance), float(dt), float(orthogonal_threshold),
                                   lat_out.data_ptr(), acc.numel(), ws.data_ptr(), _stream()), "apg_euler")
    LAUNCHES[0] += 3


def latent_unscale(latents: torch.Tensor, scaling_factor: float, shift_factor: float,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """latents / scaling_factor + shift_factor in bf16 (f_lite/pipeline.py:304)."""
    lib = _lib.load()
    _chk(latents, "latents")
    if not latents.is_contiguous():
        raise _lib.FliteError("latents must be contiguous")
    if out is None:
        out = torch.empty_like(latents)
    _lib.check(lib.flite_latent_unscale(latents.data_ptr(), out.data_ptr(), float(scaling_factor), float(shift_factor),
                                        latents.numel(), _stream()), "latent_unscale")
    LAUNCHES[0] += 1
    return out


def image_to_uint8(decoded: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """decoded [B, C, H, W] bf16 or fp32 in [-1, 1] -> uint8 [B, H, W, C] (f_lite/pipeline.py:324-327)."""
    lib = _lib.load()
    if not decoded.is_cuda or decoded.dtype not in (BF16, torch.float32) or not decoded.is_contiguous():
        raise _lib.FliteError("image_to_uint8: expected a contiguous CUDA bf16/fp32 [B, C, H, W] tensor")
    B, C, H, W = decoded.shape
    if out is None:
        out = torch.empty((B, H, W, C), dtype=torch.uint8, device=decoded.device)
    _lib.check(lib.flite_image_to_uint8(decoded.data_ptr(), int(decoded.dtype == torch.float32), out.data_ptr(), B, C,
                                        H, W, _stream()), "image_to_uint8")
    LAUNCHES[0] += 1
    return out


_GN_WS = {}


def groupnorm_silu(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, eps: float = 1e-6,
                   silu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+ SiLU) of a channels-last [N, C, H, W] bf16 activation in two launches (flite_groupnorm_silu_nhwc):
    the VAE decoder's norm -> nonlinearity pairs (diffusers ResnetBlock2D / conv_norm_out, f_lite/pipeline.py:299-307)."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == BF16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        raise _lib.FliteError("groupnorm_silu: expected a channels-last CUDA bf16 [N, C, H, W] tensor")
    N, C, H, W = x.shape
    if out is None:
        out = torch.empty_like(x)           # keeps the channels-last strides
    elif out.shape != x.shape or out.stride() != x.stride() or out.dtype != BF16:
        raise _lib.FliteError("groupnorm_silu: out must match x (shape, channels-last strides, bf16)")
    w = weight if weight.dtype == BF16 else weight.to(BF16)
    b = bias if bias.dtype == BF16 else bias.to(BF16)
    splits = max(1, min(512, (8 * torch.cuda.get_device_properties(x.device).multi_processor_count + N - 1) // N,
                        (H * W + 255) // 256))
    key = (x.device, N, groups, splits)
    ws = _GN_WS.get(key)
    if ws is None:
        _GN_WS.clear()
        ws = _GN_WS[key] = torch.empty(int(lib.flite_groupnorm_partials_bytes(N, groups, splits)) // 4, dtype=torch.float32,
                                       device=x.device)
    _lib.check(lib.flite_groupnorm_silu_nhwc(x.data_ptr(), out.data_ptr(), w.contiguous().data_ptr(), b.contiguous().data_ptr(),
                                             N, H * W, C, groups, float(eps), int(silu), ws.data_ptr(), splits, _stream()),
               "groupnorm_silu")
    LAUNCHES[0] += 2
    return out


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    """Nearest-neighbour 2x upsampling of a channels-last [N, C, H, W] bf16 activation (diffusers Upsample2D)."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == BF16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        raise _lib.FliteError("upsample_nearest2x: expected a channels-last CUDA bf16 [N, C, H, W] tensor")
    N, C, H, W = x.shape
    out = torch.empty((N, C, 2 * H, 2 * W), dtype=BF16, device=x.device, memory_format=torch.channels_last)
    _lib.check(lib.flite_upsample_nearest2x_nhwc(x.data_ptr(), out.data_ptr(), N, H, W, C, _stream()), "upsample_nearest2x")
    LAUNCHES[0] += 1
    return out


def bias_residual_add_(y: torch.Tensor, bias: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place on a channels-last [N, C, H, W] bf16 conv output: y = bf16(y + bias); y = bf16(residual + y) if given."""
    lib = _lib.load()
    cl = torch.channels_last
    if not (y.is_cuda and y.dtype == BF16 and y.dim() == 4 and y.is_contiguous(memory_format=cl)):
        raise _lib.FliteError("bias_residual_add_: expected a channels-last CUDA bf16 [N, C, H, W] tensor")
    if residual is not None and not (residual.shape == y.shape and residual.dtype == BF16
                                     and residual.is_contiguous(memory_format=cl)):
        raise _lib.FliteError("bias_residual_add_: residual must match y (shape, channels-last, bf16)")
    N, C, H, W = y.shape
    b = bias if bias.dtype == BF16 else bias.to(BF16)
    _lib.check(lib.flite_bias_residual_add_nhwc(y.data_ptr(), b.contiguous().data_ptr(), _ptr(residual), N * H * W, C,
                                                _stream()), "bias_residual_add")
    LAUNCHES[0] += 1
    return y


@_traced(lambda x, *a, **k: f"rmsnorm_modulate {tuple(x.shape)}")
def rmsnorm_modulate(x: torch.Tensor, weight: Optional[torch.Tensor], weight_mode: int,
                     scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                     rows_per_sample: int = 0, eps: float = 1e-6, out: Optional[torch.Tensor] = None):
    """x [rows, d]; scale/shift are [B, d] views into the modulation matrix (same row stride)."""
    lib = _lib.load()
    _chk(x, "x")
    rows, d = x.shape
    if out is None:
        out = torch.empty((rows, d), dtype=BF16, device=x.device)
    ld_mod = 0
    if scale is not None:
        _chk(scale, "scale")
        _chk(shift, "shift")
        ld_mod = scale.stride(0)
        if shift.stride(0) != ld_mod:
            raise _lib.FliteError("scale and shift must share a row stride")
    _lib.check(lib.flite_rmsnorm_modulate(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), _ptr(weight),
                                          weight_mode, _ptr(scale), _ptr(shift), ld_mod, rows_per_sample, rows, d,
                                          eps, _stream()), "rmsnorm_modulate")
    LAUNCHES[0] += 1
    return out


def rope_qknorm_(buf: torch.Tensor, n_slots: int, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor],
                 rows_per_sample: int = 0, eps: float = 1e-6) -> None:
    lib = _lib.load()
    _chk(buf, "buf")
    if cos is not None:
        _chk(cos, "cos")
        _chk(sin, "sin")
    _lib.check(lib.flite_rope_qknorm(buf.data_ptr(), buf.stride(0), buf.shape[0], n_slots, _ptr(cos), _ptr(sin),
                                     rows_per_sample, eps, _stream()), "rope_qknorm")
    LAUNCHES[0] += 1


_PATCH_WS = {}


@_traced(lambda x, *a, **k: f"patch_embed {tuple(x.shape)}")
def patch_embed(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, reg_tokens: torch.Tensor, patch: int,
                out: Optional[torch.Tensor] = None, tok_offset: int = 0, tok_count: int = 0) -> torch.Tensor:
    """Conv2d(k = s = patch) + register tokens (model.py:324-328,535) -> token rows [B * tok_count, d].

    C*P*P % 64 == 0 (F Lite: 16 * 2 * 2): an im2col gather (flite_patch_gather) followed by the tcgen05 GEMM (+bias)
    straight into the image rows of every sample; otherwise the CUDA-core patch_embed kernel."""
    lib = _lib.load()
    for t, n in ((x, "x"), (weight, "weight"), (bias, "bias"), (reg_tokens, "register_tokens")):
        _chk(t, n)
        if not t.is_contiguous():
            raise _lib.FliteError(f"{n} must be contiguous")
    B, C, H, W = x.shape
    d = weight.shape[0]
    n_reg = reg_tokens.shape[-2]
    L_full = n_reg + (H // patch) * (W // patch)
    if tok_count <= 0:
        tok_offset, tok_count = 0, L_full
    rows = B * tok_count
    if out is None:
        out = torch.empty((rows, d), dtype=BF16, device=x.device)
    kdim = C * patch * patch
    if kdim % 64 == 0 and lib.flite_get_tuning(11) == 0:
        n_reg_local = max(0, min(n_reg, tok_offset + tok_count) - tok_offset)
        n_img = tok_count - n_reg_local
        key = (x.device, B * n_img, kdim)
        A = _PATCH_WS.get(key)
        if A is None:
            A = torch.empty((max(B * n_img, 1), kdim), dtype=BF16, device=x.device)
            _PATCH_WS[key] = A
        _lib.check(lib.flite_patch_gather(x.data_ptr(), reg_tokens.data_ptr(), A.data_ptr(), out.data_ptr(), B, C, H, W,
                                          patch, d, n_reg, tok_offset, tok_count, _stream()), "patch_gather")
        LAUNCHES[0] += 1
        if n_img > 0:
            w2 = weight.view(d, kdim)
            for b in range(B):
                gemm.__wrapped__(A[b * n_img:(b + 1) * n_img], w2, bias,
                                 out=out[b * tok_count + n_reg_local:(b + 1) * tok_count])
        return out
    _lib.check(lib.flite_patch_embed(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), reg_tokens.data_ptr(),
                                     out.data_ptr(), B, C, H, W, patch, d, n_reg, tok_offset, tok_count, _stream()),
               "patch_embed")
    LAUNCHES[0] += 1
    return out


@_traced(lambda *a, **k: "timestep_embed")
def timestep_embed(t_f32: torch.Tensor, t_is_bf16: bool, freqs: torch.Tensor, d: int,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _chk(t_f32, "t", torch.float32)
    _chk(freqs, "freqs", torch.float32)
    B = t_f32.numel()
    if out is None:
        out = torch.empty((B, d), dtype=BF16, device=t_f32.device)
    _lib.check(lib.flite_timestep_embed(t_f32.data_ptr(), int(t_is_bf16), freqs.data_ptr(), out.data_ptr(), B, d,
                                        _stream()), "timestep_embed")
    LAUNCHES[0] += 1
    return out


@_traced(lambda tok, *a, **k: f"unpatchify {tuple(tok.shape)}")
def unpatchify(tok: torch.Tensor, B: int, C: int, H: int, W: int, patch: int, n_reg: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    _chk(tok, "tok")
    if out is None:
        out = torch.empty((B, C, H, W), dtype=BF16, device=tok.device)
    _lib.check(lib.flite_unpatchify(tok.data_ptr(), tok.stride(0), out.data_ptr(), B, C, H, W, patch, n_reg,
                                    _stream()), "unpatchify")
    LAUNCHES[0] += 1
    return out


@_traced(lambda src, *a, **k: f"pack_context {tuple(src.shape)}")
def pack_context(src: torch.Tensor, mask_f32: torch.Tensor):
    """src [B, Lc, d] bf16, mask [B, Lc] fp32 -> (packed [B*Lc, d] zero-padded, cu_seqlens int32 [B+1])."""
    lib = _lib.load()
    _chk(src, "src")
    _chk(mask_f32, "mask", torch.float32)
    B, Lc, d = src.shape
    src2 = src.reshape(B * Lc, d)
    dst = torch.zeros((B * Lc, d), dtype=BF16, device=src.device)
    pos = torch.empty(B * Lc, dtype=torch.int32, device=src.device)
    seqlens = torch.empty(B, dtype=torch.int32, device=src.device)
    cu = torch.empty(B + 1, dtype=torch.int32, device=src.device)
    _lib.check(lib.flite_pack_context(src2.data_ptr(), src2.stride(0), dst.data_ptr(), dst.stride(0),
                                      mask_f32.contiguous().data_ptr(), B, Lc, d, pos.data_ptr(), seqlens.data_ptr(),
                                      cu.data_ptr(), _stream()), "pack_context")
    LAUNCHES[0] += 3
    return dst, cu


@_traced(lambda a, w, *r, **k: f"gemm {_EPI_NAMES[k.get('epilogue', EPI_STORE)]} {a.shape[0]}x{w.shape[0]}x{a.shape[1]}"
         + (" sp-scatter" if k.get("sp_ranks", 0) else ""))
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = 0,
         epilogue: int = EPI_STORE, resid: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
         rows_per_sample: int = 0, rope_cos: Optional[torch.Tensor] = None, rope_sin: Optional[torch.Tensor] = None,
         qk_cols: int = 0, eps: float = 1e-6, variant: int = GEMM_AUTO, out: Optional[torch.Tensor] = None,
         sp_ranks: int = 0, sp_heads_per_rank: int = 0):
    """out = epilogue(a @ w.T); a [M, K], w [N, K] (nn.Linear layout), bf16."""
    lib = _lib.load()
    _chk(a, "a")
    _chk(w, "w")
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == EPI_SWIGLU else N
    if out is None:
        if sp_ranks > 0:
            raise _lib.FliteError("gemm: the sequence-parallel head scatter needs an explicit output buffer")
        out = torch.empty((M, n_out), dtype=BF16, device=a.device)
    _chk(out, "out")
    for t, n in ((bias, "bias"), (resid, "resid"), (gate, "gate")):
        if t is not None:
            _chk(t, n)
    for t, n in ((rope_cos, "rope_cos"), (rope_sin, "rope_sin")):
        if t is not None:
            _chk(t, n)
            if not t.is_contiguous():
                raise _lib.FliteError(f"{n} must be contiguous [rows_per_sample, 128]")
    hook = PROFILE_HOOK
    if hook is not None:
        hook("gemm", "begin", (M, N, K, epilogue))
    _lib.check(lib.flite_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(),
                                   out.stride(0), M, N, K, _ptr(bias), act, epilogue, _ptr(resid),
                                   resid.stride(0) if resid is not None else 0, _ptr(gate),
                                   gate.stride(0) if gate is not None else 0, rows_per_sample, _ptr(rope_cos),
                                   _ptr(rope_sin), qk_cols, eps, sp_ranks, sp_heads_per_rank, variant, _stream()),
               "gemm_bf16")
    LAUNCHES[0] += 1
    if hook is not None:
        hook("gemm", "end", (M, N, K, epilogue))
    return out


@_traced(lambda q, k, v, cu_q, cu_k, num_heads, max_q, *r, **kw: f"attention Tq={q.shape[0]} Tk={k.shape[0]} H={num_heads}")
def attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor,
                     num_heads: int, max_q: int, softmax_scale: float, out: Optional[torch.Tensor] = None,
                     variant: int = 0):
    """q [Tq, >= H*256] / k, v [Tk, >= H*256] are (possibly strided column) views of projection buffers."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, n)
    _chk(cu_q, "cu_q", torch.int32)
    _chk(cu_k, "cu_k", torch.int32)
    Tq = q.shape[0]
    if out is None:
        out = torch.empty((Tq, num_heads * 256), dtype=BF16, device=q.device)
    B = cu_q.numel() - 1
    _lib.check(lib.flite_attention_varlen(q.data_ptr(), q.stride(0), Tq, 0, k.data_ptr(), k.stride(0), k.shape[0], 0,
                                          v.data_ptr(), v.stride(0), 0, out.data_ptr(), out.stride(0),
                                          cu_q.data_ptr(), cu_k.data_ptr(), B, num_heads, max_q,
                                          float(softmax_scale), variant, _stream()), "attention_varlen")
    LAUNCHES[0] += 1
    return out


_SK_WS = {}


@_traced(lambda q, k, v, cu_q, cu_k, num_heads, q_len, k_len, *r, **kw: f"attention Tq={q.shape[0]} Tk={k.shape[0]} H={num_heads}")
def attention_streamk(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor,
                      num_heads: int, q_len: int, k_len: int, softmax_scale: float, out: Optional[torch.Tensor] = None):
    """Self-attention for uniform sequence lengths as one persistent stream-K wave (flite_attention_streamk): same
    arguments as ``attention_varlen`` plus the per-sequence lengths the host knows."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, n)
    _chk(cu_q, "cu_q", torch.int32)
    _chk(cu_k, "cu_k", torch.int32)
    Tq = q.shape[0]
    if out is None:
        out = torch.empty((Tq, num_heads * 256), dtype=BF16, device=q.device)
    ws = _sk_workspace(q.device)
    B = cu_q.numel() - 1
    _lib.check(lib.flite_attention_streamk(q.data_ptr(), q.stride(0), Tq, 0, k.data_ptr(), k.stride(0), k.shape[0], 0,
                                           v.data_ptr(), v.stride(0), 0, out.data_ptr(), out.stride(0), cu_q.data_ptr(),
                                           cu_k.data_ptr(), B, num_heads, int(q_len), int(k_len), float(softmax_scale),
                                           ws.data_ptr(), ws.numel(), _stream()), "attention_streamk")
    LAUNCHES[0] += 1
    return out


def _sk_workspace(device):
    lib = _lib.load()
    ws = _SK_WS.get(device)
    if ws is None:      # flags + one partial slot per cluster; zero-filled once, the kernel leaves the flags zeroed
        ws = torch.zeros(int(lib.flite_attention_streamk_workspace_bytes()), dtype=torch.uint8, device=device)
        _SK_WS[device] = ws
    return ws


@_traced(lambda q, k, v, cu_q, cu_k, num_heads, *r, **kw: f"attention+p2p Tq={q.shape[0]} Tk={k.shape[0]} H={num_heads}")
def attention_streamk_p2p(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor,
                          num_heads: int, q_len: int, k_len: int, softmax_scale: float, peer_out, n_peers: int,
                          tokens_per_rank: int, head0: int, ldo: int) -> None:
    """``attention_streamk`` with the fused return all-to-all of ``attention_varlen_p2p``."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, n)
    _chk(cu_q, "cu_q", torch.int32)
    _chk(cu_k, "cu_k", torch.int32)
    ws = _sk_workspace(q.device)
    B = cu_q.numel() - 1
    _lib.check(lib.flite_attention_streamk_p2p(q.data_ptr(), q.stride(0), q.shape[0], 0, k.data_ptr(), k.stride(0),
                                               k.shape[0], 0, v.data_ptr(), v.stride(0), 0, peer_out, n_peers,
                                               tokens_per_rank, head0, ldo, cu_q.data_ptr(), cu_k.data_ptr(), B, num_heads,
                                               int(q_len), int(k_len), float(softmax_scale), ws.data_ptr(), ws.numel(),
                                               _stream()), "attention_streamk_p2p")
    LAUNCHES[0] += 1


@_traced(lambda a, w, *r, **k: f"gemm qkv_rope+p2p {a.shape[0]}x{w.shape[0]}x{a.shape[1]}")
def gemm_qkv_p2p(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], rope_cos: torch.Tensor,
                 rope_sin: torch.Tensor, tokens_per_sample: int, sp_ranks: int, sp_heads_per_rank: int, sp_rank: int,
                 seq_len: int, peer_recv, eps: float = 1e-6, variant: int = GEMM_AUTO) -> None:
    """Fused QKV projection + head all-to-all: ``a @ w.T`` (+bias, RoPE, QK-norm, model.py:162-183) with every head
    stored straight into the receive buffer of the rank that owns it over NVLink peer memory.  ``peer_recv`` is a
    ctypes array of the ranks' receive-buffer addresses (``peer.SymmetricBuffer.table``)."""
    lib = _lib.load()
    _chk(a, "a")
    _chk(w, "w")
    for t, n in ((bias, "bias"), (rope_cos, "rope_cos"), (rope_sin, "rope_sin")):
        if t is not None:
            _chk(t, n)
    M, K = a.shape
    if w.shape[0] != 3 * sp_ranks * sp_heads_per_rank * 256:
        raise _lib.FliteError("gemm_qkv_p2p: weight rows must be 3 * heads * 256")
    _lib.check(lib.flite_gemm_qkv_p2p(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, K, _ptr(bias),
                                      tokens_per_sample, _ptr(rope_cos), _ptr(rope_sin), eps, sp_ranks,
                                      sp_heads_per_rank, sp_rank, seq_len, peer_recv, variant, _stream()),
               "gemm_qkv_p2p")
    LAUNCHES[0] += 1


@_traced(lambda q, k, v, cu_q, cu_k, num_heads, *r, **kw: f"attention+p2p Tq={q.shape[0]} Tk={k.shape[0]} H={num_heads}")
def attention_varlen_p2p(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor,
                         num_heads: int, max_q: int, softmax_scale: float, peer_out, n_peers: int,
                         tokens_per_rank: int, head0: int, ldo: int, variant: int = 0) -> None:
    """Fused attention + return all-to-all: query row l of this rank's ``num_heads`` heads is stored into the buffer of
    the rank that owns token l (``peer_out[l // tokens_per_rank]``, row ``b*tokens_per_rank + l % tokens_per_rank``,
    columns ``(head0 + h) * 256``, row stride ``ldo``)."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, n)
    _chk(cu_q, "cu_q", torch.int32)
    _chk(cu_k, "cu_k", torch.int32)
    B = cu_q.numel() - 1
    _lib.check(lib.flite_attention_varlen_p2p(q.data_ptr(), q.stride(0), q.shape[0], 0, k.data_ptr(), k.stride(0),
                                              k.shape[0], 0, v.data_ptr(), v.stride(0), 0, peer_out, n_peers,
                                              tokens_per_rank, head0, ldo, cu_q.data_ptr(), cu_k.data_ptr(), B,
                                              num_heads, max_q, float(softmax_scale), variant, _stream()),
               "attention_varlen_p2p")
    LAUNCHES[0] += 1


@_traced(lambda src, *r, **k: f"permute_021 {tuple(src.shape)}")
def permute_021(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n0, n1, n2] -> [n1, n0, n2] (bf16 contiguous, n2 % 8 == 0)."""
    lib = _lib.load()
    _chk(src, "src")
    if not src.is_contiguous() or src.dim() != 3:
        raise _lib.FliteError("permute_021 needs a contiguous 3-D tensor")
    n0, n1, n2 = src.shape
    if out is None:
        out = torch.empty((n1, n0, n2), dtype=BF16, device=src.device)
    _lib.check(lib.flite_permute_021(src.data_ptr(), out.data_ptr(), n0, n1, n2, _stream()), "permute_021")
    LAUNCHES[0] += 1
    return out


def poison_on_abort(buf: torch.Tensor) -> None:
    """Stream-ordered fail-closed guard: NaN-fill ``buf`` if a kernel-side wait of this process has timed out."""
    lib = _lib.load()
    _chk(buf, "buf")
    if not buf.is_contiguous():
        raise _lib.FliteError("poison_on_abort: buffer must be contiguous")
    _lib.check(lib.flite_poison_on_abort(buf.data_ptr(), buf.numel(), _stream()), "poison_on_abort")
    LAUNCHES[0] += 1


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor) -> torch.Tensor:
    """[inter, d] x2 -> [2*inter, d] with rows interleaved in groups of 64 ([g 64 | u 64] per 128 rows) so that
    one accumulator tile holds matching gate / up columns for the EPI_SWIGLU epilogue."""
    inter, d = gate_w.shape
    assert inter % 64 == 0
    return torch.stack([gate_w.view(inter // 64, 64, d), up_w.view(inter // 64, 64, d)], dim=1).reshape(2 * inter, d).contiguous()
