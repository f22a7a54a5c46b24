"""FLUX-style ``AutoencoderKL`` decoder for the tail of ``FLitePipeline.__call__``.

The reference decodes the final latents with diffusers' ``AutoencoderKL.decode`` (``f_lite/pipeline.py:299-307``;
16 latent channels, block_out_channels (128, 256, 512, 512), layers_per_block 2, 32 norm groups, mid-block
self-attention, no post-quant conv; ``scaling_factor`` 0.3611 / ``shift_factor`` 0.1159 in ``vae.config``).  diffusers
is a third-party dependency of the reference (unpinned, ``requirements.txt:1``) and is not installed here, so the
decoder is provided by this module with **diffusers' parameter names** (``decoder.conv_in.weight``,
``decoder.up_blocks.0.resnets.0.conv1.weight`` ...) so that a FLUX VAE checkpoint loads with ``load_state_dict``
(names restated from diffusers' ``models/autoencoders/vae.py``; encoder / quant-conv keys are ignored with
``strict=False``).

This is **not** part of the denoise hot path.  Profiled at 1024^2 (``tools/vae_probe.py``): the convolutions are the
cheap part of a decode (cuDNN's sm_100 implicit-GEMM kernels, ~1080 TFLOP/s, 10 % of the time) -- 85 % went to torch's
GroupNorm row-moments kernel and unvectorised elementwise kernels on the channels-last activations.  On a CUDA bf16
channels-last tensor the norm -> SiLU pairs therefore run as ``flite_groupnorm_silu_nhwc`` (``csrc/groupnorm.cuh``, two
HBM-bound launches); convolutions and the single mid-block attention stay on cuDNN / SDPA through torch (library code);
on the CPU (``tests/test_vae_cpu.py``, the oracle comparison) everything is torch.  Around the decoder the pipeline tail
is ``flite_latent_unscale`` before and ``flite_image_to_uint8`` after (``ops.py``).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn


# False = run the norm / bias / skip-connection passes through torch as well (A/B and parity tests)
USE_NATIVE_KERNELS = True


def _gn(x: torch.Tensor, norm: nn.GroupNorm, silu: bool) -> torch.Tensor:
    """GroupNorm (+ SiLU): the sm_100a kernel pair for channels-last CUDA bf16 activations, torch otherwise."""
    if (USE_NATIVE_KERNELS and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] % 8 == 0 and (x.shape[1] // norm.num_groups) % 4 == 0 and norm.num_groups <= 64):
        from . import ops
        return ops.groupnorm_silu(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu=silu)
    y = norm(x)
    return F.silu(y) if silu else y


def _native(x: torch.Tensor, channels: int) -> bool:
    return (USE_NATIVE_KERNELS and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and channels % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


def _conv(x: torch.Tensor, conv: nn.Conv2d, residual: torch.Tensor | None = None) -> torch.Tensor:
    """conv(x) [+ residual].  On the CUDA bf16 channels-last path the convolution itself is cuDNN's (library code, run
    without its bias) and the bias + skip connection are one vectorised in-place pass (flite_bias_residual_add_nhwc)
    with torch's rounding points; anywhere else plain torch."""
    if conv.bias is not None and _native(x, conv.out_channels):
        y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding)
        if y.is_contiguous(memory_format=torch.channels_last) and (
                residual is None or (residual.shape == y.shape and residual.is_contiguous(memory_format=torch.channels_last))):
            from . import ops
            return ops.bias_residual_add_(y, conv.bias, residual)
        y = y + conv.bias.view(1, -1, 1, 1)
        return y if residual is None else residual + y
    y = conv(x)
    return y if residual is None else residual + y


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = _conv(_gn(x, self.norm1, True), self.conv1)
        skip = x if self.conv_shortcut is None else _conv(x, self.conv_shortcut)
        return _conv(_gn(h, self.norm2, True), self.conv2, residual=skip)


class Attention(nn.Module):
    """Single-head spatial self-attention of the mid block (diffusers ``Attention`` with ``group_norm``)."""

    def __init__(self, c: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = _gn(x, self.group_norm, False).flatten(2).transpose(1, 2)
        a = F.scaled_dot_product_attention(self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None])[:, 0]
        return x + self.to_out[0](a).transpose(1, 2).reshape(b, c, h, w)


class _MidBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Attention(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Upsample(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        if _native(x, x.shape[1]):
            from . import ops
            return _conv(ops.upsample_nearest2x(x), self.conv)
        return _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), self.conv)


class _UpBlock(nn.Module):
    def __init__(self, cin: int, cout: int, n_layers: int, upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(n_layers)])
        self.upsamplers = nn.ModuleList([_Upsample(cout)]) if upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, latent_channels=16, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        ch = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(latent_channels, ch[0], 3, padding=1)
        self.mid_block = _MidBlock(ch[0])
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(_UpBlock(prev, c, layers_per_block + 1, upsample=i != len(ch) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(_conv(z, self.conv_in))
        for blk in self.up_blocks:
            h = blk(h)
        return self.conv_out(_gn(h, self.conv_norm_out, True))


class AutoencoderKL(nn.Module):
    """Decode-only stand-in for ``diffusers.AutoencoderKL`` with the surface ``FLitePipeline`` uses
    (pipeline.py:73-76,84-92,301-307): ``.config.scaling_factor / .shift_factor``, ``.dtype``, ``.decode(z).sample``,
    ``enable_slicing()`` (one image per decoder call) and ``enable_tiling()`` (accepted, no-op: 180 GB of HBM)."""

    def __init__(self, latent_channels=16, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 scaling_factor=0.3611, shift_factor=0.1159):
        super().__init__()
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.config = SimpleNamespace(scaling_factor=scaling_factor, shift_factor=shift_factor,
                                      latent_channels=latent_channels, block_out_channels=tuple(block_out_channels),
                                      layers_per_block=layers_per_block, use_post_quant_conv=False)
        self.use_slicing = False

    @property
    def dtype(self):
        return self.decoder.conv_in.weight.dtype

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def enable_tiling(self):
        pass

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        z = z.contiguous(memory_format=torch.channels_last)
        if self.use_slicing and z.shape[0] > 1:
            x = torch.cat([self.decoder(zi) for zi in z.split(1)])
        else:
            x = self.decoder(z)
        x = x.contiguous()
        return SimpleNamespace(sample=x) if return_dict else (x,)
