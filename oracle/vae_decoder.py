"""TEST INFRASTRUCTURE -- FLUX-style ``AutoencoderKL`` decoder restated in torch (random init).

The reference decodes the final latents with diffusers' ``AutoencoderKL.decode`` (f_lite/pipeline.py:299-307,
f_lite/train.py:602-603; FLUX 16-channel VAE, scaling 0.3611 / shift 0.1159 from ``vae.config``).  diffusers is not
installed in this image and no checkpoint can be downloaded, so the decoder architecture
(``diffusers.models.autoencoders.vae.Decoder`` with block_out_channels (128, 256, 512, 512), layers_per_block 2,
norm_num_groups 32, mid-block self-attention; version unpinned by the reference, requirements.txt:1) is restated here
and randomly initialised.  It is used ONLY to evaluate the north_star image criterion -- PSNR of the decoded image
of the new path against the decoded image of the reference path, the same decoder applied to both trajectories
 -- never by the product path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

SCALING_FACTOR = 0.3611
SHIFT_FACTOR = 0.1159


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.shortcut is None else self.shortcut(x)) + h


class AttnBlock(nn.Module):
    def __init__(self, c, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.q, self.k, self.v, self.o = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.norm(x).flatten(2).transpose(1, 2)
        a = F.scaled_dot_product_attention(self.q(t)[:, None], self.k(t)[:, None], self.v(t)[:, None])[:, 0]
        return x + self.o(a).transpose(1, 2).reshape(b, c, h, w)


class Decoder(nn.Module):
    def __init__(self, latent_channels=16, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        ch = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(latent_channels, ch[0], 3, padding=1)
        self.mid = nn.Sequential(ResnetBlock(ch[0], ch[0]), AttnBlock(ch[0]), ResnetBlock(ch[0], ch[0]))
        ups = []
        prev = ch[0]
        for i, c in enumerate(ch):
            for _ in range(layers_per_block + 1):
                ups.append(ResnetBlock(prev, c))
                prev = c
            if i != len(ch) - 1:
                ups.append(nn.Upsample(scale_factor=2.0, mode="nearest"))
                ups.append(nn.Conv2d(c, c, 3, padding=1))
        self.up = nn.Sequential(*ups)
        self.norm_out = nn.GroupNorm(32, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], out_channels, 3, padding=1)

    def forward(self, z):
        h = self.up(self.mid(self.conv_in(z)))
        return self.conv_out(F.silu(self.norm_out(h)))


def make_decoder(seed: int = 0, device="cpu") -> Decoder:
    g = torch.Generator().manual_seed(seed)
    dec = Decoder()
    for p in dec.parameters():                      # seeded, device-independent init (weights; biases/norms default)
        if p.dim() > 1:
            fan_in = p.shape[1:].numel()
            p.data.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5)
    return dec.to(device).eval()


@torch.no_grad()
def decode_to_image(dec: Decoder, latents: torch.Tensor, gain: float = 1.0) -> torch.Tensor:
    """f_lite/pipeline.py:299-307,324-326: latents / scaling + shift -> decode -> (x/2 + 0.5).clamp(0, 1).
    ``gain`` rescales the random-init decoder's output so the image uses the [0, 1] range (same gain for both images)."""
    z = latents.float() / SCALING_FACTOR + SHIFT_FACTOR
    x = dec(z) * gain
    return (x / 2 + 0.5).clamp(0, 1)


def map_diffusers_names(sd):
    """State dict with diffusers' AutoencoderKL decoder names (as used by flite_b200.vae) -> this file's names."""
    out = {}
    up_idx = 0
    for k, v in sd.items():
        k = k[len("decoder."):]
        if k.startswith("mid_block.resnets.0."):
            out["mid.0." + k[len("mid_block.resnets.0."):]] = v
        elif k.startswith("mid_block.resnets.1."):
            out["mid.2." + k[len("mid_block.resnets.1."):]] = v
        elif k.startswith("mid_block.attentions.0."):
            r = k[len("mid_block.attentions.0."):]
            r = (r.replace("group_norm.", "norm.").replace("to_q.", "q.").replace("to_k.", "k.").replace("to_v.", "v.")
                  .replace("to_out.0.", "o."))
            out["mid.1." + r] = v
        elif k.startswith("up_blocks."):
            _, i, kind, j, rest = k.split(".", 4)
            i, j = int(i), int(j)
            base = i * 5                       # 3 resnets + upsample + conv per block in the oracle's Sequential
            if kind == "resnets":
                out[f"up.{base + j}." + rest.replace("conv_shortcut.", "shortcut.")] = v
            else:                              # upsamplers.0.conv.*
                out[f"up.{base + 4}." + rest[len("conv."):]] = v
        elif k.startswith("conv_norm_out."):
            out["norm_out." + k[len("conv_norm_out."):]] = v
        else:
            out[k] = v
    return out


def toy_decode(z: torch.Tensor) -> torch.Tensor:
    """Deterministic, cheap stand-in for ``AutoencoderKL.decode`` used by the pipeline golden fixture
    (oracle/ref_pipeline_shim.py): 3 channel mixes of the latent, 8x nearest upsampling."""
    mix = torch.stack([z[:, i::3].mean(1) for i in range(3)], 1)
    return F.interpolate(mix * 0.9, scale_factor=8, mode="nearest")


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = (a.float() - b.float()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()
