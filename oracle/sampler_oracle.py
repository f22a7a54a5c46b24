"""TEST INFRASTRUCTURE -- restatement of the reference's rectified-flow Euler / CFG loops.

Two loops exist in the reference:

* ``FLitePipeline.__call__`` (f_lite/pipeline.py:244-297): batched CFG with input order
  ``[negative, positive]``, accumulator kept in the model dtype.
* ``train.py::sample_images`` (f_lite/train.py:573-599): two separate forwards, accumulator
  promoted to fp32 by ``dt * v.float()``.
"""
from __future__ import annotations

import math

import torch


def default_alpha(latent_h: int, latent_w: int) -> float:
    """f_lite/pipeline.py:240-242 / f_lite/train.py:574-575."""
    return 2 * math.sqrt((latent_h * latent_w) / (64 * 64))


def schedule(num_steps: int, alpha: float):
    """f_lite/pipeline.py:250-257: list of (t, dt) for i = N..1 with the alpha time shift."""
    out = []
    for i in range(num_steps, 0, -1):
        t = i / num_steps
        t_next = (i - 1) / num_steps
        t = t * alpha / (1 + (alpha - 1) * t)
        t_next = t_next * alpha / (1 + (alpha - 1) * t_next)
        out.append((t, t - t_next))
    return out


def cfg_combine(uncond, cond, guidance_scale: float):
    """f_lite/pipeline.py:290 / f_lite/train.py:596."""
    return uncond + guidance_scale * (cond - uncond)


def apg_combine(uncond, cond, guidance_scale: float, orthogonal_threshold: float):
    """f_lite/pipeline.py:276-287 (Augmented Parallel Guidance)."""
    dy = cond
    dd = cond - uncond
    parallel = (dy * dd).sum() / (dy * dy).sum() * dy
    orth = dd - parallel
    orth_std = orth.std()
    orth_scale = min(1, orthogonal_threshold / orth_std)
    orth = orth * orth_scale
    return dy + (guidance_scale - 1) * orth


@torch.no_grad()
def sample_pipeline(model_fn, latents, negative_embeds, prompt_embeds, mask, num_steps,
                    guidance_scale, alpha=None, apg=None, trace=None):
    """f_lite/pipeline.py:236-297.  ``model_fn(x, context, mask, t)`` is the 4-arg forward
    (the shipped 3-arg call is stale); ``mask`` is for ``[negative, positive]``.
    """
    b = latents.shape[0]
    dtype = latents.dtype
    acc = latents.clone()
    if alpha is None:
        alpha = default_alpha(latents.shape[2], latents.shape[3])
    do_cfg = guidance_scale >= 1.0                                   # pipeline.py:248
    for t, dt in schedule(num_steps, alpha):
        t_tensor = torch.tensor([t] * b, device=latents.device, dtype=dtype)
        if do_cfg:
            out = model_fn(torch.cat([latents] * 2), torch.cat([negative_embeds, prompt_embeds]),
                           mask, torch.cat([t_tensor] * 2))
            uncond, cond = out.chunk(2)
            if apg is not None:
                v = apg_combine(uncond, cond, guidance_scale, apg)
            else:
                v = cfg_combine(uncond, cond, guidance_scale)
        else:
            v = model_fn(latents, prompt_embeds, mask[b:], t_tensor)
        if trace is not None:
            trace.append(v.clone())
        acc = acc + dt * v
        latents = acc.clone()
    return latents


@torch.no_grad()
def sample_train(model_fn, latents, negative_embeds, prompt_embeds, neg_mask, pos_mask,
                 num_steps, cfg_scale):
    """f_lite/train.py:573-599 (two forwards per step, fp32 accumulator after step 1)."""
    alpha = default_alpha(latents.shape[2], latents.shape[3])
    for t, dt in schedule(num_steps, alpha):
        t_tensor = torch.tensor([t] * latents.shape[0]).to(latents.device, torch.bfloat16
                                                            if prompt_embeds.dtype == torch.bfloat16
                                                            else prompt_embeds.dtype)
        v = model_fn(latents.to(prompt_embeds.dtype), prompt_embeds, pos_mask, t_tensor)
        if cfg_scale > 1:
            u = model_fn(latents.to(prompt_embeds.dtype), negative_embeds, neg_mask, t_tensor)
            v = u + cfg_scale * (v - u)
        latents = latents + dt * v.to(dtype=torch.float32)
    return latents
