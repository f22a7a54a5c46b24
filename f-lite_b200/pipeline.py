"""Sampler loop of the F Lite pipeline on the fused CFG+Euler kernel.

``FLitePipeline`` keeps the reference's constructor and ``__call__`` signature, defaults, time-shift
schedule and return type (``/root/reference/f_lite/pipeline.py:60-62,188-202,329-331``); its loop body
calls ``flite_cfg_euler`` instead of the five elementwise launches + two clones of
``pipeline.py:290,296-297``.  ``denoise()`` is the loop alone (what ``bench.py`` times); it supports both
accumulator semantics found in the reference: bf16 (``FLitePipeline.__call__``) and fp32
(``train.py::sample_images``, f_lite/train.py:599).

Text encoding (Qwen2.5-VL) and VAE decoding are outside the hot path: they are
used if the caller supplies the modules, and can be bypassed with ``prompt_embeds=...`` /
``output_type="latent"``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, List, Optional, Union

import torch

from . import ops


@dataclass
class APGConfig:
    """f_lite/pipeline.py:25-31."""
    enabled: bool = True
    orthogonal_threshold: float = 0.03


@dataclass
class FLitePipelineOutput:
    """f_lite/pipeline.py:33-43."""
    images: Any


def default_alpha(latent_height: int, latent_width: int) -> float:
    """f_lite/pipeline.py:240-242."""
    return 2 * math.sqrt((latent_height * latent_width) / (64 * 64))


def time_shift_schedule(num_inference_steps: int, alpha: float):
    """[(t, dt)] for i = N..1, f_lite/pipeline.py:250-257."""
    out = []
    for i in range(num_inference_steps, 0, -1):
        t = i / num_inference_steps
        t_next = (i - 1) / num_inference_steps
        t = t * alpha / (1 + (alpha - 1) * t)
        t_next = t_next * alpha / (1 + (alpha - 1) * t_next)
        out.append((t, t - t_next))
    return out


@torch.no_grad()
def denoise_step(dit_model, latents, acc, context_input, mask_input, t_tensor, dt: float, guidance_scale: float,
                 do_cfg: bool = True, cfg_group=None):
    """One denoise step: CFG-batched DiT forward ([negative, positive], pipeline.py:264-271) followed by the
    fused CFG-combine + Euler update.  ``acc`` (bf16 or fp32) and ``latents`` (bf16) are updated in place.
    ``t_tensor`` holds the (already time-shifted) t for every model row, in the model dtype (pipeline.py:260).

    ``cfg_group`` (2 ranks): CFG halves on different GPUs ("C2 at 2 GPUs"): rank 0 of the group
    evaluates the negative half, rank 1 the positive half, one all-gather of the velocities per step, then both
    ranks apply the same update.  ``context_input`` / ``mask_input`` / ``t_tensor`` then hold this rank's half."""
    b = latents.shape[0]
    if do_cfg and cfg_group is not None:
        import torch.distributed as dist
        mine = dit_model(latents, context_input, mask_input, t_tensor)
        both = torch.empty((2,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(both.view(2 * b, *mine.shape[1:]), mine.contiguous(), group=cfg_group)
        ops.cfg_euler(acc, both[0], both[1], guidance_scale, dt, latents, do_cfg=True)
        return both.view(2 * b, *mine.shape[1:])
    if do_cfg:
        out = dit_model(torch.cat([latents] * 2), context_input, mask_input, t_tensor)
        uncond, cond = out[:b], out[b:]
        ops.cfg_euler(acc, uncond, cond, guidance_scale, dt, latents, do_cfg=True)
    else:
        out = dit_model(latents, context_input, mask_input, t_tensor)
        ops.cfg_euler(acc, None, out, guidance_scale, dt, latents, do_cfg=False)
    return out


@torch.no_grad()
def denoise(dit_model, latents, negative_embeds, prompt_embeds, mask=None, num_inference_steps: int = 30,
            guidance_scale: float = 6.0, alpha: Optional[float] = None, acc_dtype=torch.bfloat16,
            apg_config: Optional[APGConfig] = None, trace: Optional[list] = None, cfg_group=None,
            cuda_graph: bool = False, progress=None):
    """The sampling loop of f_lite/pipeline.py:244-297 (acc_dtype=bf16) / f_lite/train.py:573-599
    (acc_dtype=fp32).  ``mask`` covers ``[negative, positive]`` rows (None = all ones).

    ``cuda_graph``: capture the DiT forward once and replay it every step (``graphs.GraphedForward``) -- for
    launch-bound shapes; bit-identical to the eager loop.  Single-GPU only."""
    b = latents.shape[0]
    latents = latents.to(torch.bfloat16).contiguous().clone()
    acc = latents.to(acc_dtype).clone()
    if alpha is None:
        alpha = default_alpha(latents.shape[2], latents.shape[3])
    do_cfg = guidance_scale >= 1.0                                                        # pipeline.py:248
    split_cfg = do_cfg and cfg_group is not None
    if split_cfg:
        import torch.distributed as dist
        half = dist.get_rank(cfg_group)                 # 0: negative half, 1: positive half
        context_input = (negative_embeds, prompt_embeds)[half].contiguous()
        mask_input = None if mask is None else mask[half * b:(half + 1) * b].contiguous()
    elif do_cfg:
        context_input = torch.cat([negative_embeds, prompt_embeds])
        mask_input = mask
    else:
        context_input = prompt_embeds
        mask_input = None if mask is None else mask[b:]
    apg = apg_config is not None and apg_config.enabled
    sched = time_shift_schedule(num_inference_steps, alpha)
    rows = 2 * b if (do_cfg and not split_cfg) else b
    # torch.tensor([t] * batch, dtype=model dtype) of pipeline.py:260, for every step, in one H2D copy
    t_all = torch.tensor([[t] * rows for t, _ in sched], dtype=latents.dtype).to(latents.device)
    graphed = None
    if cuda_graph:
        if cfg_group is not None:
            raise ValueError("cuda_graph=True is single-GPU only")
        from .graphs import GraphedForward
        graphed = GraphedForward(dit_model, latents, context_input, mask_input, t_all[0], duplicate_latents=do_cfg)
    steps_iter = range(len(sched)) if progress is None else progress(range(len(sched)))   # pipeline.py:250 progress_bar
    for step in steps_iter:
        t, dt = sched[step]
        t_tensor = t_all[step]
        if graphed is not None:
            out = graphed(t_tensor)
            if apg and do_cfg:
                ops.apg_euler(acc, out[:b], out[b:], guidance_scale, dt, apg_config.orthogonal_threshold, latents)
            elif do_cfg:
                ops.cfg_euler(acc, out[:b], out[b:], guidance_scale, dt, latents, do_cfg=True)
            else:
                ops.cfg_euler(acc, None, out, guidance_scale, dt, latents, do_cfg=False)
        elif apg and do_cfg and not split_cfg:
            # Augmented Parallel Guidance (pipeline.py:276-287): the three global reductions and the update run on the
            # device (flite_apg_euler), no host sync.
            out = dit_model(torch.cat([latents] * 2), context_input, mask_input, t_tensor)
            ops.apg_euler(acc, out[:b], out[b:], guidance_scale, dt, apg_config.orthogonal_threshold, latents)
        elif apg and do_cfg:
            import torch.distributed as dist
            mine = dit_model(latents, context_input, mask_input, t_tensor)
            out = torch.empty((2 * b,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(out, mine.contiguous(), group=cfg_group)
            ops.apg_euler(acc, out[:b], out[b:], guidance_scale, dt, apg_config.orthogonal_threshold, latents)
        else:
            out = denoise_step(dit_model, latents, acc, context_input, mask_input, t_tensor, dt, guidance_scale, do_cfg,
                               cfg_group=cfg_group if split_cfg else None)
        if trace is not None:
            trace.append(out.clone())
    # one synchronisation per trajectory: surface a kernel-side barrier time-out (flite_watchdog_status) instead of
    # returning whatever the aborted kernels left behind
    from . import _lib
    _lib.watchdog_ok()
    return latents if acc_dtype == torch.bfloat16 else acc


class FLitePipeline:
    """Same public surface as ``f_lite.pipeline.FLitePipeline`` for the sampling path."""

    model_cpu_offload_seq = "text_encoder->dit_model->vae"

    def __init__(self, dit_model, vae=None, text_encoder=None, processor=None):
        self.dit_model = dit_model
        self.vae = vae
        self.text_encoder = text_encoder
        self.processor = processor
        # pipeline.py:80-83
        self.vae_scale_factor = 8
        self.return_index = -8

    @property
    def _execution_device(self):
        return next(self.dit_model.parameters()).device

    def to(self, torch_device=None, torch_dtype=None, silence_dtype_warnings=False):
        for m in (self.vae, self.text_encoder, self.dit_model):
            if m is not None:
                m.to(device=torch_device, dtype=torch_dtype)
        return self

    # pipeline.py:84-102: memory / progress helpers of the reference surface
    def enable_vae_slicing(self):
        if hasattr(self.vae, "enable_slicing"):
            self.vae.enable_slicing()

    def enable_vae_tiling(self):
        if hasattr(self.vae, "enable_tiling"):
            self.vae.enable_tiling()

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, **kwargs):
        config = {**(getattr(self, "_progress_bar_config", None) or {}), **kwargs}
        if config.get("disable", False):
            return iterable
        try:
            from tqdm.auto import tqdm
        except ImportError:
            return iterable
        return tqdm(iterable, **config)

    # the system turn the reference puts in front of every caption (f_lite/pipeline.py:105); F Lite was trained with it,
    # so it is part of the text-conditioning interface, not a tunable
    SYSTEM_PROMPT = (
        "You are a text-to-image generation model engineered to transform user-provided textual captions directly into "
        "high-quality, visually rich image tokens. Your core objective is to generate the best possible, highest-fidelity "
        "image that creatively interprets and expands upon the user's intent while maintaining strong semantic alignment "
        "with the original caption. You are designed for maximum visual quality, artistic flair, and implicit adherence to "
        "best practices in image generation (e.g., proper anatomy, clear focus, compelling composition), ensuring a "
        "stunning visual result from even concise descriptions.")

    def _convert_caption_to_messages(self, caption: str) -> str:
        """pipeline.py:104-124 (system turn + user caption through the processor's chat template)."""
        messages = [{"role": "system", "content": self.SYSTEM_PROMPT},
                    {"role": "user", "content": [{"type": "text", "text": caption}]}]
        return self.processor.apply_chat_template(messages, tokenize=False, add_generation_prompt=True)

    def encode_prompt_with_masks(self, prompt, negative_prompt=None, device=None, dtype=None, max_sequence_length=512,
                                 return_index=-8):
        """pipeline.py:126-175 plus the attention masks the 4-argument DiT forward needs (model.py:526): returns
        ``(prompt_embeds, negative_embeds, prompt_mask, negative_mask)``."""
        if self.text_encoder is None:
            raise RuntimeError("no text_encoder: pass prompt_embeds= / negative_embeds= to __call__")
        if isinstance(prompt, str):
            prompt = [prompt]
        device = device or self.text_encoder.device
        messages = [self._convert_caption_to_messages(p) for p in prompt]
        text_inputs = self.processor(text=messages, padding="longest", pad_to_multiple_of=8,
                                     max_length=max_sequence_length, truncation=True, return_tensors="pt").to(device)
        enc = self.text_encoder(**text_inputs, use_cache=False, return_dict=True, output_hidden_states=True)
        dtype = dtype or next(self.text_encoder.parameters()).dtype
        embeds = enc.hidden_states[return_index].to(dtype=dtype, device=device)
        mask = text_inputs["attention_mask"].to(device)
        if negative_prompt is None:
            neg, neg_mask = torch.zeros_like(embeds), torch.ones_like(mask)
        else:
            if isinstance(negative_prompt, str):
                negative_prompt = [negative_prompt]
            neg, _, neg_mask, _ = self.encode_prompt_with_masks(negative_prompt, device=device, dtype=dtype,
                                                                return_index=return_index)
        return embeds, neg, mask, neg_mask

    def encode_prompt(self, prompt, negative_prompt=None, device=None, dtype=None, max_sequence_length=512,
                      return_index=-8):
        """Same contract as the reference (pipeline.py:126-175): returns ``(prompt_embeds, negative_embeds)``."""
        embeds, neg, _, _ = self.encode_prompt_with_masks(prompt, negative_prompt, device, dtype, max_sequence_length,
                                                          return_index)
        return embeds, neg

    @torch.no_grad()
    def __call__(
        self,
        prompt: Union[str, List[str], None] = None,
        height: Optional[int] = 1024,
        width: Optional[int] = 1024,
        num_inference_steps: int = 30,
        guidance_scale: float = 6.0,
        negative_prompt: Optional[Union[str, List[str]]] = None,
        num_images_per_prompt: int = 1,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        dtype: Optional[torch.dtype] = None,
        alpha: Optional[float] = None,
        apg_config: Optional[APGConfig] = None,
        **kwargs,
    ):
        height = 1024 if height is None else height
        width = 1024 if width is None else width
        dtype = dtype or next(self.dit_model.parameters()).dtype
        apg_config = apg_config or APGConfig(enabled=False)
        device = self._execution_device

        prompt_embeds = kwargs.pop("prompt_embeds", None)
        negative_embeds = kwargs.pop("negative_embeds", None)
        prompt_mask = kwargs.pop("prompt_attention_mask", None)
        negative_mask = kwargs.pop("negative_attention_mask", None)
        output_type = kwargs.pop("output_type", "pil")
        acc_dtype = kwargs.pop("acc_dtype", dtype)
        cuda_graph = kwargs.pop("cuda_graph", False)       # replay the DiT forward from a CUDA graph (launch-bound shapes)
        if prompt_embeds is None:
            prompt_embeds, negative_embeds, prompt_mask, negative_mask = self.encode_prompt_with_masks(
                prompt, negative_prompt, device=device, dtype=dtype, return_index=self.return_index)
        if negative_embeds is None:
            negative_embeds = torch.zeros_like(prompt_embeds)                      # pipeline.py:160-161
        prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, dim=0).to(device)
        negative_embeds = negative_embeds.repeat_interleave(num_images_per_prompt, dim=0).to(device)
        batch_size = prompt_embeds.shape[0]
        if prompt_mask is None:
            prompt_mask = torch.ones(prompt_embeds.shape[:2], device=device)
        else:
            prompt_mask = prompt_mask.repeat_interleave(num_images_per_prompt, dim=0).to(device)
        if negative_mask is None:
            negative_mask = torch.ones(negative_embeds.shape[:2], device=device)
        else:
            negative_mask = negative_mask.repeat_interleave(num_images_per_prompt, dim=0).to(device)
        mask = torch.cat([negative_mask.to(torch.float32), prompt_mask.to(torch.float32)])

        latent_height = height // self.vae_scale_factor
        latent_width = width // self.vae_scale_factor
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"Got {len(generator)} generators for {batch_size} samples")
        latents = kwargs.pop("latents", None)
        if latents is None:
            shape = (batch_size, 16, latent_height, latent_width)
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=dtype).to(device)
                                     for g in generator])
            else:
                gdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)

        self.dit_model.eval()
        latents = denoise(self.dit_model, latents, negative_embeds, prompt_embeds, mask, num_inference_steps,
                          guidance_scale, alpha, acc_dtype=acc_dtype, apg_config=apg_config, cuda_graph=cuda_graph,
                          progress=self.progress_bar)
        if output_type == "latent" or self.vae is None:
            return FLitePipelineOutput(images=latents)

        # pipeline.py:299-327 (decode + post-process); VAE is a "next" row, run through the caller's module
        scaling = getattr(self.vae.config, "scaling_factor", 0.18215) if hasattr(self.vae, "config") else 0.18215
        shift = getattr(self.vae.config, "shift_factor", 0) if hasattr(self.vae, "config") else 0
        lat = ops.latent_unscale(latents.contiguous(), scaling, shift)
        vae_dtype = self.vae.dtype if hasattr(self.vae, "dtype") else dtype
        decoded = self.vae.decode(lat.to(vae_dtype))
        decoded = decoded.sample if hasattr(decoded, "sample") else decoded
        if decoded.dtype not in (torch.bfloat16, torch.float32):
            decoded = decoded.float()
        images_hwc = ops.image_to_uint8(decoded.contiguous()).cpu()           # [B, H, W, C] uint8, PIL layout
        images = images_hwc.permute(0, 3, 1, 2)                               # the reference's NCHW view
        if output_type == "pt":
            return FLitePipelineOutput(images=images)
        from PIL import Image
        return FLitePipelineOutput(images=[Image.fromarray(img.numpy()) for img in images_hwc])
