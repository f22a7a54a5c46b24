"""TEST INFRASTRUCTURE, BUILD CONTAINER ONLY -- run the UNMODIFIED reference ``FLitePipeline.__call__``.

``/root/reference/f_lite/pipeline.py`` imports diffusers (not installed).  This module registers stub modules for
exactly the names it imports (``DiffusionPipeline``, ``AutoencoderKL``, ``BaseOutput``, ``randn_tensor``), loads the
reference file by path and drives its ``__call__`` end to end with stand-ins for the three third-party models:

* text encoder / processor: return the caller's embeddings as ``hidden_states[return_index]`` so that
  ``encode_prompt`` (pipeline.py:126-175) runs unmodified;
* DiT: the real reference DiT (``ref_shim``) behind the stale 3-argument call of pipeline.py:271,293 (mask := ones);
* VAE: a deterministic toy decoder with ``config.scaling_factor / shift_factor`` that records its input.

It pins ``oracle/sampler_oracle.py`` (schedule, CFG order, APG, accumulator dtype) and the pipeline-tail restatements
against the reference's own loop.  Never imported on the GPU box (``/root/reference`` does not exist there).
"""
from __future__ import annotations

import importlib.util
import sys
import types

import torch
from torch import nn

from . import ref_shim
from .vae_decoder import SCALING_FACTOR as TOY_SCALING, SHIFT_FACTOR as TOY_SHIFT, toy_decode

from .build_ref import ref_path

REF_PIPELINE = ref_path("f_lite/pipeline.py")


def _install_stubs():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    class DiffusionPipeline:
        def __init__(self):
            pass

        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def _execution_device(self):
            return next(self.dit_model.parameters()).device

        def maybe_free_model_hooks(self):
            pass

    class BaseOutput:
        pass

    def randn_tensor(shape, generator=None, device=None, dtype=None):
        # diffusers.utils.torch_utils.randn_tensor for a single generator: sample on the generator's device, then move
        gdev = generator.device if generator is not None else device
        return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)

    d = mod("diffusers")
    d.DiffusionPipeline = DiffusionPipeline
    d.AutoencoderKL = type("AutoencoderKL", (), {})
    mod("diffusers.utils").BaseOutput = BaseOutput
    mod("diffusers.utils.torch_utils").randn_tensor = randn_tensor


def load_reference_pipeline_module():
    _install_stubs()
    # the reference imports transformers, whose availability probes choke on ref_shim's spec-less stub modules
    parked = {k: sys.modules.pop(k) for k in ("flash_attn_interface",) if k in sys.modules}
    try:
        spec = importlib.util.spec_from_file_location("_flite_ref_pipeline", REF_PIPELINE)
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
    finally:
        sys.modules.update(parked)
    return m


class _ThreeArgDiT(nn.Module):
    """pipeline.py:271,293 call ``dit_model(latents, context, t)``; the mounted model.py needs the mask."""

    def __init__(self, dit):
        super().__init__()
        self.dit = dit

    def forward(self, x, context, t):
        mask = torch.ones(context.shape[:2], dtype=x.dtype, device=x.device)
        return self.dit(x, context, mask, t)


class _Inputs(dict):
    def to(self, device=None, dtype=None):
        return self


class _Processor:
    def __init__(self):
        self.templated = []

    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True):
        self.templated.append(messages)
        return messages[-1]["content"][0]["text"]

    def __call__(self, text, **kw):
        self.kw = kw
        return _Inputs(keys=list(text))


class _TextEncoder(nn.Module):
    """Returns the embeddings registered for each prompt string as hidden_states[-8]."""

    def __init__(self, table, dtype):
        super().__init__()
        self.table = table
        self.p = nn.Parameter(torch.zeros(1, dtype=dtype))

    @property
    def device(self):
        return self.p.device

    def forward(self, keys, use_cache=False, return_dict=True, output_hidden_states=True):
        e = torch.cat([self.table[k] for k in keys])
        hs = [torch.full_like(e, float("nan"))] * 16
        hs[-8] = e
        return types.SimpleNamespace(hidden_states=hs)


class _ToyVAE:
    def __init__(self, dtype):
        self.config = types.SimpleNamespace(scaling_factor=TOY_SCALING, shift_factor=TOY_SHIFT)
        self.dtype = dtype
        self.seen = []

    def decode(self, z):
        self.seen.append(z.clone())
        return types.SimpleNamespace(sample=toy_decode(z))


def run_reference_pipeline(cfg, sd, prompt_embeds, dtype, seed, height, width, steps, guidance, apg_threshold=None):
    """-> (decode-input latents, uint8 images [B, H, W, 3]) from the reference's own __call__ on CPU."""
    import numpy as np
    P = load_reference_pipeline_module()
    dit = _ThreeArgDiT(ref_shim.build_reference_dit(cfg, sd, dtype))
    table = {f"p{i}": prompt_embeds[i:i + 1].to(dtype) for i in range(prompt_embeds.shape[0])}
    vae, proc = _ToyVAE(dtype), _Processor()
    pipe = P.FLitePipeline(dit, vae, _TextEncoder(table, dtype), proc)
    apg = P.APGConfig(enabled=True, orthogonal_threshold=apg_threshold) if apg_threshold is not None else None
    out = pipe(list(table), height=height, width=width, num_inference_steps=steps, guidance_scale=guidance,
               generator=torch.Generator().manual_seed(seed), apg_config=apg)
    imgs = torch.from_numpy(np.stack([np.asarray(im) for im in out.images]))
    return vae.seen[-1], imgs, proc
