"""TEST INFRASTRUCTURE, BUILD CONTAINER ONLY -- import the UNMODIFIED reference DiT.

``/root/reference/f_lite/model.py`` imports diffusers / peft / liger_kernel /
flash_attn_interface, none of which are usable on CPU in this image.  This module registers
minimal stub modules for exactly those import names and loads the reference file by path, so
that ``oracle/make_golden.py`` can run the *real* module and pin the restatement in
``oracle/dit_oracle.py`` against it.  ``/root/reference`` does not exist on the GPU box, so
nothing imported at GPU-test / bench / smoke time may import this file.
"""
from __future__ import annotations

import importlib.util
import inspect
import sys
import types

import torch
from torch import nn

REF_MODEL = "/root/reference/f_lite/model.py"


def _install_stubs():
    from . import dit_oracle as O

    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    class ConfigMixin:
        pass

    def register_to_config(init):
        sig = inspect.signature(init)

        def wrapped(self, *a, **kw):
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            object.__setattr__(self, "_cfg", _Cfg(cfg))
            init(self, *a, **kw)

        return wrapped

    class ModelMixin(nn.Module):
        @property
        def config(self):
            return self._cfg

    class _Empty:
        pass

    for n in ("diffusers", "diffusers.models", "diffusers.utils", "peft"):
        mod(n)
    mod("diffusers.configuration_utils").ConfigMixin = ConfigMixin
    mod("diffusers.configuration_utils").register_to_config = register_to_config
    mod("diffusers.loaders").FromOriginalModelMixin = _Empty
    mod("diffusers.loaders").PeftAdapterMixin = type("PeftAdapterMixin", (), {})
    mod("diffusers.models.modeling_utils").ModelMixin = ModelMixin
    mod("diffusers.utils.accelerate_utils").apply_forward_hook = lambda f: f
    mod("peft").get_peft_model_state_dict = lambda *a, **k: {}
    mod("peft").set_peft_model_state_dict = lambda *a, **k: None

    class LigerRMSNorm(nn.Module):
        def __init__(self, hidden_size, eps=1e-6):
            super().__init__()
            self.weight = nn.Parameter(torch.ones(hidden_size))
            self.eps = eps

        def forward(self, x):
            return O.liger_rms_norm(x, self.weight, self.eps)

    class LigerSwiGLUMLP(nn.Module):
        def __init__(self, config):
            super().__init__()
            self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
            self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
            self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)

        def forward(self, x):
            return self.down_proj(O.liger_swiglu(self.gate_proj(x), self.up_proj(x)))

    lk = mod("liger_kernel.transformers")
    mod("liger_kernel")
    lk.LigerRMSNorm = LigerRMSNorm
    lk.LigerSwiGLUMLP = LigerSwiGLUMLP

    def flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                               softmax_scale):
        return O.flash_attn_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, softmax_scale), None

    mod("flash_attn_interface").flash_attn_varlen_func = flash_attn_varlen_func


def load_reference_model_module():
    saved = {k: sys.modules.get(k) for k in ("liger_kernel", "liger_kernel.transformers")}
    _install_stubs()
    spec = importlib.util.spec_from_file_location("_flite_ref_model", REF_MODEL)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for k, v in saved.items():       # do not leave a stubbed liger_kernel behind
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return m


def build_reference_dit(cfg: dict, sd: dict, dtype=torch.float32):
    """Instantiate the reference DiT, load ``sd`` (reference keys) and cast like users do."""
    m = load_reference_model_module()
    model = m.DiT(**cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.to(dtype).eval()
