"""TEST INFRASTRUCTURE -- import the UNMODIFIED reference DiT.

``f_lite/model.py`` imports diffusers / peft (not installed), liger_kernel (Triton: GPU only) and
flash_attn_interface (FlashAttention-3: not installed, Hopper-only).  This module registers minimal stub modules for
exactly those import names and loads the reference file by path -- from ``/root/reference`` in the build container,
from the travelled copy ``oracle/_ref/f_lite/model.py`` on the GPU box (``oracle/build_ref.py``).  Two backends:

* ``backend="cpu"``: LigerRMSNorm / LigerSwiGLUMLP / flash_attn_varlen_func are the torch restatements of
  ``oracle/dit_oracle.py`` -- what ``oracle/make_golden.py`` uses to pin the restatement of everything else.
* ``backend="gpu"``: the REAL ``liger_kernel.transformers`` modules (0.8.0, Triton) and the REAL FlashAttention-2
  ``flash_attn.flash_attn_varlen_func`` (2.8.3; same call signature, returns ``out`` where FA3 returns ``(out, lse)``)
  behind the ``flash_attn_interface`` name -- "the reference bf16 path" of BASELINE.json on a B200.  Used by
  ``tests/test_reference_gpu.py`` and ``bench.py --impl reference``; never by the product.
"""
from __future__ import annotations

import importlib.util
import inspect
import sys
import types

import torch
from torch import nn

from .build_ref import ref_path


def ref_model_path() -> str:
    p = ref_path("f_lite/model.py")
    if p is None:
        raise FileNotFoundError("reference model.py not found: neither /root/reference nor oracle/_ref "
                                "(run `python -m oracle.build_ref` in the build container)")
    return p


def _install_stubs(backend: str = "cpu"):
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

    if backend == "gpu":
        # the real third-party kernels: liger_kernel (Triton) as is, FA2 behind FA3's module name
        import flash_attn
        import liger_kernel.transformers  # noqa: F401  (model.py imports the real classes from it)
        # liger 0.8.0's rms_norm references torch.distributed.tensor.DTensor without importing the submodule
        importlib.import_module("torch.distributed.tensor")

        def fa2_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, softmax_scale):
            return flash_attn.flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                                     softmax_scale=softmax_scale, causal=False), None

        mod("flash_attn_interface").flash_attn_varlen_func = fa2_varlen
        return

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


def load_reference_model_module(backend: str = "cpu"):
    names = ("liger_kernel", "liger_kernel.transformers", "flash_attn_interface")
    saved = {k: sys.modules.get(k) for k in names}
    if backend == "cpu":             # a real liger_kernel imported earlier must not shadow the CPU stubs
        for k in names:
            sys.modules.pop(k, None)
    _install_stubs(backend)
    spec = importlib.util.spec_from_file_location(f"_flite_ref_model_{backend}", ref_model_path())
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if backend == "cpu":
        for k, v in saved.items():   # do not leave a stubbed liger_kernel behind
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


def build_reference_dit(cfg: dict, sd: dict, dtype=torch.float32, backend: str = "cpu", device=None):
    """Instantiate the reference DiT, load ``sd`` (reference keys) and cast like users do."""
    m = load_reference_model_module(backend)
    if device is not None:
        # fp32 construction directly on the device (a 10B-architecture init on the host takes minutes), then the user's
        # .to(dtype) (RoPE buffers included).  model.py:342 builds inv_freq with the legacy `torch.FloatTensor([...])`
        # constructor, which ignores the device context: route that one constructor through torch.tensor for the
        # duration of __init__ (same fp32 values; the reference source is not touched).
        legacy = torch.FloatTensor
        torch.FloatTensor = lambda data: torch.tensor(data, dtype=torch.float32)
        try:
            with torch.device(device):
                model = m.DiT(**cfg)
        finally:
            torch.FloatTensor = legacy
    else:
        model = m.DiT(**cfg)
    if sd is not None:
        missing, unexpected = model.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
    model = model.to(dtype).eval()
    return model.to(device) if device is not None else model
