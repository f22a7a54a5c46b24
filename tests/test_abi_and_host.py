"""CPU: the C-ABI library loads and exports every declared symbol; host-side mirrors of the reference API."""
import ctypes
import os
import re

import pytest
import torch

import flite_b200
from flite_b200 import _lib, ops, pipeline
from oracle import sampler_oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "flite_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flite_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/flite_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.load().flite_abi_version() == 1


def test_sass_is_blackwell_native():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass     # tcgen05.mma / TMA / tcgen05.ld
    assert "UTCHMMA.2CTA" in sass                                          # cta_group::2 pair
    assert "HMMA.16" not in sass                                           # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_fails_loudly_without_gpu():
    lib = _lib.load()
    assert lib.flite_check_device() != 0
    assert lib.flite_last_error()
    m = flite_b200.DiT(**dict(synth.TINY, depth=1))
    with pytest.raises(flite_b200.FliteError):
        m(torch.zeros(1, 16, 32, 32), torch.zeros(1, 8, 4096), torch.ones(1, 8), torch.ones(1))
    with pytest.raises(flite_b200.FliteError):   # legacy 3-argument call form is routed the same way
        m(torch.zeros(1, 16, 32, 32), torch.zeros(1, 8, 4096), torch.ones(1))
    with pytest.raises(_lib.FliteError):
        ops.cfg_euler(torch.zeros(8), torch.zeros(8), torch.zeros(8), 6.0, 0.1, torch.zeros(8))


def test_bad_arguments_return_error_codes():
    lib = _lib.load()
    assert lib.flite_gemm_bf16(None, 8, None, 8, None, 8, 1, 64, 64, None, 0, 0, None, 0, None, 0, 0, None, None, 0,
                               1e-6, 0, 0, 0, None) == -1
    assert b"null" in lib.flite_last_error()
    one = ctypes.c_void_p(16)
    assert lib.flite_gemm_bf16(one, 8, one, 8, one, 8, 1, 64, 63, None, 0, 0, None, 0, None, 0, 0, None, None, 0,
                               1e-6, 0, 0, 0, None) == -1
    assert b"K = 63" in lib.flite_last_error()
    assert lib.flite_cfg_euler(one, 0, one, one, 6.0, 0.1, 1, one, 12, None) == -1
    # attention: the q / k / v views must hold col0 + 256*H columns inside their row stride (the tensor maps declare
    # exactly that width; a map wider than the view would reach past the end of the buffer)
    assert lib.flite_attention_varlen(one, 512, 128, 0, one, 512, 128, 0, one, 512, 0, one, 512, one, one, 1, 3, 128,
                                      0.0625, 0, None) == -1
    assert b"row stride" in lib.flite_last_error()
    assert lib.flite_attention_varlen(one, 768, 128, 0, one, 768, 128, 8, one, 768, 0, one, 768, one, one, 1, 3, 128,
                                      0.0625, 0, None) == -1                      # k_col0 = 8 pushes K past its stride
    assert lib.flite_apg_euler(one, 0, one, one, 6.0, 0.1, 0.03, one, 8, one, None) == -1   # numel must exceed 8
    assert lib.flite_latent_unscale(one, one, 0.0, 0.1, 16, None) == -1                     # zero scaling factor
    assert lib.flite_gemm_bf16(one, 64, one, 64, one, 64, 16, 64, 64, None, 0, 0, None, 0, None, 0, 0, None, None, 0,
                               1e-6, 0, 0, 5, None) == -1                                   # GEMV variant needs M <= 8


@pytest.mark.parametrize("cfg", [synth.TINY, dict(synth.TINY, depth=9, train_bias_and_rms=False)])
def test_state_dict_layout_matches_reference(cfg):
    m = flite_b200.DiT(**cfg)
    sd = m.state_dict()
    want = synth.param_shapes(cfg)         # key list pinned against the real reference in make_golden
    assert set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    for k in ("patch_size", "hidden_size", "use_rope", "gradient_checkpoint", "depth"):
        assert getattr(m.config, k) == cfg[k]
    # cross-attention placement: idx % 4 == 0 or idx < 8 (model.py:464)
    have = [i for i, b in enumerate(m.blocks) if b.cross_attn is not None]
    assert have == synth.cross_attn_blocks(cfg["depth"])
    # the reference zero-inits these (model.py:455-456,476-479)
    assert m.adaLN_modulation[1].weight.abs().sum() == 0 and m.final_proj.weight.abs().sum() == 0


def test_reference_state_dict_keys_are_what_the_reference_module_has():
    ref_model = "/root/reference/f_lite/model.py"
    if not os.path.exists(ref_model):
        pytest.skip("reference tree not mounted (GPU box)")
    from oracle import ref_shim
    cfg = dict(synth.TINY, depth=2)
    ref = ref_shim.build_reference_dit(cfg, synth.make_state_dict(cfg))
    mine = flite_b200.DiT(**cfg)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_schedule_matches_reference_loop():
    for n, (h, w) in ((4, (32, 32)), (30, (128, 128)), (30, (112, 168))):
        a = pipeline.default_alpha(h, w)
        assert a == sampler_oracle.default_alpha(h, w)
        assert pipeline.time_shift_schedule(n, a) == sampler_oracle.schedule(n, a)


def test_interleave_gate_up_layout():
    inter, d = 256, 8
    g = torch.arange(inter * d, dtype=torch.float32).view(inter, d)
    u = -g
    w = ops.interleave_gate_up(g, u)
    assert w.shape == (2 * inter, d)
    for n in range(2 * inter):
        grp, off = divmod(n, 128)
        src = g if off < 64 else u
        assert torch.equal(w[n], src[grp * 64 + off % 64])


def test_rope_table_matches_oracle():
    from flite_b200.model import _rope_table
    from oracle.dit_oracle import rope_tables
    cos, sin = _rope_table(256, 6, 10, 10000, 16, round_bf16=True)
    oc, os_ = rope_tables(256, 6, 10, 10000, "cpu", torch.bfloat16)
    assert torch.equal(cos, oc[0].float()) and torch.equal(sin, os_[0].float())
    assert torch.equal(cos.bfloat16().float(), cos)      # exactly representable: stored as bf16 on the device
    assert torch.all(cos[:16] == 1) and torch.all(sin[:16] == 0)


def test_pipeline_signature_matches_reference():
    import inspect
    sig = inspect.signature(pipeline.FLitePipeline.__call__)
    want = ["self", "prompt", "height", "width", "num_inference_steps", "guidance_scale", "negative_prompt",
            "num_images_per_prompt", "generator", "dtype", "alpha", "apg_config", "kwargs"]
    assert list(sig.parameters) == want           # f_lite/pipeline.py:188-202
    p = sig.parameters
    assert (p["height"].default, p["width"].default, p["num_inference_steps"].default,
            p["guidance_scale"].default) == (1024, 1024, 30, 6.0)
    fsig = inspect.signature(flite_b200.DiT.forward)
    assert list(fsig.parameters)[:5] == ["self", "x", "context", "context_attn_mask", "timesteps"]  # model.py:526


def test_pipeline_public_surface_matches_reference():
    """Attributes / helper methods of f_lite.pipeline.FLitePipeline that user code touches (pipeline.py:47-102,176-184;
    generate.py:77-78 calls pipe.vae.enable_slicing / enable_tiling through these)."""
    P = pipeline.FLitePipeline
    for name in ("enable_vae_slicing", "enable_vae_tiling", "set_progress_bar_config", "progress_bar", "encode_prompt",
                 "to", "_convert_caption_to_messages"):
        assert callable(getattr(P, name)), name
    assert P.model_cpu_offload_seq == "text_encoder->dit_model->vae"

    class V:
        sliced = tiled = False

        def enable_slicing(self): self.sliced = True
        def enable_tiling(self): self.tiled = True

    pipe = P(torch.nn.Linear(1, 1), V(), None, None)
    pipe.enable_vae_slicing(); pipe.enable_vae_tiling()
    assert pipe.vae.sliced and pipe.vae.tiled and pipe.vae_scale_factor == 8 and pipe.return_index == -8
    pipe.set_progress_bar_config(disable=True)
    assert list(pipe.progress_bar(range(3))) == [0, 1, 2]
    o = pipeline.FLitePipelineOutput(images=[1])
    assert o.images == [1]
    a = pipeline.APGConfig()
    assert a.enabled is True and a.orthogonal_threshold == 0.03          # pipeline.py:25-31


def test_encode_prompt_contract_with_stand_in_encoder():
    """Text-encoder side of the pipeline (f_lite/pipeline.py:104-175) exercised on the CPU with stand-ins for the
    third-party processor / encoder: chat template -> processor(text=..., padding='longest', pad_to_multiple_of=8,
    max_length=512, truncation=True) -> hidden_states[return_index]; negative prompt None => zeros (pipeline.py:160-161);
    same 2-tuple return as the reference, masks through encode_prompt_with_masks."""
    import torch
    from flite_b200 import FLitePipeline

    class Batch(dict):
        def to(self, *a, **k):
            return self

    class Proc:
        def __init__(self):
            self.calls = []

        def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True):
            assert messages[0]["role"] == "system" and messages[1]["role"] == "user"
            assert tokenize is False and add_generation_prompt is True
            return "<sys>" + messages[0]["content"][:8] + "<user>" + messages[1]["content"][0]["text"]

        def __call__(self, text, padding, pad_to_multiple_of, max_length, truncation, return_tensors):
            self.calls.append(dict(text=text, padding=padding, pad_to_multiple_of=pad_to_multiple_of,
                                   max_length=max_length, truncation=truncation, return_tensors=return_tensors))
            lens = [len(t) % 13 + 3 for t in text]
            L = (max(lens) + 7) // 8 * 8
            ids = torch.zeros(len(text), L, dtype=torch.long)
            mask = torch.zeros(len(text), L, dtype=torch.long)
            for i, n in enumerate(lens):
                ids[i, :n] = torch.arange(1, n + 1) + i
                mask[i, :n] = 1
            return Batch(input_ids=ids, attention_mask=mask)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(64, 32)
            self.device = torch.device("cpu")

        def forward(self, input_ids, attention_mask, use_cache, return_dict, output_hidden_states):
            assert use_cache is False and return_dict and output_hidden_states
            h = self.emb(input_ids)
            from types import SimpleNamespace
            return SimpleNamespace(hidden_states=[h * (i + 1) for i in range(12)])

    proc, enc = Proc(), Enc()
    pipe = FLitePipeline(None, None, enc, proc)
    out = pipe.encode_prompt(["a red fox", "two cats on a sofa"], dtype=torch.float32)
    assert isinstance(out, tuple) and len(out) == 2                     # the reference's contract (pipeline.py:175)
    emb, neg = out
    assert proc.calls[0]["padding"] == "longest" and proc.calls[0]["pad_to_multiple_of"] == 8
    assert proc.calls[0]["max_length"] == 512 and proc.calls[0]["truncation"] is True
    assert emb.shape[0] == 2 and emb.shape[1] % 8 == 0 and torch.equal(neg, torch.zeros_like(emb))
    ids = proc(text=proc.calls[0]["text"], padding="longest", pad_to_multiple_of=8, max_length=512, truncation=True,
               return_tensors="pt")["input_ids"]
    assert torch.equal(emb, enc.emb(ids) * 5)                           # hidden_states[-8] of 12
    e2, n2, m2, nm2 = pipe.encode_prompt_with_masks("a red fox", negative_prompt="blurry", dtype=torch.float32)
    assert e2.shape[0] == 1 and n2.shape[0] == 1 and m2.dtype == torch.long and nm2.sum() > 0
    assert not torch.equal(n2, torch.zeros_like(n2))                    # a real negative prompt is encoded, not zeros
    assert "<user>blurry" in proc.calls[-1]["text"][0] and "<user>a red fox" in proc.calls[-2]["text"][0]


def test_custom_op_registrations_and_fake_shapes():
    """the hot entry points are also registered with torch.library.custom_op (opaque operators with
    fake / meta implementations and declared mutations) for callers that live in a torch graph."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    from flite_b200 import custom_ops
    for name in custom_ops.REGISTERED:
        assert hasattr(torch.ops.flite_b200, name), name
    with FakeTensorMode():
        a = torch.empty(5, 64, dtype=torch.bfloat16, device="cuda")
        w = torch.empty(128, 64, dtype=torch.bfloat16, device="cuda")
        assert tuple(torch.ops.flite_b200.linear(a, w, None).shape) == (5, 128)
        q = torch.empty(40, 512, dtype=torch.bfloat16, device="cuda")
        cu = torch.empty(3, dtype=torch.int32, device="cuda")
        assert tuple(torch.ops.flite_b200.attention_varlen(q, q, q, cu, cu, 2, 20, 0.0625).shape) == (40, 512)
        x = torch.empty(7, 256, dtype=torch.bfloat16, device="cuda")
        assert tuple(torch.ops.flite_b200.rmsnorm_modulate(x, x[0], x[:1], x[:1], 7).shape) == (7, 256)
    schema = str(torch.ops.flite_b200.cfg_euler_.default._schema)
    assert "Tensor(a0!) acc" in schema and "lat_out" in schema            # in-place arguments are declared as mutated
