"""CPU: the oracle restatement against the committed golden vectors of the REAL reference module."""
import os

import numpy as np
import pytest
import torch

from oracle import dit_oracle, sampler_oracle, synth
from oracle.make_golden import CASES, build_case


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_fp32(name, golden_dir):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    sd, x, ctx, mask, t = build_case(g["recipe"])
    y = dit_oracle.dit_forward(sd, g["recipe"]["cfg"], x, ctx, mask, t)
    assert y.shape == g["velocity_fp32"].shape
    assert _rel(y, g["velocity_fp32"]) <= 1e-5           # fp32: same ops, tolerance only for BLAS threading
    assert g["velocity_fp32"].std() > 0.1                 # de-zeroed init: the parity is not vacuous (D10)


@pytest.mark.parametrize("name", ["tiny_256", "tiny_nobias"])
def test_forward_matches_reference_bf16(name, golden_dir):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    sd, x, ctx, mask, t = build_case(g["recipe"])
    sd = {k: v.bfloat16() for k, v in sd.items()}
    y = dit_oracle.dit_forward(sd, g["recipe"]["cfg"], x.bfloat16(), ctx.bfloat16(), mask.bfloat16(), t.bfloat16())
    assert _rel(y, g["velocity_bf16"]) <= 2e-3            # bit-equal here; slack for other CPUs' bf16 GEMM paths


def test_sampler_matches_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_256_sampler.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec)
    b = rec["batch"]
    fn = lambda *a: dit_oracle.dit_forward(sd, rec["cfg"], *a)
    trace = []
    lat = sampler_oracle.sample_pipeline(fn, x[:b], ctx[:b], ctx[b:], mask, g["steps"], g["guidance"], trace=trace)
    assert _rel(lat, g["latents_pipeline"]) <= 1e-5
    assert _rel(torch.stack(trace), g["velocities"]) <= 1e-5
    lat2 = sampler_oracle.sample_train(fn, x[:b], ctx[:b], ctx[b:], mask[:b], mask[b:], g["steps"], g["guidance"])
    assert _rel(lat2, g["latents_train"]) <= 1e-5


def test_schedule_constants():
    # alpha = 2*sqrt(h*w/4096)
    assert sampler_oracle.default_alpha(32, 32) == 1.0
    assert sampler_oracle.default_alpha(128, 128) == 4.0
    assert sampler_oracle.default_alpha(256, 256) == 8.0
    s = sampler_oracle.schedule(30, 4.0)
    assert len(s) == 30 and abs(s[0][0] - 1.0) < 1e-12
    assert abs(sum(dt for _, dt in s) - 1.0) < 1e-9        # the dt's telescope from t=1 to t=0


def test_hash_generator_is_portable():
    # splitmix64 in torch int64 arithmetic == the same in numpy uint64 (well-defined wrap-around)
    n, seed = 1000, 7
    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64((seed * 0x9E3779B97F4A7C15) % 2**64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    ref = (z >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)
    got = synth.hash_u01(n, seed).numpy()
    assert np.array_equal(ref, got)
    assert 0.45 < got.mean() < 0.55


def test_empty_and_ragged_masks():
    # ragged valid lengths incl. a length-1 context; attention over the packed rows only
    cfg = dict(synth.TINY, depth=1)
    sd = synth.make_state_dict(cfg, 0)
    x, ctx, mask = synth.make_inputs(cfg, 2, 64, 64, 8, valid_len=[1, 8])
    t = torch.full((4,), 0.5)
    y = dit_oracle.dit_forward(sd, cfg, torch.cat([x, x]), ctx, mask, t)
    # padded context rows must not influence the result
    ctx2 = ctx.clone()
    ctx2[2, 1:] = 123.0
    y2 = dit_oracle.dit_forward(sd, cfg, torch.cat([x, x]), ctx2, mask, t)
    assert torch.equal(y, y2)


def _pipeline_inputs(g):
    rec = g["recipe"]
    sd, _, ctx, _, _ = build_case(rec)
    b = rec["batch"]
    pos = ctx[b:]
    lat0 = torch.randn((b, 16, g["height"] // 8, g["width"] // 8), generator=torch.Generator().manual_seed(g["seed"]))
    return rec, sd, pos, lat0


@pytest.mark.parametrize("mode", ["cfg", "apg"])
def test_sampler_and_pipeline_tail_match_the_reference_pipeline_call(mode, golden_dir):
    """tiny_256_pipeline.pt was produced by the UNMODIFIED reference FLitePipeline.__call__ (oracle/ref_pipeline_shim):
    it pins the restated loop (schedule, [negative, positive] order, zero negative embeddings, CFG / APG combine,
    accumulator), the latent unscale and the uint8 post-process."""
    from oracle import vae_decoder
    g = torch.load(os.path.join(golden_dir, "tiny_256_pipeline.pt"), weights_only=False)
    rec, sd, pos, lat0 = _pipeline_inputs(g)
    fn = lambda *a: dit_oracle.dit_forward(sd, rec["cfg"], *a)
    mask = torch.ones(2 * pos.shape[0], pos.shape[1])                # the shipped 3-argument call has no mask
    lat = sampler_oracle.sample_pipeline(fn, lat0, torch.zeros_like(pos), pos, mask, g["steps"], g["guidance"],
                                         apg=g["apg_threshold"] if mode == "apg" else None)
    z = lat / vae_decoder.SCALING_FACTOR + vae_decoder.SHIFT_FACTOR                    # pipeline.py:304
    assert _rel(z, g[f"decode_input_{mode}_fp32"]) <= 1e-5
    if mode == "cfg":
        img = ((vae_decoder.toy_decode(z) / 2 + 0.5).clamp(0, 1) * 255).round().clamp(0, 255).to(torch.uint8)
        ref = g["images_cfg_fp32"]                                   # [B, H, W, 3] from the PIL images
        diff = (img.permute(0, 2, 3, 1).int() - ref.int()).abs()
        assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3
    else:
        # APG and plain CFG really differ in the reference output (the fixture is not degenerate)
        assert _rel(g["decode_input_apg_fp32"], g["decode_input_cfg_fp32"]) > 1e-4


def test_chat_template_messages_match_the_reference(golden_dir):
    """Pipeline glue (pipeline.py:104-124): same system turn + user caption handed to the processor."""
    import flite_b200
    g = torch.load(os.path.join(golden_dir, "tiny_256_pipeline.pt"), weights_only=False)

    class Proc:
        def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True):
            self.messages, self.kw = messages, (tokenize, add_generation_prompt)
            return "x"

    proc = Proc()
    pipe = flite_b200.FLitePipeline.__new__(flite_b200.FLitePipeline)
    pipe.processor = proc
    pipe._convert_caption_to_messages("p0")
    assert proc.messages == g["chat_messages"] and proc.kw == (False, True)
