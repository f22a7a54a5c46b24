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
