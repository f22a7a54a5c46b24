"""TEST INFRASTRUCTURE, BUILD CONTAINER ONLY -- generate tests/golden/*.pt from the REAL reference.

Runs the unmodified ``/root/reference/f_lite/model.py`` (through ``oracle/ref_shim.py``) on
hash-seeded weights / inputs (``oracle/synth.py``) and stores only the *outputs* plus the
recipe (config, seeds, shapes).  Tests regenerate weights / inputs from the recipe, so the
fixtures stay small.  Re-run with:  ``python -m oracle.make_golden``
"""
from __future__ import annotations

import os

import torch

from . import ref_shim, sampler_oracle, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (cfg overrides, batch, H, W, ctx_len, valid_len, t)
    "tiny_256": (dict(), 1, 256, 256, 24, [17], 0.7),
    "tiny_rect_b2": (dict(depth=2), 2, 128, 192, 16, [16, 5], 0.31),
    "tiny_nobias": (dict(depth=1, train_bias_and_rms=False), 1, 64, 64, 8, [8], 0.95),
}


def case_recipe(name):
    over, batch, H, W, lc, valid, t = CASES[name]
    cfg = dict(synth.TINY)
    cfg.update(over)
    return dict(cfg=cfg, batch=batch, H=H, W=W, ctx_len=lc, valid_len=valid, t=t,
                weight_seed=0, input_seed=1234)


def build_case(rec, dtype=torch.float32, device="cpu"):
    cfg = rec["cfg"]
    sd = synth.make_state_dict(cfg, rec["weight_seed"], device=device)
    x, ctx, mask = synth.make_inputs(cfg, rec["batch"], rec["H"], rec["W"], rec["ctx_len"],
                                     rec["valid_len"], rec["input_seed"], device=device)
    t = torch.full((2 * rec["batch"],), rec["t"], device=device)
    return sd, torch.cat([x, x]).to(dtype), ctx.to(dtype), mask.to(dtype), t.to(dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        rec = case_recipe(name)
        out = dict(recipe=rec)
        for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            sd, x, ctx, mask, t = build_case(rec)
            ref = ref_shim.build_reference_dit(rec["cfg"], sd, dtype)
            with torch.no_grad():
                y = ref(x.to(dtype), ctx.to(dtype), mask.to(dtype), t.to(dtype))
            out["velocity_" + tag] = y.clone()
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, "std", out["velocity_fp32"].std().item(), "bf16 vs fp32 rel-L2",
              ((out["velocity_bf16"].float() - out["velocity_fp32"]).norm()
               / out["velocity_fp32"].norm()).item())

    # sampler trajectory: config C1 of BASELINE.json (tiny, 256^2, 4 Euler steps, CFG 6, B=1, fp32)
    rec = case_recipe("tiny_256")
    sd, x, ctx, mask, _ = build_case(rec)
    ref = ref_shim.build_reference_dit(rec["cfg"], sd, torch.float32)
    b = rec["batch"]
    trace = []
    lat = sampler_oracle.sample_pipeline(lambda *a: ref(*a), x[:b], ctx[:b], ctx[b:], mask, 4, 6.0,
                                         trace=trace)
    lat2 = sampler_oracle.sample_train(lambda *a: ref(*a), x[:b], ctx[:b], ctx[b:], mask[:b], mask[b:],
                                       4, 6.0)
    torch.save(dict(recipe=rec, steps=4, guidance=6.0, latents_pipeline=lat, latents_train=lat2,
                    velocities=torch.stack(trace)),
               os.path.join(OUT, "tiny_256_sampler.pt"))
    print("sampler: pipeline-vs-train rel-L2", ((lat - lat2).norm() / lat.norm()).item())

    # the reference's own FLitePipeline.__call__ end to end (pipeline.py:188-331): schedule, CFG / APG combine,
    # accumulator dtype, latent unscale, post-process -- with stand-ins for the three third-party models
    from . import ref_pipeline_shim
    rec = case_recipe("tiny_256")
    sd, _, ctx, _, _ = build_case(rec)
    out = dict(recipe=rec, seed=77, steps=4, guidance=6.0, apg_threshold=0.03, height=256, width=256)
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        for apg, atag in ((None, "cfg"), (0.03, "apg")):
            z, imgs, proc = ref_pipeline_shim.run_reference_pipeline(rec["cfg"], sd, ctx[rec["batch"]:], dtype, 77, 256, 256,
                                                                     4, 6.0, apg)
            out[f"decode_input_{atag}_{tag}"] = z
            if apg is None:
                out[f"images_{atag}_{tag}"] = imgs
            print("pipeline", tag, atag, "decode-input std", z.float().std().item(), "image mean", imgs.float().mean().item())
    out["chat_messages"] = proc.templated[-1]
    torch.save(out, os.path.join(OUT, "tiny_256_pipeline.pt"))


if __name__ == "__main__":
    main()
