"""GPU parity of the whole denoise path against the reference: golden vectors of the real module
(tests/golden, generated in the build container), the oracle restatement run on the same device, and
size-independent properties at the full 10B-architecture width."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-2          # north_star: bf16 per-step velocity relative L2 <= 1e-2 vs the reference bf16 path


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _model(cfg, sd):
    import flite_b200
    m = flite_b200.DiT(**cfg)
    m.load_state_dict(sd)
    return m.to(DEV, torch.bfloat16).eval()


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    from flite_b200 import _lib
    _lib.check(_lib.load().flite_check_device(), "flite_check_device")
    yield
    _lib.watchdog_ok()


@pytest.mark.parametrize("name", ["tiny_256", "tiny_rect_b2", "tiny_nobias"])
def test_forward_vs_reference_golden_and_oracle(name, golden_dir):
    from oracle import dit_oracle
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, t = build_case(rec, device=DEV)
    m = _model(rec["cfg"], sd)
    xb, cb, mb, tb = x.bfloat16(), ctx.bfloat16(), mask.bfloat16(), t.bfloat16()
    v = m(xb, cb, mb, tb)
    assert v.shape == xb.shape and v.dtype == torch.bfloat16
    # (1) the REAL reference module's bf16 output (CPU, committed fixture)
    r_gold = rel(v.cpu(), g["velocity_bf16"])
    # (2) the oracle restatement in bf16 on this device (cuBLAS GEMMs + fp32 attention)
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    v_or = dit_oracle.dit_forward(sdb, rec["cfg"], xb, cb, mb, tb)
    r_or = rel(v, v_or)
    # (3) both against the fp32 oracle: the new path must not be less accurate than the reference's own bf16
    sdf = {k: w.float() for k, w in sdb.items()}
    v32 = dit_oracle.dit_forward(sdf, rec["cfg"], xb.float(), cb.float(), mb.float(), tb, rope_dtype=torch.bfloat16)
    r_mine32, r_ref32 = rel(v, v32), rel(v_or, v32)
    print(f"{name}: vs golden(bf16) {r_gold:.2e} vs oracle(bf16,gpu) {r_or:.2e} | vs fp32: mine {r_mine32:.2e} ref-bf16 {r_ref32:.2e}")
    assert r_gold <= TOL and r_or <= TOL
    assert r_mine32 <= max(1.5 * r_ref32, 5e-3)


def test_legacy_three_argument_call_and_mask_none(golden_dir):
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256.pt"), weights_only=False)
    rec = dict(g["recipe"], valid_len=None)
    sd, x, ctx, mask, t = build_case(rec, device=DEV)
    m = _model(rec["cfg"], sd)
    a = m(x.bfloat16(), ctx.bfloat16(), mask.bfloat16(), t.bfloat16())
    b = m(x.bfloat16(), ctx.bfloat16(), t.bfloat16())            # pipeline.py:271 call form
    assert torch.equal(a, b)


def test_context_cache_is_invalidated_by_in_place_updates(golden_dir):
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256.pt"), weights_only=False)
    sd, x, ctx, mask, t = build_case(g["recipe"], device=DEV)
    m = _model(g["recipe"]["cfg"], sd)
    cb = ctx.bfloat16()
    a = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
    cb.mul_(0.5)                                                 # same storage, new contents
    b = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
    m.hoist_context = False
    c = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
    assert not torch.equal(a, b) and torch.equal(b, c)


def test_inference_mode_tensors_and_in_place_weight_edits(golden_dir):
    """Weights / embeddings created under torch.inference_mode() have no version counter: the forward must
    run, must not serve a stale hoisted context for them, and invalidate_caches() must pick up an in-place weight edit."""
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256.pt"), weights_only=False)
    sd, x, ctx, mask, t = build_case(g["recipe"], device=DEV)
    ref = _model(g["recipe"]["cfg"], sd)(x.bfloat16(), ctx.bfloat16(), mask.bfloat16(), t.bfloat16())
    with torch.inference_mode():
        m = _model(g["recipe"]["cfg"], {k: v.clone() for k, v in sd.items()})
        cb = ctx.bfloat16()
        a = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
        assert torch.equal(a, ref)
        cb.mul_(0.5)                                              # same storage, no version counter to notice it
        b = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
        assert not torch.equal(a, b)
        m.blocks[0].mlp.gate_proj.weight.mul_(0.5)                # e.g. a LoRA merge under inference mode
        m.blocks[0].cross_attn.context_kv.bias.add_(0.25)
        m.invalidate_caches()
        c = m(x.bfloat16(), cb, mask.bfloat16(), t.bfloat16())
        assert not torch.equal(b, c)
    m2 = _model(g["recipe"]["cfg"], sd)                           # normal tensors: version counters catch the edit
    b2 = m2(x.bfloat16(), cb.clone(), mask.bfloat16(), t.bfloat16())
    assert torch.equal(b2, b)
    with torch.no_grad():
        m2.blocks[0].mlp.gate_proj.weight.mul_(0.5)
        m2.blocks[0].cross_attn.context_kv.bias.add_(0.25)
    assert torch.equal(m2(x.bfloat16(), cb.clone(), mask.bfloat16(), t.bfloat16()), c)


def test_sampler_trajectory_vs_reference(golden_dir):
    """Config C1 of BASELINE.json: tiny DiT, 256x256, 4 Euler steps, CFG 6, batch 1."""
    import flite_b200
    from oracle import dit_oracle, sampler_oracle
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256_sampler.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    b = rec["batch"]
    m = _model(rec["cfg"], sd)
    trace = []
    lat = flite_b200.denoise(m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, g["steps"],
                             g["guidance"], trace=trace)
    # bf16 oracle trajectory on this device (the "reference bf16 path")
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    fn = lambda *a: dit_oracle.dit_forward(sdb, rec["cfg"], *a)
    otrace = []
    olat = sampler_oracle.sample_pipeline(fn, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(),
                                          mask.bfloat16(), g["steps"], g["guidance"], trace=otrace)
    v0 = trace[0][:b] + g["guidance"] * (trace[0][b:] - trace[0][:b])     # my trace holds [uncond; cond]
    gv0 = g["velocities"][0].to(DEV)                             # fp32 reference, CFG-combined (x6 amplification)
    assert rel(v0, gv0) <= max(1.5 * rel(otrace[0], gv0), 2 * TOL)
    r_final = rel(lat, olat)
    r_gold = rel(lat.cpu(), g["latents_pipeline"])               # fp32 reference trajectory (CPU fixture)
    r_ref_gold = rel(olat.cpu(), g["latents_pipeline"])
    print(f"final latents: vs oracle bf16 {r_final:.2e}; vs fp32 reference: mine {r_gold:.2e}, ref-bf16 {r_ref_gold:.2e}")
    assert r_final <= 3e-2 and r_gold <= max(1.5 * r_ref_gold, 2e-2)
    # fp32-accumulator semantics of train.py::sample_images
    lat32 = flite_b200.denoise(m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, g["steps"],
                               g["guidance"], acc_dtype=torch.float32)
    assert lat32.dtype == torch.float32 and rel(lat32.cpu(), g["latents_train"]) <= max(1.5 * r_ref_gold, 2e-2)


def test_pipeline_call_latent_output(golden_dir):
    import flite_b200
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    m = _model(rec["cfg"], sd)
    pipe = flite_b200.FLitePipeline(m, None, None, None)
    gen = torch.Generator(device=DEV).manual_seed(3)
    out = pipe(prompt=None, height=256, width=256, num_inference_steps=2, guidance_scale=6.0, generator=gen,
               prompt_embeds=ctx[1:].bfloat16(), prompt_attention_mask=mask[1:], output_type="latent")
    assert out.images.shape == (1, 16, 32, 32) and torch.isfinite(out.images.float()).all()


def test_wide_model_properties_and_oracle():
    """10B-architecture WIDTH (d 3072, 12 heads, 1024^2 => 2 x 4112 tokens) at depth 2 so the oracle finishes in
    seconds; plus properties that hold at any size: batched-CFG == two separate calls (pipeline.py:264-271 vs
    train.py:591-595) and batch-permutation equivariance."""
    from oracle import dit_oracle, synth
    cfg = dict(synth.ARCH_10B, depth=2)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    x, ctx, mask = synth.make_inputs(cfg, 1, 1024, 1024, 256, valid_len=[200], device=DEV)
    xb, cb, mb = torch.cat([x, x]).bfloat16(), ctx.bfloat16(), mask.bfloat16()
    t = torch.tensor([0.81, 0.81], device=DEV).bfloat16()
    v = m(xb, cb, mb, t)
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    v_or = dit_oracle.dit_forward(sdb, cfg, xb, cb, mb, t)
    r = rel(v, v_or)
    print("wide depth-2 vs oracle bf16:", r, "std", v_or.float().std().item())
    assert r <= TOL
    v0 = m(xb[:1], cb[:1], mb[:1], t[:1])
    v1 = m(xb[1:], cb[1:], mb[1:], t[1:])
    assert rel(torch.cat([v0, v1]), v) <= 2e-3
    vp = m(xb.flip(0), cb.flip(0), mb.flip(0), t.flip(0))
    assert rel(vp.flip(0), v) <= 2e-3


def test_full_size_c2_velocity_parity():
    """BASELINE.json configs[1] at FULL size: 10B architecture (depth 40, d 3072, 12 heads), 1024x1024, CFG batch of 2.
    north_star tolerance: bf16 per-step velocity relative L2 <= 1e-2 against the reference bf16 path (here the oracle
    restatement of the reference, bit-equal to the real module on CPU, run in bf16 on this GPU with cuBLAS GEMMs and
    fp32 softmax attention)."""
    from oracle import dit_oracle, synth
    cfg = dict(synth.ARCH_10B)
    sd = synth.make_state_dict(cfg, 0, device=DEV, dtype=torch.bfloat16)
    import flite_b200
    m = flite_b200.DiT(**cfg).to(torch.bfloat16)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x, ctx, mask = synth.make_inputs(cfg, 1, 1024, 1024, 256, valid_len=[200], device=DEV)
    xb, cb, mb = torch.cat([x, x]).bfloat16(), ctx.bfloat16(), mask.bfloat16()
    for tval in (0.95, 0.3):
        t = torch.tensor([tval, tval], device=DEV).bfloat16()
        v = m(xb, cb, mb, t)
        v_or = dit_oracle.dit_forward(sd, cfg, xb, cb, mb, t)
        r = rel(v, v_or)
        print(f"C2 full size, t={tval}: rel-L2 vs reference bf16 path {r:.3e} (output std {v_or.float().std().item():.2f})")
        assert v_or.float().std().item() > 0.1 and r <= TOL
    del m, sd
    torch.cuda.empty_cache()


def test_decoded_image_psnr_vs_reference_trajectory(golden_dir):
    """north_star image criterion: final decoded image PSNR >= 40 dB against the reference image.  Config C1 (tiny DiT,
    256x256, 4 Euler steps, CFG 6): the new path's trajectory and the reference bf16 trajectory (oracle) are decoded by the
    SAME FLUX-style decoder (random init, oracle/vae_decoder.py) and post-processed like f_lite/pipeline.py:324-326."""
    import flite_b200
    from oracle import dit_oracle, sampler_oracle, vae_decoder
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256_sampler.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    b = rec["batch"]
    m = _model(rec["cfg"], sd)
    lat = flite_b200.denoise(m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, g["steps"], g["guidance"])
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    fn = lambda *a: dit_oracle.dit_forward(sdb, rec["cfg"], *a)
    olat = sampler_oracle.sample_pipeline(fn, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask.bfloat16(),
                                          g["steps"], g["guidance"])
    dec = vae_decoder.make_decoder(0, DEV)
    img, ref = vae_decoder.decode_to_image(dec, lat), vae_decoder.decode_to_image(dec, olat)
    ref32 = vae_decoder.decode_to_image(dec, g["latents_pipeline"].to(DEV))          # fp32 reference trajectory (CPU fixture)
    p = vae_decoder.psnr(img, ref)
    print(f"PSNR vs reference bf16 image {p:.1f} dB | vs fp32 reference image: mine {vae_decoder.psnr(img, ref32):.1f} dB, "
          f"reference-bf16 {vae_decoder.psnr(ref, ref32):.1f} dB | image std {ref.std().item():.2f}")
    assert ref.std().item() > 0.1           # the decoded image is not flat
    assert p >= 40.0


def test_sampler_apg_trajectory_vs_reference(golden_dir):
    """Augmented Parallel Guidance (pipeline.py:276-287) through flite_b200.denoise vs the reference loop (oracle) on
    the same device: 4 steps, tiny DiT.  APG's global scalars are bf16, so one-ulp flips are possible: rel-L2 <= 3e-2
    on the final latents (same bound as the plain-CFG trajectory test)."""
    import flite_b200
    from oracle import dit_oracle, sampler_oracle
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256_sampler.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    b = rec["batch"]
    m = _model(rec["cfg"], sd)
    apg = flite_b200.APGConfig(enabled=True, orthogonal_threshold=0.03)
    lat = flite_b200.denoise(m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, g["steps"],
                             g["guidance"], apg_config=apg)
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    fn = lambda *a: dit_oracle.dit_forward(sdb, rec["cfg"], *a)
    olat = sampler_oracle.sample_pipeline(fn, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(),
                                          mask.bfloat16(), g["steps"], g["guidance"], apg=0.03)
    plain = flite_b200.denoise(m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, g["steps"],
                               g["guidance"])
    r = rel(lat, olat)
    x0 = x[:b].bfloat16().float()
    d_mine, d_ref, d_plain = lat.float() - x0, olat.float() - x0, plain.float() - x0     # displacements over the 4 steps
    print(f"APG final latents vs oracle bf16: {r:.2e}; displacement: mine vs oracle {rel(d_mine, d_ref):.2e}, "
          f"plain CFG vs oracle-APG {rel(d_plain, d_ref):.2e}")
    # (cond and uncond outputs of this tiny random model nearly coincide, so APG ~ CFG ~ dy here; that the kernel
    # applies APG and not CFG is asserted on synthetic velocities in test_kernels_gpu.py::test_apg_euler_*)
    assert r <= 3e-2 and rel(d_mine, d_ref) <= 5e-2


def test_pipeline_call_decodes_to_uint8_images(golden_dir):
    """FLitePipeline.__call__ end to end with caller-supplied embeddings and a VAE module (pipeline.py:299-327):
    latent unscale + uint8 post-process run on the flite kernels; output equals the reference's torch op sequence
    applied to the same decoder output."""
    import flite_b200
    from types import SimpleNamespace
    from oracle import vae_decoder
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    m = _model(rec["cfg"], sd)
    dec = vae_decoder.make_decoder(0, DEV).to(torch.bfloat16)

    class VAE:
        config = SimpleNamespace(scaling_factor=vae_decoder.SCALING_FACTOR, shift_factor=vae_decoder.SHIFT_FACTOR)
        dtype = torch.bfloat16
        seen = []

        def decode(self, z):
            self.seen.append(z)
            return SimpleNamespace(sample=dec(z))

    vae = VAE()
    pipe = flite_b200.FLitePipeline(m, vae, None, None)
    kw = dict(prompt=None, height=256, width=256, num_inference_steps=2, guidance_scale=6.0,
              prompt_embeds=ctx[1:].bfloat16(), prompt_attention_mask=mask[1:])
    lat = pipe(generator=torch.Generator(device=DEV).manual_seed(3), output_type="latent", **kw).images
    out = pipe(generator=torch.Generator(device=DEV).manual_seed(3), output_type="pt", **kw).images
    assert out.dtype == torch.uint8 and out.shape == (1, 3, 256, 256)
    z = lat / vae_decoder.SCALING_FACTOR + vae_decoder.SHIFT_FACTOR                      # pipeline.py:304
    assert torch.equal(vae.seen[-1], z)
    ref = ((dec(z) / 2 + 0.5).clamp(0, 1) * 255).round().clamp(0, 255).to(torch.uint8).cpu()   # pipeline.py:324-326
    assert torch.equal(out, ref)
    pil = pipe(generator=torch.Generator(device=DEV).manual_seed(3), **kw).images
    assert pil[0].size == (256, 256) and pil[0].mode == "RGB"


def test_programmatic_dependent_launch_is_bit_identical():
    """FLITE_TUNE_PDL = 1 (opt-in): the GEMM / attention / rmsnorm kernels start under programmatic dependent launch
    (their prologue overlaps the previous kernel's tail).  Several denoise steps at the 10B width with changing
    timesteps must give exactly the bits of the fully serialised launches (stale-read / ordering check)."""
    import flite_b200
    from flite_b200 import _lib
    from oracle import synth
    cfg = dict(synth.ARCH_10B, depth=3)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    m.hoist_context = False
    x, ctx, mask = synth.make_inputs(cfg, 1, 1024, 1024, 256, valid_len=[200], device=DEV)
    lat0, neg, pos = x.bfloat16(), ctx[:1].bfloat16(), ctx[1:].bfloat16()
    lib = _lib.load()
    outs = []
    for on in (0, 1, 1):
        lib.flite_set_tuning(8, on)
        try:
            outs.append(flite_b200.denoise(m, lat0, neg, pos, mask, 5, 6.0))
        finally:
            lib.flite_set_tuning(8, 0)
    _lib.watchdog_ok()
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("mode", ["cfg", "apg"])
def test_pipeline_call_vs_reference_pipeline_golden(mode, golden_dir):
    """flite_b200.FLitePipeline.__call__ against the fixture written by the UNMODIFIED reference FLitePipeline.__call__
    (tests/golden/tiny_256_pipeline.pt): same CPU generator seed, same embeddings, toy VAE.  Checks the decode input
    (trajectory + latent unscale) against the reference's bf16 run and the uint8 image."""
    import flite_b200
    from types import SimpleNamespace
    from oracle import vae_decoder
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256_pipeline.pt"), weights_only=False)
    rec = g["recipe"]
    sd, _, ctx, _, _ = build_case(rec, device=DEV)
    m = _model(rec["cfg"], sd)

    class VAE:
        config = SimpleNamespace(scaling_factor=vae_decoder.SCALING_FACTOR, shift_factor=vae_decoder.SHIFT_FACTOR)
        dtype = torch.bfloat16
        seen = []

        def decode(self, z):
            self.seen.append(z)
            return SimpleNamespace(sample=vae_decoder.toy_decode(z))

    vae = VAE()
    pipe = flite_b200.FLitePipeline(m, vae, None, None)
    apg = flite_b200.APGConfig(enabled=True, orthogonal_threshold=g["apg_threshold"]) if mode == "apg" else None
    out = pipe(prompt=None, height=g["height"], width=g["width"], num_inference_steps=g["steps"],
               guidance_scale=g["guidance"], generator=torch.Generator().manual_seed(g["seed"]), apg_config=apg,
               prompt_embeds=ctx[rec["batch"]:].bfloat16(), output_type="pt").images
    z = vae.seen[-1].float().cpu()
    # (the fixture's fp32 run starts from different noise -- randn in fp32 vs bf16 -- so only the bf16 run is comparable)
    r_bf16 = rel(z, g[f"decode_input_{mode}_bf16"])
    print(f"{mode}: decode input vs the reference pipeline's bf16 run {r_bf16:.2e}")
    assert r_bf16 <= 3e-2
    if mode == "cfg":
        ref_img = g["images_cfg_bf16"].permute(0, 3, 1, 2)
        diff = (out.int() - ref_img.int()).abs().float()
        print("uint8 image: mean |diff|", diff.mean().item(), "max", diff.max().item())
        assert diff.mean().item() < 1.0


def _poison_allocator(nbytes=6 << 30):
    """Leave NaN bit patterns in the caching allocator's free blocks so that every later torch.empty() starts as NaN."""
    torch.cuda.empty_cache()
    junk = [torch.full((nbytes // 8,), -1, dtype=torch.int16, device=DEV) for _ in range(4)]   # 0xFFFF = bf16 / fp32 NaN
    del junk


@pytest.mark.parametrize("hoist", [True, False])
def test_no_uninitialised_reads_with_poisoned_workspaces(hoist):
    """Every buffer the path allocates with torch.empty (activation workspaces, per-call outputs) is poisoned with NaN
    before the run: the result must be finite and bit-identical to a clean run.  10B width, 2 images (4 sequences,
    T = 16448: multi-band GEMM rasterisation and tail units), ragged context, 3 steps."""
    import flite_b200
    from oracle import synth
    cfg = dict(synth.ARCH_10B, depth=2)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    m.hoist_context = hoist
    x, ctx, mask = synth.make_inputs(cfg, 2, 1024, 1024, 256, valid_len=[200, 77], device=DEV)
    lat0, neg, pos = x.bfloat16(), ctx[:2].bfloat16(), ctx[2:].bfloat16()
    clean = flite_b200.denoise(m, lat0, neg, pos, mask, 3, 6.0)
    assert torch.isfinite(clean.float()).all()
    for buf in m._ws.values():
        if buf.is_floating_point():
            buf.fill_(float("nan"))
    m._ctx_cache = None
    _poison_allocator()
    again = flite_b200.denoise(m, lat0, neg, pos, mask, 3, 6.0)
    assert torch.isfinite(again.float()).all()
    assert torch.equal(again, clean)


@pytest.mark.parametrize("mode", ["cfg", "apg", "nocfg", "fp32acc"])
def test_cuda_graph_replay_is_bit_identical(mode, golden_dir):
    """denoise(cuda_graph=True): the DiT forward captured once and replayed every step (graphs.GraphedForward) gives
    exactly the eager loop's bits -- tiny model (the launch-bound case it exists for), 4 steps."""
    import flite_b200
    from oracle.make_golden import build_case
    g = torch.load(os.path.join(golden_dir, "tiny_256_sampler.pt"), weights_only=False)
    rec = g["recipe"]
    sd, x, ctx, mask, _ = build_case(rec, device=DEV)
    b = rec["batch"]
    m = _model(rec["cfg"], sd)
    kw = dict(num_inference_steps=4, guidance_scale=0.5 if mode == "nocfg" else 6.0,
              apg_config=flite_b200.APGConfig(True, 0.03) if mode == "apg" else None,
              acc_dtype=torch.float32 if mode == "fp32acc" else torch.bfloat16)
    args = (m, x[:b].bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask)
    eager = flite_b200.denoise(*args, **kw)
    graphed = flite_b200.denoise(*args, cuda_graph=True, **kw)
    assert torch.isfinite(eager.float()).all() and torch.equal(eager, graphed)


def test_cuda_graph_replay_wide_model_without_context_hoisting():
    import flite_b200
    from oracle import synth
    cfg = dict(synth.ARCH_10B, depth=2)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    m.hoist_context = False                      # the context path is then captured inside the graph as well
    x, ctx, mask = synth.make_inputs(cfg, 1, 1024, 1024, 256, valid_len=[200], device=DEV)
    args = (m, x.bfloat16(), ctx[:1].bfloat16(), ctx[1:].bfloat16(), mask)
    eager = flite_b200.denoise(*args, num_inference_steps=3)
    graphed = flite_b200.denoise(*args, num_inference_steps=3, cuda_graph=True)
    assert torch.equal(eager, graphed)


@pytest.mark.parametrize("name,H,W,batch,valid,depth", [
    ("C3 shape: 1344x896 (rectangular RoPE grid 84x56, L = 4720), 2 prompts", 1344, 896, 2, [256, 131], 2),
    ("C4 shape: 2048x2048 (L = 16400), 1 prompt", 2048, 2048, 1, [256], 1),
])
def test_baseline_config_shapes_vs_oracle(name, H, W, batch, valid, depth):
    """BASELINE.json configs[2] / [3] at their full width, resolution and sequence length (depth reduced so the
    torch oracle -- which materialises the attention scores -- finishes in seconds): velocity rel-L2 <= 1e-2 against the
    reference bf16 op sequence on the same device, batched-CFG layout [negative, positive]."""
    from oracle import dit_oracle, synth
    cfg = dict(synth.ARCH_10B, depth=depth)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    x, ctx, mask = synth.make_inputs(cfg, batch, H, W, 256, valid_len=valid, device=DEV)
    xb, cb, mb = torch.cat([x, x]).bfloat16(), ctx.bfloat16(), mask.bfloat16()
    t = torch.full((2 * batch,), 0.43, device=DEV).bfloat16()
    v = m(xb, cb, mb, t)
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    v_or = dit_oracle.dit_forward(sdb, cfg, xb, cb, mb, t)
    r = rel(v, v_or)
    print(f"{name}: rel-L2 vs oracle bf16 {r:.2e}, output std {v_or.float().std().item():.3f}")
    assert v.shape == xb.shape and r <= TOL
    del m, sd, sdb
    torch.cuda.empty_cache()


def test_all_masked_context_rows_give_zero_cross_attention():
    """Edge case of the varlen path (model.py:31-64,190-210): a sample whose context mask is all zeros has NO keys;
    flash-attn returns zeros there, so its output must equal the output of a model without cross-attention
    contribution for that sample, and must not disturb the other sample."""
    from oracle import dit_oracle, synth
    cfg = dict(synth.TINY, depth=2)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    x, ctx, mask = synth.make_inputs(cfg, 1, 128, 128, 16, valid_len=[9], device=DEV)
    mask = mask.clone()
    mask[0] = 0                                   # the negative row: no valid context token at all
    xb, cb, mb = torch.cat([x, x]).bfloat16(), ctx.bfloat16(), mask.bfloat16()
    t = torch.full((2,), 0.6, device=DEV).bfloat16()
    v = m(xb, cb, mb, t)
    sdb = {k: w.bfloat16() for k, w in sd.items()}
    v_or = dit_oracle.dit_forward(sdb, cfg, xb, cb, mb, t)
    assert torch.isfinite(v.float()).all() and rel(v, v_or) <= TOL
    cb2 = cb.clone(); cb2[0] = 77.0               # fully masked rows are never read
    assert torch.equal(m(xb, cb2, mb, t), v)


def test_product_vae_decoder_bf16_channels_last_vs_fp32_restatement():
    """flite_b200.vae.AutoencoderKL (diffusers parameter names, bf16, channels_last, cuDNN) against the fp32 oracle
    decoder with the same weights: decoded image PSNR, then the flite pipeline tail on top of it."""
    from flite_b200 import ops, vae
    from oracle import vae_decoder
    torch.manual_seed(0)
    m = vae.AutoencoderKL().to(DEV)
    for p_ in m.parameters():
        if p_.dim() > 1:
            p_.data.uniform_(-1, 1).mul_((3.0 / p_.shape[1:].numel()) ** 0.5)
    ref = vae_decoder.Decoder().to(DEV).eval()
    ref.load_state_dict(vae_decoder.map_diffusers_names(m.state_dict()), strict=True)
    mb = vae.AutoencoderKL().to(DEV)
    mb.load_state_dict(m.state_dict())
    mb = mb.to(torch.bfloat16).to(memory_format=torch.channels_last).eval()
    lat = torch.randn(2, 16, 32, 32, device=DEV).bfloat16()
    z = ops.latent_unscale(lat, mb.config.scaling_factor, mb.config.shift_factor)
    img = ops.image_to_uint8(mb.decode(z).sample)                                  # [B, H, W, 3] uint8
    with torch.no_grad():
        x32 = ref(z.float())
    ref_img = ((x32 / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1)
    mse = ((img.float() - ref_img.float()) / 255).pow(2).mean().item()
    psnr = 10 * math.log10(1.0 / max(mse, 1e-12))
    print(f"bf16 product decoder vs fp32 restatement: PSNR {psnr:.1f} dB, image std {ref_img.float().std().item():.1f}")
    assert img.shape == (2, 256, 256, 3) and ref_img.float().std().item() > 10 and psnr >= 30.0


def test_batch4_per_gpu_layout_is_bit_identical_to_per_image_runs():
    """C3 / C5 layout (>= 4 images = 8 CFG sequences per GPU): the default path is batch-invariant -- every
    row of every GEMM / norm and every (sequence, head) of attention is computed the same way whatever else is in the
    batch -- so denoising 4 images together gives exactly the bits of 4 single-image runs (10B width, depth 2, 512^2),
    and the stream-K attention (opt-in) stays within the velocity tolerance of the default path."""
    import flite_b200
    from oracle import synth
    cfg = dict(synth.ARCH_10B, depth=2)
    sd = synth.make_state_dict(cfg, 0, device=DEV)
    m = _model(cfg, sd)
    g = torch.Generator(device=DEV).manual_seed(5)
    b = 4
    lat = torch.randn(b, 16, 64, 64, device=DEV, generator=g).bfloat16()
    pos = torch.randn(b, 64, 4096, device=DEV, generator=g).bfloat16()
    neg = torch.zeros_like(pos)
    mask = torch.ones(2 * b, 64, device=DEV)
    mask[b + 1, 40:] = 0                                     # one ragged prompt
    together = flite_b200.denoise(m, lat, neg, pos, mask, 2, 6.0)
    for i in range(b):
        mi = torch.cat([mask[i:i + 1], mask[b + i:b + i + 1]])
        alone = flite_b200.denoise(m, lat[i:i + 1], neg[i:i + 1], pos[i:i + 1], mi, 2, 6.0)
        assert torch.equal(alone[0], together[i]), f"image {i} differs between the batched and the single-image run"
    t = torch.full((2 * b,), 0.6, device=DEV).bfloat16()
    x2, c2 = torch.cat([lat, lat]), torch.cat([neg, pos])
    v0 = m(x2, c2, mask, t)
    default = m.attn_streamk
    try:
        m.attn_streamk = "0"                 # one cluster per unit: the round-robin default must reproduce it bit for bit
        assert torch.equal(m(x2, c2, mask, t), v0)
        for mode in ("1", "hybrid"):         # stream-K shares: split units are merged in fp32
            m.attn_streamk = mode
            assert rel(m(x2, c2, mask, t), v0) <= TOL
    finally:
        m.attn_streamk = default
    assert default == "rr"
