#!/usr/bin/env python
"""bench.py -- F Lite denoise steps/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]

A "step" is one denoise step of the reference sampler: one CFG-batched DiT forward on [negative, positive]
(f_lite/pipeline.py:264-271 with the 4-argument forward of f_lite/model.py:526) + CFG combine + Euler update
(pipeline.py:290,296-297).  Workload at N=1 is BASELINE.json configs[1] ("C2"): the 10B architecture as the mounted
model.py instantiates it (d 3072, depth 40, 12 heads, 6.84 B params), 1024x1024, batch 1 => 2 sequences of 4112 tokens,
text context 256 tokens of width 4096, bf16, synthetic latents/embeddings, random-init (de-zeroed) weights.
For N>1 every rank runs its own image (data parallel over prompts, no collective on the data path): weak scaling,
value = N images' steps per second.

Our arm (default).  Printed JSON keys follow the driver contract: value (inputs resident in HBM), e2e (host buffers,
H2D/D2H inside the timed region, through flite_b200.denoise_step = the public API), roofline (dominant kernel = MLP
gate/up tcgen05 GEMM, CUDA events on the launching stream during the timed region; traffic read from the newest
profiles/*ncu_full*.csv), cpu_baseline (N=1 only; bounded sample of the same workload on the host cores), clocks,
gpu_launches.  Extra keys: `c1` (BASELINE.json configs[0] run in full on this GPU: 4 Euler steps of the tiny DiT) and,
for N>1, `multi_gpu` -- the COMMUNICATING layouts of SURVEY.md 8(e) run after the data-parallel timing with the same
weights: CFG halves on two GPUs (C2) and Ulysses sequence parallelism with the exchange fused into the kernels over
NVLink peer memory (C4, one 2048^2 image), each with ms/step (max over ranks), speed-up over the 1-GPU step measured in
the same process, and rel-L2 of the step's result against the 1-GPU path (0.0 = bit-identical).

Reference arm (--impl reference; rank 0 only).  The UNMODIFIED reference module (oracle/_ref/f_lite/model.py, installed
by oracle/build_ref.py) on the host cores: `value` = C2 steps/s from a bounded sample per step (one cross-attention
DiTBlock + one plain DiTBlock of the reference's own classes at the full C2 token count, scaled to the 16 + 24 block
mix; ms_per_step is the time actually measured per sample), plus `c1_full` (configs[0] run in full, nothing
extrapolated) and `reference_gpu` (the same unmodified module in bf16 on the B200 with the real liger_kernel and
flash-attn 2 kernels at C2 -- what the reference itself would run on this box).
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import math
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "10B DiT denoise steps/s @1024^2 bf16"
UNIT = "steps/s"

ARCH_10B = dict(in_channels=16, patch_size=2, hidden_size=3072, depth=40, num_heads=12, mlp_ratio=4.0,
                cross_attn_input_size=4096, train_bias_and_rms=True, use_rope=True)
ARCH_TINY = dict(ARCH_10B, hidden_size=512, depth=4, num_heads=2)
ARCH_7B = dict(ARCH_10B, depth=28)     # ASSUMED: "7B" is not defined anywhere in the reference
WORKLOADS = {
    # name: (arch, height, width, ctx_len, images per GPU)   -- BASELINE.json configs[0..4]
    "c2": (ARCH_10B, 1024, 1024, 256, 1),
    "c1": (ARCH_TINY, 256, 256, 256, 1),
    "c3": (ARCH_7B, 1344, 896, 256, 8),     # 64 prompts over 8 GPUs = 8 images (16 CFG sequences) per GPU
    "c4": (ARCH_10B, 2048, 2048, 256, 1),   # single 2048^2 image (1-GPU form; Ulysses form: `multi_gpu` key at N>1)
    "c5": (ARCH_10B, 1024, 1024, 256, 4),   # batch 32 over 8 GPUs = 4 images per GPU (with decode: tools/pipeline_bench.py)
}
GUIDANCE = 6.0


def flops_per_step(cfg, height, width, ctx_len, images):
    """Algorithmic FLOPs of one CFG-batched step: 2MNK per GEMM, 4 Lq Lk d per attention."""
    d, depth, ci, p = cfg["hidden_size"], cfg["depth"], cfg["cross_attn_input_size"], cfg["patch_size"]
    B = 2 * images
    L = 16 + (height // 8 // p) * (width // 8 // p)
    T, Tc = B * L, B * ctx_len
    X = len([i for i in range(depth) if i % 4 == 0 or i < 8])
    f = depth * (2 * T * d * 3 * d + 2 * T * d * d + B * 4 * L * L * d + 3 * 2 * T * d * 4 * d)
    f += X * (2 * 2 * T * d * d + 2 * Tc * d * 2 * d + B * 4 * L * ctx_len * d)
    f += 2 * Tc * ci * d + 2 * 2 * B * (L - 16) * 64 * d + 2 * B * (d * 4 * d * 2 + d * 9 * d + d * 2 * d)
    return float(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            return dict(bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                        hbm=j["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def workload_desc(name, cfg, height, width, ctx_len, images, n_gpus):
    L = 16 + (height // 16) * (width // 16)
    return (f"{name.upper()}: F Lite DiT d{cfg['hidden_size']} depth{cfg['depth']} heads{cfg['num_heads']} "
            f"{height}x{width}, CFG-batched [neg,pos] => {2 * images} seq x {L} tokens per GPU, ctx {ctx_len}x"
            f"{cfg['cross_attn_input_size']}, {images} image(s)/GPU x {n_gpus} GPU(s)")


def bench_config(name, cfg, height, width, ctx_len, images, n_gpus):
    """The `config` object -- built by ONE function so both arms print the same thing for the same flags."""
    return {"workload": workload_desc(name, cfg, height, width, ctx_len, images, n_gpus),
            "parallelism": f"dp{n_gpus} (one image set per GPU, no data-path collective)",
            "weights": "random-init, de-zeroed (seed 0), replicated per GPU",
            "context_kv": "recomputed every step (hoisting disabled)",
            "l2": "GBs of weights streamed per step >> 126 MB L2, no flush needed",
            "flops_per_step_per_gpu": flops_per_step(cfg, height, width, ctx_len, images)}


def ncu_traffic(kernel_regex, grid_hint=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the first kernel whose name matches `kernel_regex`,
    from the newest (by its per-round name) profiles/*ncu_full*.csv (an `ncu -i ... --page raw --csv` export).  Returns (bytes, file) or
    (None, None) -- never a hard-coded number."""
    # files are named per round (r1f_..., r2a_...): the lexicographically last one is the newest capture
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full*.csv")), key=os.path.basename, reverse=True)
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in files:
        try:
            rows = list(csv.reader(open(path, newline="")))
        except Exception:
            continue
        hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
        if hdr is None or hdr + 2 >= len(rows):
            continue
        names, units = rows[hdr], rows[hdr + 1]
        try:
            kn, rd, wr = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
        except ValueError:
            continue
        for r in rows[hdr + 2:]:
            if len(r) <= max(kn, rd, wr) or not re.search(kernel_regex, r[kn]):
                continue
            if grid_hint is not None and grid_hint not in ",".join(r):
                continue
            try:
                total = (float(r[rd].replace(",", "")) * unit_scale.get(units[rd], 1.0)
                         + float(r[wr].replace(",", "")) * unit_scale.get(units[wr], 1.0))
            except ValueError:
                continue
            return total, os.path.relpath(path, ROOT)
    return None, None


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference module on the host cores
# --------------------------------------------------------------------------------------------------------------
def _dezero_(module, seed):
    """Seeded de-zeroed default init for a reference module (f_lite/model.py zero-inits adaLN / final layers, D10)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for name, p in module.named_parameters():
        if p.abs().sum().item() == 0:
            p.data.copy_(torch.randn(p.shape, generator=g) * 0.02)


def cpu_reference_sample(cfg, height, width, ctx_len, images, repeats=1):
    """Bounded sample of one C-workload step on the host cores with the reference's OWN classes: one
    `DiTBlock(do_cross_attn=True)` + one `DiTBlock(do_cross_attn=False)` of the unmodified f_lite/model.py (fp32, all
    host threads; LigerRMSNorm / LigerSwiGLUMLP / flash_attn_varlen_func are Triton / CUDA-only, so on the CPU they are
    the torch restatements of oracle/dit_oracle.py) at the workload's full token count, scaled to the step's block
    mix.  The embedders and the final head (< 0.1 % of the FLOPs) are not sampled.
    Returns (extrapolated step seconds per repeat, measured sample seconds per repeat, cores, kind, description)."""
    import torch

    from oracle import build_ref, dit_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d, nh, depth, p = cfg["hidden_size"], cfg["num_heads"], cfg["depth"], cfg["patch_size"]
    B = 2 * images
    h, w = height // 8 // p, width // 8 // p
    L = 16 + h * w
    n_cross = len([i for i in range(depth) if i % 4 == 0 or i < 8])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B * L, d, generator=g)
    ctx = torch.randn(B * ctx_len, d, generator=g)
    cu_x = torch.arange(B + 1, dtype=torch.int32) * L
    cu_c = torch.arange(B + 1, dtype=torch.int32) * ctx_len
    mod = tuple(0.1 * torch.randn(B, d, generator=g).repeat_interleave(L, 0) for _ in range(9))
    cos, sin = dit_oracle.rope_tables(d // nh, h, w, 10000, "cpu", torch.float32)
    rope = (cos.repeat(1, B, 1), sin.repeat(1, B, 1))
    if build_ref.ref_path("f_lite/model.py") is not None:
        from oracle import ref_shim
        m = ref_shim.load_reference_model_module("cpu")
        torch.manual_seed(0)
        blk_x = m.DiTBlock(d, nh, do_cross_attn=True, mlp_ratio=cfg["mlp_ratio"], qkv_bias=True).eval()
        blk_p = m.DiTBlock(d, nh, do_cross_attn=False, mlp_ratio=cfg["mlp_ratio"], qkv_bias=True).eval()
        run_x = lambda: blk_x(x, cu_x, L, ctx, cu_c, ctx_len, mod, rope)
        run_p = lambda: blk_p(x, cu_x, L, ctx, cu_c, ctx_len, mod, rope)
        kind, what = "reference", "f_lite/model.py::DiTBlock (unmodified, oracle/_ref)"
    else:   # oracle/_ref not installed: the port
        from oracle import synth
        scfg = dict(synth.TINY, **{k: cfg[k] for k in cfg})
        scfg["depth"] = 2
        shapes = {k: v for k, v in synth.param_shapes(scfg).items() if k.startswith("blocks.")}
        sd = {}
        for k, shp in shapes.items():
            if "norm" in k:
                sd[k] = torch.ones(shp)
            else:
                fan_in = shp[1] if len(shp) > 1 else d
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        run_x = lambda: dit_oracle.dit_block(sd, 0, x, cu_x, ctx, cu_c, mod, rope, nh, True, dit_oracle.flash_attn_varlen)
        run_p = lambda: dit_oracle.dit_block(sd, 1, x, cu_x, ctx, cu_c, mod, rope, nh, False, dit_oracle.flash_attn_varlen)
        kind, what = "port", "oracle/dit_oracle.py::dit_block (oracle/_ref not installed)"
    full, sample = [], []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            run_x()
            t1 = time.perf_counter()
            run_p()
            t2 = time.perf_counter()
            full.append(n_cross * (t1 - t0) + (depth - n_cross) * (t2 - t1))
            sample.append(t2 - t0)
    desc = (f"per step: 1 cross-attention block ({t1 - t0:.2f} s) + 1 plain block ({t2 - t1:.2f} s) of {what} at "
            f"{B}x{L} tokens, fp32 torch on {cores} threads; full-step time = {n_cross}*cross + {depth - n_cross}*plain")
    return full, sample, cores, kind, desc


def cpu_reference_c1_full(steps=4, repeats=3):
    """BASELINE.json configs[0] in FULL, nothing extrapolated: the unmodified reference DiT (tiny: d 512, depth 4),
    256x256, 4 Euler steps, CFG 6 (batched [neg, pos]), batch 1, fp32 on all host cores, driven by the reference's
    sampler loop as restated in oracle/sampler_oracle.py (f_lite/pipeline.py:244-297)."""
    import torch

    from oracle import build_ref, ref_shim, sampler_oracle
    if build_ref.ref_path("f_lite/model.py") is None:
        return {"unavailable": "oracle/_ref not installed"}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    cfg = {k: v for k, v in ARCH_TINY.items()}
    model = ref_shim.build_reference_dit(cfg, None, torch.float32, backend="cpu")
    _dezero_(model, 0)
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(1, 16, 32, 32, generator=g)
    pos = torch.randn(1, 256, 4096, generator=g)
    neg = torch.zeros_like(pos)
    mask = torch.ones(2, 256)
    fn = lambda *a: model(*a)
    times = []
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        out = sampler_oracle.sample_pipeline(fn, lat, neg, pos, mask, steps, GUIDANCE)
        times.append((time.perf_counter() - t0) / steps)
    best = min(times[1:])
    return {"config": "C1: tiny DiT d512 depth4 heads2, 256x256, 4 Euler steps, CFG 6, batch 1, fp32, unmodified "
                      "f_lite/model.py (oracle/_ref) through the reference sampler loop",
            "ms_per_step": best * 1e3, "steps_per_s": 1.0 / best, "cores": cores, "kind": "reference",
            "finite": bool(torch.isfinite(out).all())}


def reference_gpu_c2(cfg, height, width, ctx_len, images, steps, warmup=2):
    """The unmodified reference module on the B200 in bf16 with the REAL third-party kernels (liger_kernel Triton
    RMSNorm / SwiGLU, FlashAttention-2 behind the flash_attn_interface name) -- what the reference runs on this box.
    One step = CFG-batched forward + the reference's own torch CFG / Euler ops (pipeline.py:290,296-297)."""
    import torch
    if not torch.cuda.is_available():
        return {"unavailable": "no GPU visible"}
    try:
        from oracle import build_ref, ref_shim
        if build_ref.ref_path("f_lite/model.py") is None:
            return {"unavailable": "oracle/_ref not installed"}
        import flash_attn
        import liger_kernel
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        torch.manual_seed(0)
        model = ref_shim.build_reference_dit(cfg, None, torch.bfloat16, backend="gpu", device=dev)
        g = torch.Generator(device=dev).manual_seed(0)
        for name, p in model.named_parameters():
            if p.abs().sum().item() == 0:
                p.data.copy_(torch.randn(p.shape, device=dev, generator=g) * 0.02)
        b = images
        lat = torch.randn((b, 16, height // 8, width // 8), device=dev, generator=g).bfloat16()
        pos = torch.randn((b, ctx_len, cfg["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
        ctx = torch.cat([torch.zeros_like(pos), pos])
        mask = torch.ones((2 * b, ctx_len), device=dev)
        from flite_b200.pipeline import default_alpha, time_shift_schedule
        sched = time_shift_schedule(30, default_alpha(height // 8, width // 8))

        def step(i):
            nonlocal lat
            t, dt = sched[i % 30]
            tt = torch.tensor([t] * (2 * b), device=dev, dtype=torch.bfloat16)
            with torch.no_grad():
                out = model(torch.cat([lat] * 2), ctx, mask, tt)
            u, c = out.chunk(2)
            lat = lat + dt * (u + GUIDANCE * (c - u))

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        fl = flops_per_step(cfg, height, width, ctx_len, images)
        out = {"ms_per_step": ms, "steps_per_s": images * 1e3 / ms, "steps": steps, "warmup": warmup, "dtype": "bf16",
               "tflops": fl / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(lat.float()).all()),
               "module": "unmodified f_lite/model.py (oracle/_ref), eager PyTorch",
               "attention_backend": f"flash_attn {flash_attn.__version__} flash_attn_varlen_func (FA2; the reference "
                                    "imports FA3's flash_attn_interface, Hopper-only and not installed)",
               "norm_mlp_backend": f"liger_kernel {getattr(liger_kernel, '__version__', '0.8.0')} LigerRMSNorm / LigerSwiGLUMLP (Triton)",
               "gemm_backend": f"torch {torch.__version__} nn.Linear (cuBLASLt)"}
        del model
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # a missing / broken third-party kernel must not take the CPU line down with it
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


def run_reference(args, cfg, height, width, ctx_len, images, workload_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full, sample, cores, kind, desc = cpu_reference_sample(cfg, height, width, ctx_len, images,
                                                           repeats=args.warmup + args.steps)
    full_t, sample_t = full[args.warmup:], sample[args.warmup:]
    step_s = sum(full_t) / len(full_t)
    sample_s = sum(sample_t) / len(sample_t)
    val = images / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sample_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(workload_name, cfg, height, width, ctx_len, images, 1),
        "note": ("reference arm = the reference's CPU path on this box's host cores; GPUs unused for `value`. "
                 "`ms_per_step` is the time MEASURED per step (the bounded sample: 2 of the step's blocks); `value` "
                 "= 1 / ms_per_full_step_extrapolated.  Nothing-extrapolated numbers: `c1_full` (configs[0] in full) "
                 "and `reference_gpu` (the unmodified module on the B200)."),
        "ms_per_full_step_extrapolated": step_s * 1e3,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        line["c1_full"] = cpu_reference_c1_full()
    except Exception as e:
        line["c1_full"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    if not args.no_reference_gpu and workload_name in ("c1", "c2"):
        line["reference_gpu"] = reference_gpu_c2(cfg, height, width, ctx_len, images, steps=min(args.steps, 5))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def random_init_(model, seed):
    """Seeded 'de-zeroed default init' on the device: the reference zero-inits adaLN / final layers
    (f_lite/model.py:455-456,476-479), which makes every block a no-op; use N(0, 0.02) there instead."""
    import torch
    g = torch.Generator(device=model.device).manual_seed(seed)
    for name, p in model.named_parameters():
        if name == "register_tokens" or name.startswith(("adaLN_modulation.1.", "final_modulation.1.", "final_proj.")):
            p.data.copy_(torch.randn(p.shape, device=p.device, generator=g) * 0.02)
        elif "norm" in name:
            p.data.fill_(1.0)
        else:
            fan_in = p.shape[1:].numel() if p.dim() > 1 else None
            if fan_in is None:  # bias: bound from the matching weight's fan-in
                w = dict(model.named_parameters())[name[:-4] + "weight"]
                fan_in = w.shape[1:].numel()
            p.data.copy_((torch.rand(p.shape, device=p.device, generator=g) * 2 - 1) / math.sqrt(fan_in))


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def build_model(cfg, dev, seed=0):
    import torch

    import flite_b200
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = flite_b200.DiT(**cfg)
    torch.set_default_dtype(prev)
    random_init_(model, seed=seed)
    model.eval()
    model.hoist_context = False          # recompute the (t-independent) context path every step: nothing cached
    return model


def c1_on_gpu(dev):
    """BASELINE.json configs[0] in full on this GPU: tiny DiT, 256x256, 4 Euler steps, CFG 6, batch 1 (bf16),
    through flite_b200.denoise (eager and CUDA-graph replay); the like-for-like partner of the reference arm's c1_full."""
    import torch

    import flite_b200
    model = build_model(ARCH_TINY, dev)
    g = torch.Generator(device=dev).manual_seed(1234)
    lat = torch.randn((1, 16, 32, 32), device=dev, generator=g).bfloat16()
    pos = torch.randn((1, 256, 4096), device=dev, generator=g).bfloat16()
    neg = torch.zeros_like(pos)
    mask = torch.ones((2, 256), device=dev)
    out = {"config": "C1: tiny DiT d512 depth4 heads2, 256x256, 4 Euler steps, CFG 6, batch 1, bf16, flite_b200.denoise"}
    best = None
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = flite_b200.denoise(model, lat, neg, pos, mask, 4, GUIDANCE)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        if rep > 0:
            best = dt if best is None else min(best, dt)
    out["ms_per_step_eager"] = best * 1e3
    out["steps_per_s_eager"] = 1.0 / best
    out["finite"] = bool(torch.isfinite(res.float()).all())
    # CUDA-graph replay of the forward (what denoise(cuda_graph=True) does per step), capture outside the timed region
    from flite_b200 import ops
    from flite_b200.graphs import GraphedForward
    from flite_b200.pipeline import default_alpha, time_shift_schedule
    sched = time_shift_schedule(4, default_alpha(32, 32))
    t_all = torch.tensor([[t] * 2 for t, _ in sched], dtype=torch.bfloat16).to(dev)
    l2, acc = lat.clone(), lat.clone()
    gf = GraphedForward(model, l2, torch.cat([neg, pos]), mask, t_all[0], duplicate_latents=True)
    best = None
    for rep in range(4):
        l2.copy_(lat); acc.copy_(lat)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(4):
            v = gf(t_all[i])
            ops.cfg_euler(acc, v[:1], v[1:], GUIDANCE, sched[i][1], l2, do_cfg=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        if rep > 0:
            best = dt if best is None else min(best, dt)
    out["ms_per_step_cuda_graph"] = best * 1e3
    out["steps_per_s_cuda_graph"] = 1.0 / best
    out["cuda_graph_bit_identical_to_eager"] = bool(torch.equal(l2, res))
    out["timing"] = "wall clock around the 4-step loop incl. host launch overhead (graph capture excluded), best of 3"
    return out


def multi_gpu_layouts(model, dev, world, rank, steps):
    """The communicating layouts of SURVEY.md 8(e), run with the weights already on the GPUs.  Every rank first runs the
    1-GPU step itself (reference result + time), then each layout; results are compared on the velocity the step
    produced (rel-L2, 0.0 = bit-identical)."""
    import torch
    import torch.distributed as dist

    import flite_b200
    from flite_b200 import _lib, ops, parallel

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm()).item()

    def timed(fn, n):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    layouts = {2: [("c2", 2, 1), ("c4", 1, 2), ("c4", 2, 1)], 4: [("c4", 1, 4), ("c4", 2, 2)], 8: [("c4", 2, 4)]}.get(world, [])
    results = []
    base = {}      # workload -> (1-GPU velocity, 1-GPU ms/step)
    n = max(2, min(steps, 4))
    for wl, cfg_ranks, sp_ranks in layouts:
        arch, height, width, ctx_len, _ = WORKLOADS[wl]
        g = torch.Generator(device=dev).manual_seed(4321)          # same inputs on every rank
        lat0 = torch.randn((1, 16, height // 8, width // 8), device=dev, generator=g).bfloat16()
        pos = torch.randn((1, ctx_len, arch["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
        neg = torch.zeros_like(pos)
        ctx2, mask2 = torch.cat([neg, pos]), torch.ones((2, ctx_len), device=dev)
        t2 = torch.full((2,), 0.9, device=dev).bfloat16()
        if wl not in base:
            model.enable_sequence_parallel(None)
            lat, acc = lat0.clone(), lat0.clone()
            v1 = flite_b200.denoise_step(model, lat, acc, ctx2, mask2, t2, 0.01, GUIDANCE, True).clone()
            ms1 = timed(lambda: flite_b200.denoise_step(model, lat, acc, ctx2, mask2, t2, 0.01, GUIDANCE, True), n)
            base[wl] = (v1, ms1)
        v1, ms1 = base[wl]
        sp_group, cfg_group, _, n_rep = parallel.make_groups(cfg_ranks, sp_ranks)
        model.enable_sequence_parallel(sp_group, fused=sp_group is not None)
        if cfg_group is not None:
            half = dist.get_rank(cfg_group)
            ctx_in, mask_in, t_in = (neg, pos)[half], mask2[:1], t2[:1]
        else:
            ctx_in, mask_in, t_in = ctx2, mask2, t2
        lat, acc = lat0.clone(), lat0.clone()
        v = flite_b200.denoise_step(model, lat, acc, ctx_in, mask_in, t_in, 0.01, GUIDANCE, True, cfg_group=cfg_group).clone()
        _lib.watchdog_ok()
        r = torch.tensor([rel(v, v1)], device=dev, dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        step = lambda: flite_b200.denoise_step(model, lat, acc, ctx_in, mask_in, t_in, 0.01, GUIDANCE, True, cfg_group=cfg_group)
        ms = timed(step, n)
        hs = None
        if sp_group is not None:      # time spent in the peer-memory flag handshakes of one step (rank-local, max over ranks)
            ops.TRACE = []
            step()
            rep = ops.trace_report()
            ops.TRACE = None
            h = torch.tensor([rep.get("p2p signal+wait", (0, 0.0))[1]], device=dev, dtype=torch.float64)
            dist.all_reduce(h, op=dist.ReduceOp.MAX)
            hs = float(h.item())
        # same layout with the hybrid self-attention schedule where its heuristic picks it (model.attn_streamk = "auto": long
        # sequences whose whole-unit rounds waste >= 10 % of the last one; whole rounds in lock step + stream-K shares over
        # the rest).  Not batch-invariant, hence a separate number.
        sk = None
        if sp_group is not None and model._use_streamk_auto(1 if cfg_group is not None else 2,
                                                            arch["num_heads"] // sp_ranks, 16 + (height // 16) * (width // 16)):
            sk_default = model.attn_streamk
            model.attn_streamk = "auto"
            lat, acc = lat0.clone(), lat0.clone()
            v_sk = flite_b200.denoise_step(model, lat, acc, ctx_in, mask_in, t_in, 0.01, GUIDANCE, True, cfg_group=cfg_group).clone()
            r_sk = torch.tensor([rel(v_sk, v1)], device=dev, dtype=torch.float64)
            dist.all_reduce(r_sk, op=dist.ReduceOp.MAX)
            ms_sk = timed(step, n)
            model.attn_streamk = sk_default
            sk = {"schedule": "hybrid: whole rounds in lock step + stream-K shares over the last 1..2 units per cluster",
                  "ms_per_step": ms_sk, "speedup_vs_1gpu": ms1 / ms_sk, "rel_l2_vs_1gpu": float(r_sk.item())}
        _lib.watchdog_ok()
        model.enable_sequence_parallel(None)
        name = "x".join(p for p in ((f"cfg{cfg_ranks}" if cfg_ranks > 1 else ""), (f"sp{sp_ranks}" if sp_ranks > 1 else "")) if p)
        results.append({"workload": workload_desc(wl, arch, height, width, ctx_len, 1, 1).split(",")[0] + f" ({height}x{width}, one image)",
                        "layout": name, "gpus_per_image": cfg_ranks * sp_ranks, "replicas": n_rep,
                        "exchange": ("Ulysses: QKV-GEMM / attention epilogues store into the owner rank over NVLink peer memory "
                                     "(fused, 2 flag handshakes per block)" if sp_ranks > 1 else "") +
                                    (" + " if sp_ranks > 1 and cfg_ranks > 1 else "") +
                                    ("CFG halves on different GPUs: one NCCL all-gather of the velocity per step" if cfg_ranks > 1 else ""),
                        "ms_per_step": ms, "ms_per_step_1gpu": ms1, "speedup_vs_1gpu": ms1 / ms,
                        "rel_l2_vs_1gpu": float(r.item()), "handshake_ms_per_step": hs, "steps_timed": n,
                        "with_streamk_attention": sk})
    return results


def run_ours(args, cfg, height, width, ctx_len, images, workload_name):
    import torch
    import torch.distributed as dist

    import flite_b200
    from flite_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().flite_check_device(), "flite_check_device")

    # ---- model + synthetic inputs (weights replicated per GPU; each rank denoises its own images)
    model = build_model(cfg, dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    b = images
    lat0 = torch.randn((b, 16, height // 8, width // 8), device=dev, generator=g).bfloat16()
    pos = torch.randn((b, ctx_len, cfg["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
    ctx = torch.cat([torch.zeros_like(pos), pos])                  # [negative(zeros), positive]  pipeline.py:160,266
    mask = torch.ones((2 * b, ctx_len), device=dev)
    alpha = flite_b200.pipeline.default_alpha(height // 8, width // 8)
    n_sched = max(30, args.steps + args.warmup)
    sched = flite_b200.pipeline.time_shift_schedule(n_sched, alpha)
    t_all = torch.tensor([[t] * (2 * b) for t, _ in sched], dtype=torch.bfloat16).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-input run (value): roofline events around the dominant kernel
    lat = lat0.clone()
    acc = lat0.clone()
    step_i = [0]

    def one_step():
        i = step_i[0] % n_sched
        flite_b200.denoise_step(model, lat, acc, ctx, mask, t_all[i], sched[i][1], GUIDANCE, True)
        step_i[0] += 1

    for _ in range(args.warmup):
        one_step()
    barrier()
    inter = int(cfg["hidden_size"] * cfg["mlp_ratio"])
    dom_key = (2 * b * (16 + (height // 16) * (width // 16)), 2 * inter, cfg["hidden_size"], ops.EPI_SWIGLU)
    dom_events = []

    def hook(name, phase, key):
        if key[:4] != dom_key:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if phase == "begin":
            dom_events.append([e, None])
        else:
            dom_events[-1][1] = e

    clocks = ClockSampler(local)
    clocks.start()
    ops.PROFILE_HOOK = hook
    launches0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ops.PROFILE_HOOK = None
    launches = ops.LAUNCHES[0] - launches0
    clk = clocks.stop()
    _lib.watchdog_ok()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * images / (ms_step / 1e3)     # image-steps per second over all ranks
    assert torch.isfinite(lat.float()).all(), "latents diverged"

    fl = flops_per_step(cfg, height, width, ctx_len, images)
    peaks = measured_peaks()
    dom_ms = [a.elapsed_time(bb) for a, bb in dom_events if bb is not None]
    M, N, K, _ = dom_key
    dom_flops = 2.0 * M * N * K
    roof = None
    if dom_ms:
        avg = sum(dom_ms) / len(dom_ms)
        ach = dom_flops / (avg * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel<2,256,6,EPI_SWIGLU> (MLP gate|up, tcgen05 cta_group::2)",
                "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "flops_per_launch": dom_flops, "avg_launch_ms": avg, "launches_timed": len(dom_ms),
                "share_of_step": avg * len(dom_ms) / args.steps / ms_step, "traffic": None}
        if (M, N, K) == (8224, 24576, 3072):
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, read from the newest ncu --set full
            # export under profiles/ (algorithmic: A + W + C = 2(MK + NK + MN/2) bytes)
            # template arguments <cta_group 2, BLOCK_N 256, stages, epilogue 2 = EPI_SWIGLU>, with or without "(int)" casts
            tb, src = ncu_traffic(r"gemm_bf16_kernel<\D*2,\D*256,\D*\d+,\D*2>")
            roof["traffic"] = tb
            roof["traffic_unit"] = f"bytes of DRAM traffic per launch (ncu --set full, {src})" if src else None
            roof["algorithmic_bytes_per_launch"] = 2.0 * (M * K + N * K + M * N // 2)

    # ---- end-to-end run: host (pinned) buffers in, host buffer out, every step
    lat_h = lat0.cpu().pin_memory()
    ctx_h = ctx.cpu().pin_memory()
    mask_h = mask.cpu().pin_memory()
    t_h = t_all.cpu().pin_memory()
    out_h = torch.empty_like(lat_h).pin_memory()
    lat_d, acc_d = torch.empty_like(lat0), torch.empty_like(lat0)
    ctx_d, mask_d, t_d = torch.empty_like(ctx), torch.empty_like(mask), torch.empty_like(t_all[0])
    h2d = lat_h.numel() * 2 + ctx_h.numel() * 2 + mask_h.numel() * 4 + t_h[0].numel() * 2
    d2h = out_h.numel() * 2

    def e2e_step(i):
        i = i % n_sched
        lat_d.copy_(lat_h, non_blocking=True)
        ctx_d.copy_(ctx_h, non_blocking=True)
        mask_d.copy_(mask_h, non_blocking=True)
        t_d.copy_(t_h[i], non_blocking=True)
        acc_d.copy_(lat_d)
        flite_b200.denoise_step(model, lat_d, acc_d, ctx_d, mask_d, t_d, sched[i][1], GUIDANCE, True)
        out_h.copy_(lat_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lat_h.copy_(out_h)                      # the host owns the state between steps

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    lat_h.copy_(lat0.cpu())
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_val = world * images / (e2e_ms / 1e3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": bench_config(workload_name, cfg, height, width, ctx_len, images, world),
        "tflops_per_gpu": fl / (ms_step * 1e-3) / 1e12,
        "tensor_frac_of_burst_peak": fl / (ms_step * 1e-3) / 1e12 / peaks["bf16"],
        "tensor_frac_of_sustained_peak": fl / (ms_step * 1e-3) / 1e12 / peaks["bf16_sustained"],
        "tensor_frac_of_nominal_2250": fl / (ms_step * 1e-3) / 1e12 / 2250.0,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": "flite_b200.denoise_step (DiT.forward + flite_cfg_euler) on host buffers"},
        "gpu_launches": launches, "clocks": clk, "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline and args.workload in ("c1", "c2"):
        full, sample, cores, kind, desc = cpu_reference_sample(cfg, height, width, ctx_len, images, repeats=1)
        line["cpu_baseline"] = {"value": images / full[0], "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    if world == 1 and not args.no_c1 and args.workload == "c2":
        try:
            line["c1"] = c1_on_gpu(dev)
        except Exception as e:
            line["c1"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    if world > 1 and not args.no_multi_gpu and args.workload == "c2":
        try:
            line["multi_gpu"] = multi_gpu_layouts(model, dev, world, rank, args.steps)
        except Exception as e:
            line["multi_gpu"] = {"failed": f"{type(e).__name__}: {str(e)[:300]}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FLITE_BENCH_WORKLOAD", "c2"), choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c1", action="store_true")
    ap.add_argument("--no-multi-gpu", action="store_true", help="N>1: skip the communicating layouts (multi_gpu key)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="reference arm: skip the GPU run of the unmodified module")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg, height, width, ctx_len, images = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, height, width, ctx_len, images, args.workload)
    else:
        run_ours(args, cfg, height, width, ctx_len, images, args.workload)


if __name__ == "__main__":
    main()
