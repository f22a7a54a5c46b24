"""First GPU probe: validates every kernel of libflite_b200.so against torch on a B200 and records
library baselines (cuBLAS bf16 on the hot GEMM shapes, FA2 / SDPA at head_dim 256).  Diagnostics only;
the parity tests proper live in tests/."""
import json, os, sys, time, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flite_b200
from flite_b200 import ops, _lib

dev = "cuda"
OUT = {}
torch.manual_seed(0)

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()

def bench(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

SECTIONS = {}
def section(name):
    def deco(f):
        SECTIONS[name] = f
        return f
    return deco

def run_section(name):
    print(f"\n=== {name} ===", flush=True)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), "cpus", os.cpu_count())
    _lib.check(_lib.load().flite_check_device(), "check_device")
    try:
        SECTIONS[name]()
        _lib.watchdog_ok()
    except Exception as e:
        print(f"!!! {name} FAILED: {type(e).__name__}: {e}")
        traceback.print_exc()
        OUT[name] = "FAILED: " + str(e)[:300]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(OUT, open(f"gpurun_out/probe1_{name}.json", "w"), indent=1)

@section("cfg_euler")
def _():
    n = 16 * 128 * 128
    u = torch.randn(n, device=dev).bfloat16(); c = torch.randn(n, device=dev).bfloat16()
    acc = torch.randn(n, device=dev).bfloat16(); acc0 = acc.clone()
    lat = torch.empty_like(acc)
    ops.cfg_euler(acc, u, c, 6.0, 0.0123, lat)
    v = u + 6.0 * (c - u)
    ref = acc0 + 0.0123 * v
    print("bf16 acc max abs diff", (acc.float() - ref.float()).abs().max().item(), "lat==acc", torch.equal(lat, acc))
    accf = acc0.float().clone(); lat2 = torch.empty_like(acc)
    ops.cfg_euler(accf, u, c, 6.0, 0.0123, lat2)
    reff = acc0.float() + 0.0123 * v.float()
    print("fp32 acc max abs diff", (accf - reff).abs().max().item())
    OUT["cfg_euler_bitexact"] = bool(torch.equal(acc, ref))

@section("rmsnorm_modulate")
def _():
    T, d, B = 2 * 4112, 3072, 2
    x = torch.randn(T, d, device=dev).bfloat16(); w = (1 + 0.1 * torch.randn(d, device=dev)).bfloat16()
    mod = (0.5 * torch.randn(B, 9 * d, device=dev)).bfloat16()
    sc, sh = mod[:, d:2 * d], mod[:, 0:d]
    y = ops.rmsnorm_modulate(x, w, 1, sc, sh, rows_per_sample=T // B)
    xf = x.float(); rstd = torch.rsqrt((xf * xf).sum(-1, keepdim=True) / d + 1e-6)
    n = (xf * rstd).bfloat16() * w
    ref = n * (1 + sc.repeat_interleave(T // B, 0)) + sh.repeat_interleave(T // B, 0)
    print("rel", rel(y, ref), "exact frac", (y == ref).float().mean().item())
    ms = bench(lambda: ops.rmsnorm_modulate(x, w, 1, sc, sh, rows_per_sample=T // B, out=y))
    print("ms", ms, "GB/s", 2 * T * d * 2 / ms / 1e6)
    OUT["rmsnorm_rel"] = rel(y, ref); OUT["rmsnorm_gbs"] = 2 * T * d * 2 / ms / 1e6

def gemm_case(M, N, K, variant, tag):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    ref = torch.nn.functional.linear(a, w, bias)
    out = ops.gemm(a, w, bias, variant=variant)
    _lib.watchdog_ok()
    r = rel(out, ref)
    bad = (out.float() - ref.float()).abs() > 0.05 * ref.float().abs().max()
    print(f"[{tag}] M{M} N{N} K{K} variant {variant}: rel {r:.3e} bad {bad.float().mean().item():.4f}", flush=True)
    if r > 1e-2:
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("   bad rows", rows[:16].tolist(), "... count", rows.numel(), "bad cols", cols[:16].tolist(), "count", cols.numel())
        print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
    return r

@section("gemm_1cta_small")
def _():
    OUT["gemm_1cta_n128_small"] = gemm_case(128, 128, 64, 3, "1cta n128 single k-block")
    OUT["gemm_1cta_n128_k256"] = gemm_case(128, 128, 256, 3, "1cta n128")
    OUT["gemm_1cta_n64"] = gemm_case(200, 64, 512, 4, "1cta n64")
    OUT["gemm_1cta_n256"] = gemm_case(256, 512, 512, 1, "1cta n256")
    OUT["gemm_1cta_n256_big"] = gemm_case(8224, 3072, 3072, 1, "1cta n256")

@section("gemm_2cta")
def _():
    OUT["gemm_2cta_small"] = gemm_case(256, 256, 64, 2, "2cta single tile single kb")
    OUT["gemm_2cta_med"] = gemm_case(512, 512, 512, 2, "2cta")
    OUT["gemm_2cta_big"] = gemm_case(8224, 3072, 3072, 2, "2cta")

@section("gemm_epilogues")
def _():
    M, N, K, B = 8224, 3072, 3072, 2
    for variant in (1, 2):
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        resid = torch.randn(M, N, device=dev).bfloat16(); gate = torch.randn(B, N, device=dev).bfloat16()
        out = ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=resid, gate=gate, rows_per_sample=M // B, variant=variant)
        ref = resid + torch.nn.functional.linear(a, w) * gate.repeat_interleave(M // B, 0)
        print("gated variant", variant, "rel", rel(out, ref)); OUT[f"gemm_gated_v{variant}"] = rel(out, ref)
        inter = 4 * 3072
        wg = (torch.randn(inter, K, device=dev) * 0.05).bfloat16(); wu = (torch.randn(inter, K, device=dev) * 0.05).bfloat16()
        wi = ops.interleave_gate_up(wg, wu)
        out = ops.gemm(a, wi, None, epilogue=ops.EPI_SWIGLU, variant=variant)
        g = torch.nn.functional.linear(a, wg); u = torch.nn.functional.linear(a, wu)
        ref = torch.nn.functional.silu(g.float()).bfloat16() * u
        print("swiglu variant", variant, "rel", rel(out, ref)); OUT[f"gemm_swiglu_v{variant}"] = rel(out, ref)
        # qkv + rope + qknorm
        L = M // B
        cos = torch.rand(L, 128, device=dev) * 2 - 1; sin = torch.sqrt(1 - cos * cos)
        wq = (torch.randn(3 * N, K, device=dev) * 0.05).bfloat16(); bq = torch.randn(3 * N, device=dev).bfloat16()
        out = ops.gemm(a, wq, bq, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * N, rows_per_sample=L, variant=variant)
        qkv = torch.nn.functional.linear(a, wq, bq)
        plain = ops.gemm(a, wq, bq, variant=variant)
        print("  plain qkv rel", rel(plain, qkv))
        ops.rope_qknorm_(plain, 2 * N // 256, cos, sin, rows_per_sample=L)
        x = qkv[:, :2 * N].reshape(M, -1, 256).float()
        c = cos.repeat(B, 1)[:, None, :]; s = sin.repeat(B, 1)[:, None, :]
        x1, x2 = x[..., :128], x[..., 128:]
        y = torch.cat([x1 * c + x2 * s, x1 * (-s) + x2 * c], -1).bfloat16().float()
        y = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().reshape(M, 2 * N)
        ref = torch.cat([y, qkv[:, 2 * N:]], 1)
        print("qkv_rope fused variant", variant, "rel", rel(out, ref), " standalone rope kernel rel", rel(plain, ref))
        OUT[f"gemm_qkvrope_v{variant}"] = rel(out, ref); OUT[f"rope_kernel_v{variant}"] = rel(plain, ref)

@section("gemm_perf")
def _():
    shapes = [(8224, 9216, 3072), (8224, 3072, 3072), (8224, 24576, 3072), (8224, 3072, 12288)]
    for (M, N, K) in shapes:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2 * M * N * K
        ms = bench(lambda: torch.matmul(a, w.t(), out=out)); cub = fl / ms / 1e9
        res = {"cublas_tflops": cub}
        for variant in (1, 2):
            try:
                ms = bench(lambda: ops.gemm(a, w, None, variant=variant, out=out)); res[f"v{variant}_tflops"] = fl / ms / 1e9
            except Exception as e:
                res[f"v{variant}"] = str(e)[:80]
        print((M, N, K), res, flush=True); OUT[f"perf_{M}x{N}x{K}"] = res

@section("attention")
def _():
    from oracle.dit_oracle import flash_attn_varlen
    for (B, H, Lq, Lk, tag) in [(1, 1, 128, 128, "one tile"), (1, 2, 256, 384, "multi tile"), (2, 2, 272, 272, "tiny self"),
                                (2, 2, 272, 17, "tiny cross"), (2, 12, 4112, 4112, "C2 self")]:
        d = H * 256
        qkv = torch.randn(B * Lq, 3 * d, device=dev).bfloat16() if Lq == Lk else None
        if qkv is not None:
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        else:
            q = torch.randn(B * Lq, d, device=dev).bfloat16(); kv = torch.randn(B * Lk, 2 * d, device=dev).bfloat16()
            k, v = kv[:, :d], kv[:, d:]
        # normalise q,k like the model does so logits are O(sqrt(256)) scaled
        cu_q = torch.arange(0, B + 1, device=dev, dtype=torch.int32) * Lq
        cu_k = torch.arange(0, B + 1, device=dev, dtype=torch.int32) * Lk
        out = ops.attention_varlen(q, k, v, cu_q, cu_k, H, Lq, 256 ** -0.5)
        _lib.watchdog_ok()
        ref = flash_attn_varlen(q.reshape(B * Lq, H, 256), k.reshape(B * Lk, H, 256), v.reshape(B * Lk, H, 256), cu_q, cu_k, 256 ** -0.5)
        r = rel(out, ref.reshape(B * Lq, d))
        print(f"attention [{tag}] B{B} H{H} Lq{Lq} Lk{Lk}: rel {r:.3e}", flush=True)
        OUT[f"attn_{tag}"] = r
        if Lq == 4112:
            ms = bench(lambda: ops.attention_varlen(q, k, v, cu_q, cu_k, H, Lq, 256 ** -0.5, out=out))
            fl = 4 * B * H * Lq * Lk * 256
            print("   ours ms", ms, "TFLOP/s", fl / ms / 1e9); OUT["attn_c2_tflops"] = fl / ms / 1e9
            try:
                from flash_attn import flash_attn_varlen_func
                q3, k3, v3 = (t.reshape(B * Lq, H, 256).contiguous() for t in (q, k, v))
                o2 = flash_attn_varlen_func(q3, k3, v3, cu_q, cu_k, Lq, Lk, softmax_scale=256 ** -0.5)
                print("   FA2 rel vs ref", rel(o2.reshape(B * Lq, d), ref.reshape(B * Lq, d)))
                ms = bench(lambda: flash_attn_varlen_func(q3, k3, v3, cu_q, cu_k, Lq, Lk, softmax_scale=256 ** -0.5))
                print("   FA2 ms", ms, "TFLOP/s", fl / ms / 1e9); OUT["fa2_c2_tflops"] = fl / ms / 1e9
            except Exception as e:
                print("   FA2 failed:", repr(e)[:300]); OUT["fa2"] = repr(e)[:200]
            try:
                q4, k4, v4 = (t.reshape(B, Lq, H, 256).transpose(1, 2).contiguous() for t in (q, k, v))
                f = lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)
                ms = bench(f); print("   SDPA ms", ms, "TFLOP/s", fl / ms / 1e9); OUT["sdpa_c2_tflops"] = fl / ms / 1e9
            except Exception as e:
                print("   SDPA failed:", repr(e)[:300])

@section("small_ops")
def _():
    B, C, H, W, P, d = 2, 16, 64, 96, 2, 512
    x = torch.randn(B, C, H, W, device=dev).bfloat16(); w = (torch.randn(d, C, P, P, device=dev) * 0.1).bfloat16()
    b = torch.randn(d, device=dev).bfloat16(); reg = torch.randn(1, 16, d, device=dev).bfloat16()
    tok = ops.patch_embed(x, w, b, reg, P)
    ref = torch.nn.functional.conv2d(x, w, b, stride=P).flatten(2).transpose(1, 2)
    ref = torch.cat([reg.repeat(B, 1, 1), ref], 1).reshape(-1, d)
    print("patch_embed rel", rel(tok, ref)); OUT["patch_embed"] = rel(tok, ref)
    t = torch.tensor([0.9333, 0.25], device=dev).bfloat16()
    import math
    freqs = torch.exp(-math.log(10000) * torch.arange(0, d // 2, dtype=torch.float32) / (d // 2)).to(dev)
    e = ops.timestep_embed(t.float(), True, freqs, d)
    args = (t * 1000)[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1).bfloat16()
    print("timestep_embed max diff", (e.float() - ref.float()).abs().max().item()); OUT["timestep"] = (e.float() - ref.float()).abs().max().item()
    L = 16 + (H // P) * (W // P)
    tk = torch.randn(B * L, 64, device=dev).bfloat16()
    o = ops.unpatchify(tk, B, C, H, W, P, 16)
    r = tk.view(B, L, 64)[:, 16:].view(B, H // P, W // P, P, P, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)
    print("unpatchify exact", torch.equal(o, r)); OUT["unpatchify"] = bool(torch.equal(o, r))
    ctx = torch.randn(3, 40, 512, device=dev).bfloat16(); mask = (torch.rand(3, 40, device=dev) > 0.3).float()
    packed, cu = ops.pack_context(ctx, mask)
    idx = mask.reshape(-1).nonzero().flatten()
    refp = ctx.reshape(-1, 512)[idx]
    print("pack exact", torch.equal(packed[: idx.numel()], refp), cu.tolist(), mask.sum(1).tolist()); OUT["pack"] = bool(torch.equal(packed[: idx.numel()], refp))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_section(sys.argv[1])
    else:
        import subprocess
        allout = {}
        for name in SECTIONS:
            try:
                r = subprocess.run([sys.executable, __file__, name], timeout=300)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
            allout[name + "_rc"] = rc
            try:
                allout.update(json.load(open(f"gpurun_out/probe1_{name}.json")))
            except Exception:
                pass
        print("\nSUMMARY", json.dumps(allout, indent=1))
        json.dump(allout, open("gpurun_out/probe1.json", "w"), indent=1)
