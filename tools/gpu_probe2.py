"""Perf A/B probe (round 1, second pass): attention variants, QKV epilogue, rmsnorm, GEMM shapes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
dev = "cuda"; OUT = {}
def bench(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
T, d, H = 8224, 3072, 12
torch.manual_seed(0)
# attention
for (B, L, tag) in [(2, 4112, "c2"), (2, 16400, "c4")]:
    qkv = torch.randn(B * L, 3 * d, device=dev).bfloat16()
    cu = torch.arange(B + 1, device=dev, dtype=torch.int32) * L
    o1 = torch.empty(B * L, d, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o1)
    fl = 4 * B * H * L * L * 256
    for var, o in ((1, o1), (3, o2), (4, o2)):
        ms = bench(lambda: ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, 256 ** -0.5, out=o, variant=var), n=5 if L > 5000 else 20)
        OUT[f"attn_{tag}_wg{var}_tflops"] = fl / ms / 1e9
        print(tag, "variant", var, "ms", ms, "TF/s", fl / ms / 1e9, flush=True)
    OUT[f"attn_{tag}_cg2_vs_v1_rel"] = rel(o2, o1); print("  cg2 vs v1 rel", rel(o2, o1))
    _lib.watchdog_ok()
# qkv epilogue
a = (torch.randn(T, d, device=dev) * 0.5).bfloat16()
wq = (torch.randn(3 * d, d, device=dev) * 0.02).bfloat16(); bq = torch.randn(3 * d, device=dev).bfloat16()
cos = torch.rand(T // 2, 128, device=dev).bfloat16(); sin = torch.rand(T // 2, 128, device=dev).bfloat16()
qkv = torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16)
fl = 2 * T * 3 * d * d
ms = bench(lambda: ops.gemm(a, wq, bq, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=T // 2, out=qkv))
OUT["qkv_rope_tflops"] = fl / ms / 1e9; print("qkv rope epilogue TF/s", fl / ms / 1e9)
ms = bench(lambda: ops.gemm(a, wq, bq, out=qkv)); OUT["qkv_plain_tflops"] = fl / ms / 1e9; print("qkv plain TF/s", fl / ms / 1e9)
ms = bench(lambda: torch.nn.functional.linear(a, wq, bq)); print("cublas TF/s", fl / ms / 1e9); OUT["qkv_cublas_tflops"] = fl / ms / 1e9
# rmsnorm
x = torch.randn(T, d, device=dev).bfloat16(); w = torch.ones(d, device=dev).bfloat16()
mod = torch.randn(2, 9 * d, device=dev).bfloat16(); y = torch.empty_like(x)
ms = bench(lambda: ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=T // 2, out=y), n=50)
OUT["rmsnorm_gbs"] = 2 * T * d * 2 / ms / 1e6; print("rmsnorm GB/s", OUT["rmsnorm_gbs"], "us", ms * 1e3)
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(OUT, open("gpurun_out/probe2.json", "w"), indent=1)
print(json.dumps(OUT, indent=1))
