"""Runs the three dominant kernels at C2 shapes a few times each (target of `ncu --set full`)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
REP = int(os.environ.get("FLITE_PROFILE_REPEATS", "6"))
T, d = 8224, 3072
torch.manual_seed(0)
if which in ("gemm", "all"):
    a = (torch.randn(T, d, device=dev) * 0.5).bfloat16()
    wgu = (torch.randn(8 * d, d, device=dev) * 0.02).bfloat16()
    out = torch.empty(T, 4 * d, device=dev, dtype=torch.bfloat16)
    for _ in range(REP):
        ops.gemm(a, wgu, None, epilogue=ops.EPI_SWIGLU, out=out)
    wdn = (torch.randn(d, 4 * d, device=dev) * 0.02).bfloat16()
    x = torch.randn(T, d, device=dev).bfloat16(); gate = torch.randn(2, d, device=dev).bfloat16()
    for _ in range(REP):
        ops.gemm(out, wdn, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x)
    wq = (torch.randn(3 * d, d, device=dev) * 0.02).bfloat16(); bq = torch.randn(3 * d, device=dev).bfloat16()
    cos = torch.rand(T // 2, 128, device=dev).bfloat16(); sin = torch.rand(T // 2, 128, device=dev).bfloat16()
    qkv = torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16)
    for _ in range(REP):
        ops.gemm(a, wq, bq, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=T // 2, out=qkv)
    wo = (torch.randn(d, d, device=dev) * 0.02).bfloat16()
    for _ in range(REP):   # attention / cross-attention output projection (gated residual, K = 3072)
        ops.gemm(a, wo, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x)
if which in ("attn", "all"):
    qkv = torch.randn(T, 3 * d, device=dev).bfloat16()
    cu = torch.arange(3, device=dev, dtype=torch.int32) * (T // 2)
    o = torch.empty(T, d, device=dev, dtype=torch.bfloat16)
    var = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    for _ in range(REP):     # one cluster per 256-query unit
        ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, 12, T // 2, 256 ** -0.5, out=o, variant=var)
    _lib.load().flite_set_tuning(15, 1)
    for _ in range(REP):     # the DiT's default self-attention path: persistent kernel, whole units round-robin
        ops.attention_streamk(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, 12, T // 2, T // 2, 256 ** -0.5, out=o)
    _lib.load().flite_set_tuning(15, 0)
    ctx_kv = torch.randn(512, 2 * d, device=dev).bfloat16()      # cross-attention: 2 x 256 text keys
    cuk = torch.arange(3, device=dev, dtype=torch.int32) * 256
    for _ in range(REP):
        ops.attention_varlen(qkv[:, :d], ctx_kv[:, :d], ctx_kv[:, d:], cu, cuk, 12, T // 2, 256 ** -0.5, out=o)
if which in ("norm", "all"):
    x = torch.randn(T, d, device=dev).bfloat16(); w = torch.ones(d, device=dev).bfloat16()
    mod = torch.randn(2, 9 * d, device=dev).bfloat16(); y = torch.empty_like(x)
    for _ in range(REP):
        ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=T // 2, out=y)
torch.cuda.synchronize()
_lib.watchdog_ok()
print("ok")
