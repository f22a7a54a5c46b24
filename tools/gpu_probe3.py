"""Attention bottleneck experiments (2-CTA kernel): which stage paces the tile loop?"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"; d, H = 3072, 12
def bench(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
OUT = {}
for (B, L, tag) in [(2, 4112, "c2"), (2, 16400, "c4")]:
    qkv = torch.randn(B * L, 3 * d, device=dev).bfloat16()
    cu = torch.arange(B + 1, device=dev, dtype=torch.int32) * L
    o = torch.empty(B * L, d, device=dev, dtype=torch.bfloat16)
    fl = 4 * B * H * L * L * 256
    for var in (5, 7, 8):
        for dbg in (0,):
            lib.flite_set_tuning(3, dbg)
            ms = bench(lambda: ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, 256 ** -0.5, out=o, variant=var))
            OUT[f"{tag}_v{var}_dbg{dbg}"] = fl / ms / 1e9
            print(tag, "variant", var, "debug", dbg, "ms %.3f" % ms, "TF/s %.0f" % (fl / ms / 1e9), flush=True)
    lib.flite_set_tuning(3, 0)
    _lib.watchdog_ok()
# rmsnorm A/B
T = 8224
x = torch.randn(T, d, device=dev).bfloat16(); w = torch.ones(d, device=dev).bfloat16()
mod = torch.randn(2, 9 * d, device=dev).bfloat16(); y = torch.empty_like(x)
for mode in (1, 2):
    lib.flite_set_tuning(0, mode)
    ms = bench(lambda: ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=T // 2, out=y), n=50)
    OUT[f"rmsnorm_mode{mode}_gbs"] = 2 * T * d * 2 / ms / 1e6
    print("rmsnorm mode", mode, "GB/s", OUT[f"rmsnorm_mode{mode}_gbs"])
lib.flite_set_tuning(0, 0)
json.dump(OUT, open("gpurun_out/probe3.json", "w"), indent=1)
