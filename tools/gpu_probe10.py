"""Cross-attention at the C2 shape (2 x 4112 queries, 256 keys per sequence, 12 heads): general 2-CTA kernel vs the
persistent resident-K/V kernel, isolated, L2-warm (the q projection has just been written in the real step)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
dev = "cuda"; H, d = 12, 3072
def bench(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
OUT = {}
for (B, Lq, Lk) in [(2, 4112, 256), (2, 2056, 256), (2, 4112, 128), (16, 4720, 256)]:
    q = torch.randn(B * Lq, d, device=dev).bfloat16()
    kv = torch.randn(B * Lk, 2 * d, device=dev).bfloat16()
    cu_q = torch.arange(B + 1, device=dev, dtype=torch.int32) * Lq
    cu_k = torch.arange(B + 1, device=dev, dtype=torch.int32) * Lk
    o = torch.empty(B * Lq, d, device=dev, dtype=torch.bfloat16)
    fl = 4 * B * H * Lq * Lk * 256
    for var in (5, 9):
        ms = bench(lambda: ops.attention_varlen(q, kv[:, :d], kv[:, d:], cu_q, cu_k, H, Lq, 256 ** -0.5, out=o, variant=var))
        OUT[f"B{B}_Lq{Lq}_Lk{Lk}_v{var}"] = {"us": ms * 1e3, "tflops": fl / ms / 1e9}
        print((B, Lq, Lk), "variant", var, round(ms * 1e3, 1), "us", round(fl / ms / 1e9), "TF/s", flush=True)
_lib.watchdog_ok()
json.dump(OUT, open("gpurun_out/probe10_cross_attn.json", "w"), indent=1)
