"""rmsnorm+modulate kernel variants at the C2 / C5 shapes: isolated GB/s (algorithmic bytes = read x + write n)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"; d = 3072
def bench(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
OUT = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for T in (8224, 32896):
    x = torch.randn(T, d, device=dev).bfloat16(); w = torch.ones(d, device=dev).bfloat16()
    mod = torch.randn(2, 9 * d, device=dev).bfloat16(); y = torch.empty_like(x)
    for mode in (1, 2, 3):
        lib.flite_set_tuning(0, mode)
        ms = bench(lambda: ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=T // 2, out=y))
        # cold variant: flush L2 between launches (input not L2-resident)
        def cold():
            flush.zero_(); ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=T // 2, out=y)
        ms_both = bench(cold, n=20); ms_flush = bench(lambda: flush.zero_(), n=20)
        OUT[f"T{T}_mode{mode}"] = {"warm_us": ms * 1e3, "warm_gbs": 2 * T * d * 2 / ms / 1e6,
                                   "cold_us": (ms_both - ms_flush) * 1e3, "cold_gbs": 2 * T * d * 2 / (ms_both - ms_flush) / 1e6}
        print(T, mode, OUT[f"T{T}_mode{mode}"], flush=True)
lib.flite_set_tuning(0, 0)
_lib.watchdog_ok()
json.dump(OUT, open("gpurun_out/probe5.json", "w"), indent=1)
