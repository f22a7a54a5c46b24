"""Timestep-path skinny GEMMs at the 10B shapes: GEMV (weight streaming) vs the tcgen05 tile variant; GB/s of weights."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    ms = 0.0
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ms += e0.elapsed_time(e1)
    return ms / n
OUT = {}
for name, (M, N, K) in {"time_embed.0": (2, 12288, 3072), "time_embed.2": (2, 3072, 12288), "adaLN_modulation": (2, 27648, 3072),
                        "final_modulation": (2, 6144, 3072), "adaLN_modulation B=16": (16, 27648, 3072)}.items():
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); b = torch.randn(N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = {}
    for tag, var in (("gemv", 5), ("tcgen05_1cta_n128", 3)):
        if var == 5 and M > 8: continue
        ms = timed(lambda: ops.gemm(a, w, b, act=1, variant=var, out=out))
        res[tag] = {"us": ms * 1e3, "weight_GBs": N * K * 2 / ms / 1e6}
    OUT[f"{name} {M}x{N}x{K}"] = res
    print(name, (M, N, K), {k: (round(v["us"], 1), round(v["weight_GBs"])) for k, v in res.items()}, flush=True)
_lib.watchdog_ok()
json.dump(OUT, open("gpurun_out/probe9_gemv.json", "w"), indent=1)
