"""GEMM tail-split A/B at the hot shapes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
def bench(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
OUT = {}
T, d = 8224, 3072
for (M, N, K, epi) in [(75520, 3072, 3072, 1), (75520, 9216, 3072, 0), (75520, 24576, 3072, 2), (75520, 3072, 12288, 1), (32896, 24576, 3072, 2), (32896, 3072, 12288, 1)]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    x = torch.randn(M, N if epi != 2 else N // 2, device=dev).bfloat16(); gate = torch.randn(2, N, device=dev).bfloat16(); torch.cuda.empty_cache()
    fl = 2 * M * N * K
    res = {}
    for split in (2, 4, 8, 12, 16, 24, 32):
        lib.flite_set_tuning(5, split)
        if epi == 1: f = lambda: ops.gemm(a, w, None, epilogue=1, resid=x, gate=gate, rows_per_sample=M // 2, out=x)
        elif epi == 2: f = lambda: ops.gemm(a, w, None, epilogue=2, out=x)
        else: f = lambda: ops.gemm(a, w, None, out=x)
        res[f"G{split}"] = fl / bench(f) / 1e9
    ms = bench(lambda: torch.matmul(a, w.t())); res["cublas"] = fl / ms / 1e9
    print((M, N, K, epi), {k: round(v) for k, v in res.items()}, flush=True)
    OUT[f"{M}x{N}x{K}_epi{epi}"] = res
lib.flite_set_tuning(5, 0)
_lib.watchdog_ok()
json.dump(OUT, open("gpurun_out/probe4.json", "w"), indent=1)
