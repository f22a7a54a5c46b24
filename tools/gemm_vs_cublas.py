"""Same-process A/B of the four hot C2 GEMM shapes: flite_gemm_bf16 (plain-store epilogue) vs torch.matmul (cuBLASLt),
interleaved launch by launch so both see the same clocks / power state; cold-ish L2 (operands of the other shapes are
touched in between).  Writes gpurun_out/gemm_vs_cublas.json.

  python tools/gemm_vs_cublas.py [--reps 30]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--m", type=int, default=8224)
args = ap.parse_args()
dev = "cuda"
_lib.check(_lib.load().flite_check_device(), "flite_check_device")
M = args.m
SHAPES = {"qkv": (M, 9216, 3072), "proj": (M, 3072, 3072), "gate_up": (M, 24576, 3072), "down": (M, 3072, 12288)}
g = torch.Generator(device=dev).manual_seed(0)
out = {"M": M, "reps": args.reps, "method": "CUDA events per launch, own kernel and torch.matmul alternating, median"}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (m, n, k) in SHAPES.items():
    a = (torch.randn(m, k, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev, generator=g) * 0.02).bfloat16()
    c1 = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    c2 = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    wt = w.t()
    for _ in range(3):
        ops.gemm(a, w, None, out=c1)
        torch.matmul(a, wt, out=c2)
    rel = ((c1.float() - c2.float()).norm() / c2.float().norm()).item()
    t_own, t_cub = [], []
    for i in range(args.reps):
        for which in ((0, 1) if i % 2 == 0 else (1, 0)):
            flush.zero_()                                   # evict both operands from L2: every launch starts cold
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if which == 0:
                ops.gemm(a, w, None, out=c1)
            else:
                torch.matmul(a, wt, out=c2)
            e1.record()
            torch.cuda.synchronize()
            (t_own if which == 0 else t_cub).append(e0.elapsed_time(e1))
    t_own.sort(); t_cub.sort()
    fl = 2.0 * m * n * k
    mo, mc = t_own[len(t_own) // 2], t_cub[len(t_cub) // 2]
    out[name] = {"shape": [m, n, k], "own_ms": mo, "cublas_ms": mc, "own_tflops": fl / mo / 1e9, "cublas_tflops": fl / mc / 1e9,
                 "own_over_cublas": mc / mo, "rel_l2_own_vs_cublas": rel}
    print(name, out[name], flush=True)
    del a, w, c1, c2
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_vs_cublas.json", "w"), indent=1)
