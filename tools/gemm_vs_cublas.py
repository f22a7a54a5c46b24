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
ap.add_argument("--variants", default="0", help="own GEMM variants to time (0 = auto; 6 = seven-stage ring), comma-separated")
ap.add_argument("--sustained", type=int, default=0, help="also time N back-to-back launches per contender (power-capped regime)")
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
    variants = [int(v) for v in args.variants.split(",")]
    fns = {f"own_v{v}": (lambda v=v: ops.gemm(a, w, None, out=c1, variant=v)) for v in variants}
    fns["cublas"] = lambda: torch.matmul(a, wt, out=c2)
    names = list(fns)
    for f in fns.values():
        f()
    ts = {nm: [] for nm in names}
    for i in range(args.reps):
        for nm in (names if i % 2 == 0 else names[::-1]):
            flush.zero_()                                   # evict both operands from L2: every launch starts cold
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fns[nm]()
            e1.record()
            torch.cuda.synchronize()
            ts[nm].append(e0.elapsed_time(e1))
    fl = 2.0 * m * n * k
    med = {nm: sorted(t)[len(t) // 2] for nm, t in ts.items()}
    mo, mc = med[f"own_v{variants[0]}"], med["cublas"]
    out[name] = {"shape": [m, n, k], "own_ms": mo, "cublas_ms": mc, "own_tflops": fl / mo / 1e9, "cublas_tflops": fl / mc / 1e9,
                 "own_over_cublas": mc / mo, "rel_l2_own_vs_cublas": rel,
                 "tflops": {nm: fl / t_ / 1e9 for nm, t_ in med.items()}}
    if args.sustained:
        sus = {}
        for rnd in range(2):
            for nm in names:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.sustained):
                    fns[nm]()
                e1.record(); torch.cuda.synchronize()
                sus.setdefault(nm, []).append(fl / (e0.elapsed_time(e1) / args.sustained) / 1e9)
        out[name]["tflops_sustained"] = {nm: max(v) for nm, v in sus.items()}
    print(name, out[name], flush=True)
    del a, w, c1, c2
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_vs_cublas.json", "w"), indent=1)
