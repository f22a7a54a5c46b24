"""Isolated timing of the self-attention kernels at the C2 / C4 shapes: one cluster per query tile (flite_attention_varlen)
vs the persistent stream-K wave (flite_attention_streamk), interleaved, CUDA events, L2 flushed between launches.
Writes gpurun_out/attn_bench.json.   python tools/attn_bench.py [--reps 20]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = "cuda"
_lib.check(_lib.load().flite_check_device(), "flite_check_device")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for name, (B, H, L) in {"c2_self": (2, 12, 4112), "c4_self": (2, 12, 16400), "c4_sp4_rank": (2, 3, 16400)}.items():
    g = torch.Generator(device=dev).manual_seed(0)
    d = H * 256
    qkv = torch.randn(B * L, 3 * d, device=dev, generator=g).bfloat16()
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    cu = (torch.arange(B + 1, dtype=torch.int32) * L).to(dev)
    o = torch.empty(B * L, d, dtype=torch.bfloat16, device=dev)
    scale = 256 ** -0.5
    fns = {"per_tile": lambda: ops.attention_varlen(q, k, v, cu, cu, H, L, scale, out=o),
           "streamk": lambda: ops.attention_streamk(q, k, v, cu, cu, H, L, L, scale, out=o)}
    for f in fns.values():
        f(); f()
    t = {n: [] for n in fns}
    for i in range(args.reps):
        for n in (list(fns) if i % 2 == 0 else list(fns)[::-1]):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fns[n](); e1.record(); torch.cuda.synchronize()
            t[n].append(e0.elapsed_time(e1))
    fl = 4.0 * B * H * L * L * 256
    out[name] = {n: {"ms": sorted(v_)[len(v_) // 2], "tflops": fl / sorted(v_)[len(v_) // 2] / 1e9} for n, v_ in t.items()}
    # back-to-back (sustained, power-capped regime): 40 launches in a row
    for n, f in fns.items():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40 if L < 10000 else 8):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (40 if L < 10000 else 8)
        out[name][n]["ms_back_to_back"] = ms
        out[name][n]["tflops_back_to_back"] = fl / ms / 1e9
    print(name, out[name], flush=True)
    del qkv, o
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/attn_bench.json", "w"), indent=1)
