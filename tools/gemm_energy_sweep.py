"""Sustained (power-capped) rate of the K = 12288 down-projection and the K = N = 3072 projection GEMMs under the kernel's
tuning switches, next to cuBLAS: which choice costs energy per FLOP?  Writes gpurun_out/gemm_energy_sweep.json."""
import json, os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops
from tools.sustained_probe_lib import sample_loop

lib = _lib.load()
_lib.check(lib.flite_check_device(), "flite_check_device")
dev = "cuda"
T, d = 8224, 3072
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
shapes = {"down": (rn(T, 4 * d, sc=0.5), rn(d, 4 * d, sc=0.02)), "proj": (rn(T, d, sc=0.5), rn(d, d, sc=0.02)),
          "qkv": (rn(T, d, sc=0.5), rn(3 * d, d, sc=0.02))}
gate = rn(2, d); x = rn(T, d)
res = {}
secs = float(os.environ.get("SWEEP_SECONDS", "1.0"))


def run(tag, fn, flops):
    ms, clk, pw = sample_loop(fn, secs)
    r = {"ms": ms, "sm_mhz": clk, "power_w": pw, "tflops": flops / ms / 1e9, "pj_per_flop": (pw or 0) * ms * 1e-3 / flops * 1e12}
    res[tag] = r
    print(tag, r, flush=True)


for name, (a, w) in shapes.items():
    M, K = a.shape
    N = w.shape[0]
    fl = 2.0 * M * N * K
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    run(f"{name} cublas", lambda: torch.matmul(a, w.t(), out=out), fl)
    run(f"{name} store default", lambda: ops.gemm(a, w, None, out=out), fl)
    if N == d:
        run(f"{name} gated_res default", lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate,
                                                           rows_per_sample=T // 2, out=x), fl)
    for band in (1, 4, 8, 17):
        lib.flite_set_tuning(5, band)
        run(f"{name} store band{band}", lambda: ops.gemm(a, w, None, out=out), fl)
    lib.flite_set_tuning(5, 0)
    lib.flite_set_tuning(4, 1)
    run(f"{name} store no-tail-split", lambda: ops.gemm(a, w, None, out=out), fl)
    lib.flite_set_tuning(4, 0)
    lib.flite_set_tuning(14, 1)
    run(f"{name} store padded-M-tail", lambda: ops.gemm(a, w, None, out=out), fl)
    lib.flite_set_tuning(14, 0)
    for var, vn in ((1, "1cta_n256"), (3, "1cta_n128")):
        run(f"{name} store {vn}", lambda: ops.gemm(a, w, None, out=out, variant=var), fl)
    # exact multiple of the tile: no ragged M at all
    a2 = a[:8192]
    o2 = out[:8192]
    run(f"{name} store M=8192", lambda: ops.gemm(a2, w, None, out=o2), 2.0 * 8192 * N * K)
    run(f"{name} cublas M=8192", lambda: torch.matmul(a2, w.t(), out=o2), 2.0 * 8192 * N * K)
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_energy_sweep.json", "w"), indent=1)
