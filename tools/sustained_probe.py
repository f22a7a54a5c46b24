"""Sustained (power-capped) rate, SM clock and board power of each hot kernel run alone in a ~1.5 s loop, next to
torch.matmul (cuBLAS) on the same shapes: tells which kernels are limited by the 1000 W cap (clock drops, power at the
cap) and which by their own pipeline (power below the cap at a high clock).  Writes gpurun_out/sustained_probe.json.

  python tools/sustained_probe.py [--seconds 1.5]
"""
import argparse, json, os, subprocess, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=1.5)
args = ap.parse_args()
dev = "cuda"
_lib.check(_lib.load().flite_check_device(), "flite_check_device")
T, d = 8224, 3072
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
a = rn(T, d, sc=0.5); hm = rn(T, 4 * d, sc=0.5)
wgu = rn(8 * d, d, sc=0.02); wdn = rn(d, 4 * d, sc=0.02); wq = rn(3 * d, d, sc=0.02); wo = rn(d, d, sc=0.02)
bq = rn(3 * d); gate = rn(2, d); x = rn(T, d)
cos = torch.rand(T // 2, 128, device=dev).bfloat16(); sin = torch.rand(T // 2, 128, device=dev).bfloat16()
o_gu = torch.empty(T, 4 * d, dtype=torch.bfloat16, device=dev); qkv = torch.empty(T, 3 * d, dtype=torch.bfloat16, device=dev)
o_at = torch.empty(T, d, dtype=torch.bfloat16, device=dev); nrm = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
o_full = torch.empty(T, 8 * d, dtype=torch.bfloat16, device=dev)
cu = (torch.arange(3, dtype=torch.int32) * (T // 2)).to(dev)
nw = torch.ones(d, device=dev).bfloat16(); mod = rn(2, 2 * d, sc=0.1)
qkv_in = rn(T, 3 * d)
ssq = torch.empty(T, d // 128, dtype=torch.float32, device=dev)
KERNELS = {
    "gemm_swiglu 8224x24576x3072": (lambda: ops.gemm(a, wgu, None, epilogue=ops.EPI_SWIGLU, out=o_gu), 2.0 * T * 8 * d * d),
    "cublas 8224x24576x3072": (lambda: torch.matmul(a, wgu.t(), out=o_full), 2.0 * T * 8 * d * d),
    "gemm_down 8224x3072x12288 (gated res)": (lambda: ops.gemm(hm, wdn, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x), 2.0 * T * d * 4 * d),
    "cublas 8224x3072x12288": (lambda: torch.matmul(hm, wdn.t(), out=o_at), 2.0 * T * d * 4 * d),
    "gemm_qkv_rope 8224x9216x3072": (lambda: ops.gemm(a, wq, bq, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=T // 2, out=qkv), 2.0 * T * 3 * d * d),
    "cublas 8224x9216x3072": (lambda: torch.matmul(a, wq.t(), out=qkv), 2.0 * T * 3 * d * d),
    "gemm_proj 8224x3072x3072 (gated res)": (lambda: ops.gemm(a, wo, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x), 2.0 * T * d * d),
    "cublas 8224x3072x3072": (lambda: torch.matmul(a, wo.t(), out=o_at), 2.0 * T * d * d),
    "attention per-tile 2x12x4112^2": (lambda: ops.attention_varlen(qkv_in[:, :d], qkv_in[:, d:2 * d], qkv_in[:, 2 * d:], cu, cu, 12, T // 2, 256 ** -0.5, out=o_at), 4.0 * 2 * 12 * (T // 2) ** 2 * 256),
    "attention stream-K 2x12x4112^2": (lambda: ops.attention_streamk(qkv_in[:, :d], qkv_in[:, d:2 * d], qkv_in[:, 2 * d:], cu, cu, 12, T // 2, T // 2, 256 ** -0.5, out=o_at), 4.0 * 2 * 12 * (T // 2) ** 2 * 256),
    "rmsnorm_modulate two-pass 8224x3072": (lambda: ops.rmsnorm_modulate(x, nw, 1, mod[:, :d], mod[:, d:], rows_per_sample=T // 2, out=nrm), 0.0),
}
from tools.sustained_probe_lib import sample_loop

res = {}
for name, (fn, flops) in KERNELS.items():
    ms, clk, pw = sample_loop(fn, args.seconds)
    r = {"ms": ms, "sm_mhz": clk, "power_w": pw}
    if flops:
        r["tflops"] = flops / ms / 1e9
        if clk:
            r["tensor_util_at_clock"] = r["tflops"] / (148 * 8192 * clk * 1e6 / 1e12)
        if pw:
            r["pj_per_flop"] = pw * ms * 1e-3 / flops * 1e12
    res[name] = r
    print(name, r, flush=True)
    time.sleep(0.5)
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sustained_probe.json", "w"), indent=1)
