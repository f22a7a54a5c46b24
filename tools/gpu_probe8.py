"""GEMM rasterisation band x L2 eviction hints at the C2 hot shapes: time per launch with a cold L2 (flushed between
launches, CUDA events around the GEMM only).  Run the same script under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --kernel-name regex:gemm_bf16_kernel
with FLITE_PROBE_ONCE=1 to get the DRAM bytes of every configuration in the same order (one launch each)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import ops, _lib
lib = _lib.load(); dev = "cuda"
ONCE = os.environ.get("FLITE_PROBE_ONCE") == "1"
T, d = 8224, 3072
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=8):
    ms = 0.0
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ms += e0.elapsed_time(e1)
    return ms / n
a = (torch.randn(T, d, device=dev) * 0.5).bfloat16()
h = (torch.randn(T, 4 * d, device=dev) * 0.5).bfloat16()
wgu = (torch.randn(8 * d, d, device=dev) * 0.02).bfloat16(); hmid = torch.empty(T, 4 * d, device=dev, dtype=torch.bfloat16)
wdn = (torch.randn(d, 4 * d, device=dev) * 0.02).bfloat16(); x = torch.randn(T, d, device=dev).bfloat16(); gate = torch.randn(2, d, device=dev).bfloat16()
wq = (torch.randn(3 * d, d, device=dev) * 0.02).bfloat16(); bq = torch.randn(3 * d, device=dev).bfloat16()
cos = torch.rand(T // 2, 128, device=dev).bfloat16(); sin = torch.rand(T // 2, 128, device=dev).bfloat16(); qkv = torch.empty(T, 3 * d, device=dev, dtype=torch.bfloat16)
wo = (torch.randn(d, d, device=dev) * 0.02).bfloat16()
CASES = {
    "swiglu_8224x24576x3072": (lambda: ops.gemm(a, wgu, None, epilogue=ops.EPI_SWIGLU, out=hmid), 2 * T * 8 * d * d),
    "down_8224x3072x12288": (lambda: ops.gemm(h, wdn, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x), 2 * T * d * 4 * d),
    "qkv_8224x9216x3072": (lambda: ops.gemm(a, wq, bq, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=T // 2, out=qkv), 2 * T * 3 * d * d),
    "proj_8224x3072x3072": (lambda: ops.gemm(a, wo, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x), 2 * T * d * d),
}
BANDS = (0, 1, 4, 8)
HINTS = ((1, 1), (3, 2), (2, 3), (3, 3))
OUT = {"order": []}
for name, (fn, fl) in CASES.items():
    for band in BANDS:
        for ha, hb in HINTS:
            lib.flite_set_tuning(5, band); lib.flite_set_tuning(9, ha); lib.flite_set_tuning(10, hb)
            key = f"{name}|band{band}|A{ha}B{hb}"
            OUT["order"].append(key)
            if ONCE:
                fn(); torch.cuda.synchronize()
            else:
                ms = timed(fn); OUT[key] = {"us": ms * 1e3, "tflops": fl / ms / 1e9}
                print(key, round(ms * 1e3, 1), "us", round(fl / ms / 1e9), "TF/s", flush=True)
for k in (5, 9, 10): lib.flite_set_tuning(k, 0)
_lib.watchdog_ok()
if not ONCE:
    json.dump(OUT, open("gpurun_out/probe8_hints.json", "w"), indent=1)
else:
    json.dump(OUT["order"], open("gpurun_out/probe8_order.json", "w"))
