"""Isolated and back-to-back timing of the gated-residual GEMM at the C2 proj / down shapes in its three forms:
plain, + ssq slots, + fused norm of finished rows (and the separate single-pass norm launch for reference)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops
dev = "cuda"
_lib.check(_lib.load().flite_check_device(), "flite_check_device")
T, d = 8224, 3072
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
gate = rn(2, d); x = rn(T, d); nw = torch.ones(d, device=dev).bfloat16(); mod = rn(2, 2 * d, sc=0.1)
ssq = torch.empty(T, d // 64, dtype=torch.float32, device=dev)
nb = torch.empty(T, d, dtype=torch.bfloat16, device=dev)
cnt = torch.zeros(2 * ((T + 127) // 128), dtype=torch.int32, device=dev)
res = {}
for name, K in (("proj", d), ("down", 4 * d)):
    a = rn(T, K, sc=0.5); w = rn(d, K, sc=0.02)
    nd = dict(out=nb, weight=nw, weight_mode=1, scale=mod[:, :d], shift=mod[:, d:], counters=cnt)
    forms = {
        "plain": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x),
        "ssq": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq),
        "ssq+norm_launch": lambda: (ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq),
                                    ops.rmsnorm_modulate(x, nw, 1, mod[:, :d], mod[:, d:], rows_per_sample=T // 2, out=nb, ssq=ssq)),
        "fused_norm": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq, norm=nd),
        "fused_norm[no job]": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq, norm=nd),
        "fused_norm[no fence]": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq, norm=nd),
        "fused_norm[no job, no fence]": lambda: ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq, norm=nd),
    }
    dbg = {"fused_norm[no job]": 1, "fused_norm[no fence]": 2, "fused_norm[no job, no fence]": 3}
    for fn_name, fn in forms.items():
        _lib.load().flite_set_tuning(15, dbg.get(fn_name, 0))
        cnt.zero_()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[f"{name} {fn_name}"] = e0.elapsed_time(e1) / 50 * 1000
        print(f"{name:5s} {fn_name:18s} {res[f'{name} {fn_name}']:8.1f} us", flush=True)
import ctypes
lib = _lib.load()
lib.flite_debug_nf_read.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_ulonglong * 4)()
for name, K in (("proj", d), ("down", 4 * d)):
    a = rn(T, K, sc=0.5); w = rn(d, K, sc=0.02)
    nd = dict(out=nb, weight=nw, weight_mode=1, scale=mod[:, :d], shift=mod[:, d:], counters=cnt)
    lib.flite_set_tuning(15, 4)
    cnt.zero_()
    lib.flite_debug_nf_read(buf)
    for _ in range(10):
        ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x, ssq_out=ssq, norm=nd)
    lib.flite_debug_nf_read(buf)
    print(name, "job max us", buf[0] / 1e3, "mean us", buf[1] / max(buf[2], 1) / 1e3, "jobs/launch", buf[2] / 10, "units/launch (CTAs)", buf[3] / 10, flush=True)
lib.flite_set_tuning(15, 0)
_lib.watchdog_ok()
json.dump(res, open("gpurun_out/fused_norm_probe.json", "w"), indent=1)
