"""How much of the gated-residual GEMM launches is their epilogue's global traffic?  proj / down shapes with the real
epilogue vs the same launch with the residual loads and C stores switched off (FLITE_TUNE_GEMM_DEBUG bit0), interleaved,
isolated (L2 flushed) and back to back.  Writes gpurun_out/gemm_epilogue_probe.json."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops
lib = _lib.load()
_lib.check(lib.flite_check_device(), "flite_check_device")
dev = "cuda"
T, d = 8224, 3072
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gate = rn(2, d); out = {}
for name, K in (("proj", d), ("down", 4 * d)):
    a, w, x = rn(T, K, sc=0.5), rn(d, K, sc=0.02), rn(T, d)
    fl = 2.0 * T * d * K
    def run(dbg):
        lib.flite_set_tuning(17, dbg)
        ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=T // 2, out=x)
        lib.flite_set_tuning(17, 0)
    cases = {"real": 0, "no_global_traffic": 1}
    ts = {c: [] for c in cases}
    for c in cases: run(cases[c])
    for i in range(16):
        for c in (list(cases) if i % 2 == 0 else list(cases)[::-1]):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(cases[c]); e1.record(); torch.cuda.synchronize()
            ts[c].append(e0.elapsed_time(e1))
    res = {}
    for c in cases:
        t = sorted(ts[c])[len(ts[c]) // 2]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(200): run(cases[c])
        e1.record(); torch.cuda.synchronize()
        b = e0.elapsed_time(e1) / 200
        res[c] = {"ms": t, "tflops": fl / t / 1e9, "ms_b2b": b, "tflops_b2b": fl / b / 1e9}
    out[name] = res
    print(name, res, flush=True)
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_epilogue_probe.json", "w"), indent=1)
