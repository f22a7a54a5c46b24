"""Pacing experiments on the self-attention kernels at the C2 / C4 shapes.  Cases (interleaved rep by rep so that all see
the same clocks; L2 flushed before every isolated launch; median) + 40 back-to-back launches per case:
  v5 / v6_2wg                 one cluster per 256-query unit: default, two softmax warpgroups
  v5_*                        debug bits of FLITE_TUNE_ATTN_DEBUG (1 no softmax math, 2 no K/V reloads), staged stores,
                              per-thread row stores instead of the TMA-store epilogue
  sk0 / sk1 / sk2             persistent kernel: stream-K shares, whole units round-robin, hybrid
Writes gpurun_out/attn_probe.json.     python tools/attn_probe.py [--reps 12] [--shapes ..] [--cases ..]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--shapes", default="c2_self,c4_self")
ap.add_argument("--cases", default="")
args = ap.parse_args()
dev = "cuda"
lib = _lib.load()
_lib.check(lib.flite_check_device(), "flite_check_device")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = {"c2_cross": (2, 12, 4112, 256), "c3_cross": (16, 12, 4720, 256), "c2_self": (2, 12, 4112), "c4_self": (2, 12, 16400), "c4_sp4_rank": (2, 3, 16400), "c4_cfg2sp4_rank": (1, 3, 16400), "c4_sp2_rank": (2, 6, 16400), "c5_self": (8, 12, 4112),
          "c3_self": (16, 12, 4720)}
# tag: (kind, variant, {tuning key: value})   kind "u" = one cluster per unit, "p" = persistent
CASES = {
    "v5": ("u", 5, {}), "v6_2wg": ("u", 6, {}), 
    "v5_nosoftmax": ("u", 5, {3: 1}), "v5_noloads": ("u", 5, {3: 2}), "v5_neither": ("u", 5, {3: 3}),
    "v5_staged_stores": ("u", 5, {7: 1}), 
    "v10_persistent": ("u", 10, {}), "v9_xres": ("u", 9, {}),
    "v5_rowstores": ("u", 5, {16: 1}), "v6_rowstores": ("u", 6, {16: 1}),
    "sk0_streamk": ("p", 0, {15: 0}), "sk1_roundrobin": ("p", 0, {15: 1}), "sk2_hybrid": ("p", 0, {15: 2}),
    "sk1_rr_noepilogue": ("p", 0, {15: 1, 3: 8}),
}
names = args.cases.split(",") if args.cases else list(CASES)
out = {}
for shape in args.shapes.split(","):
    B, H, L = SHAPES[shape][:3]
    Lk = SHAPES[shape][3] if len(SHAPES[shape]) > 3 else L       # cross-attention shapes: Lk text keys per sequence
    g = torch.Generator(device=dev).manual_seed(0)
    d = H * 256
    qkv = torch.randn(B * L, 3 * d, device=dev, generator=g).bfloat16()
    q, k, v = qkv[:, :d], qkv[:B * Lk, d:2 * d], qkv[:B * Lk, 2 * d:]
    cu = (torch.arange(B + 1, dtype=torch.int32) * L).to(dev)
    cuk = (torch.arange(B + 1, dtype=torch.int32) * Lk).to(dev)
    scale = 256 ** -0.5
    fl = 4.0 * B * H * L * Lk * 256
    ref = torch.empty(B * L, d, dtype=torch.bfloat16, device=dev)
    ops.attention_varlen(q, k, v, cu, cuk, H, L, scale, out=ref, variant=5)
    o = torch.empty_like(ref)

    def run(tag):
        kind, variant, tune = CASES[tag]
        for key, val in tune.items():
            lib.flite_set_tuning(key, val)
        if kind == "u":
            ops.attention_varlen(q, k, v, cu, cuk, H, L, scale, out=o, variant=variant)
        else:
            ops.attention_streamk(q, k, v, cu, cuk, H, L, Lk, scale, out=o)
        for key in tune:
            lib.flite_set_tuning(key, 0)

    res = {n: {} for n in names}
    for n in names:
        o.zero_()
        run(n); run(n)
        torch.cuda.synchronize()
        if not (CASES[n][2].get(3, 0) & 11):
            res[n]["bit_equal_to_v5"] = bool(torch.equal(o, ref))
            res[n]["rel_l2_vs_v5"] = ((o.float() - ref.float()).norm() / ref.float().norm()).item()
    ts = {n: [] for n in names}
    for i in range(args.reps):
        for n in (names if i % 2 == 0 else names[::-1]):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(n); e1.record(); torch.cuda.synchronize()
            ts[n].append(e0.elapsed_time(e1))
    nb = 40 if B * H * L * Lk < 2 * 12 * 10000 * 10000 else 6
    for rnd in range(2):
        for n in names:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(nb):
                run(n)
            e1.record(); torch.cuda.synchronize()
            res[n].setdefault("b2b", []).append(e0.elapsed_time(e1) / nb)
    for n in names:
        t = sorted(ts[n])
        med = t[len(t) // 2]
        b2b = min(res[n].pop("b2b"))
        res[n].update({"ms": med, "ms_min": t[0], "tflops": fl / med / 1e9, "ms_b2b": b2b, "tflops_b2b": fl / b2b / 1e9})
        print(shape, n, {k_: (round(v_, 4) if isinstance(v_, float) else v_) for k_, v_ in res[n].items()}, flush=True)
    out[shape] = res
    del qkv, ref, o
_lib.watchdog_ok()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/attn_probe.json", "w"), indent=1)
