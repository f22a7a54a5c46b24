"""A/B of tuning knobs on one box, same process: alternates settings of FLITE_TUNE_* and times denoise steps.

  python tools/ab_step.py --workload c2 --knob 8 --values 1,0 --rounds 3 --steps 5
  python tools/ab_step.py --workload c2 --settings "base;k16=1;sk=rr;sk=hybrid" --rounds 4     (k<key>=<value>, sk=<DiT.attn_streamk>)
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import flite_b200
from flite_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--knob", type=int, default=-1)
ap.add_argument("--settings", default="")
ap.add_argument("--values", default="1,0")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--out", default=None)
args = ap.parse_args()
vals = [int(v) for v in args.values.split(",")]
cfg, H, W, Lc, images = bench.WORKLOADS[args.workload]
dev = torch.device("cuda", 0)
prev = torch.get_default_dtype(); torch.set_default_dtype(torch.bfloat16)
with torch.device(dev): model = flite_b200.DiT(**cfg)
torch.set_default_dtype(prev)
bench.random_init_(model, 0); model.eval(); model.hoist_context = False
g = torch.Generator(device=dev).manual_seed(1234)
lat = torch.randn((images, 16, H // 8, W // 8), device=dev, generator=g).bfloat16(); acc = lat.clone()
pos = torch.randn((images, Lc, cfg["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
ctx = torch.cat([torch.zeros_like(pos), pos]); mask = torch.ones((2 * images, Lc), device=dev)
t = torch.full((2 * images,), 0.9, device=dev).bfloat16()
step = lambda: flite_b200.denoise_step(model, lat, acc, ctx, mask, t, 0.01, 6.0, True)
lib = _lib.load()
for _ in range(4): step()
if args.settings:
    names = args.settings.split(";")
    sk_default = model.attn_streamk
    res = {n: [] for n in names}
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for r in range(args.rounds):
        for n in (names if r % 2 == 0 else names[::-1]):
            touched = []
            model.attn_streamk = sk_default
            for a in ([] if n == "base" else n.split(",")):
                key, val = a.split("=")
                if key == "sk": model.attn_streamk = val
                else:
                    lib.flite_set_tuning(int(key[1:]), int(val)); touched.append(int(key[1:]))
            step(); torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps): step()
            e1.record(); torch.cuda.synchronize()
            res[n].append(e0.elapsed_time(e1) / args.steps)
            for key in touched: lib.flite_set_tuning(key, 0)
    model.attn_streamk = sk_default
    _lib.watchdog_ok()
    out = {"workload": args.workload, "ms_per_step": res, "median": {n: sorted(v)[len(v) // 2] for n, v in res.items()}}
    print(json.dumps(out))
    if args.out: json.dump(out, open(args.out, "w"))
    sys.exit(0)
res = {v: [] for v in vals}
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
for r in range(args.rounds):
    for v in vals:
        lib.flite_set_tuning(args.knob, v)
        step(); torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps): step()
        e1.record(); torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / args.steps)
lib.flite_set_tuning(args.knob, 0)
_lib.watchdog_ok()
out = {"workload": args.workload, "knob": args.knob, "ms_per_step": {str(v): res[v] for v in vals},
       "median": {str(v): sorted(res[v])[len(res[v]) // 2] for v in vals}}
print(json.dumps(out))
if args.out: json.dump(out, open(args.out, "w"))
