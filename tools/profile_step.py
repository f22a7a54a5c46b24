"""Per-op CUDA-event breakdown of one denoise step of a bench workload (in situ: real clocks, no profiler).

  python tools/profile_step.py [--workload c2] [--steps 3]

Prints {label: [launches, total_ms, share]} per step (ops.TRACE) and the untraced remainder (launch gaps + the
few untraced tiny ops).  Event pairs cost ~2 us each, so the traced step is slightly slower than bench.py's."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import flite_b200
from flite_b200 import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--out", default=None)
args = ap.parse_args()
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
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(args.steps): step()
e1.record(); torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / args.steps
ops.TRACE = []
e0.record()
for _ in range(args.steps): step()
e1.record()
rep = ops.trace_report(); ops.TRACE = None
traced_ms = e0.elapsed_time(e1) / args.steps
_lib.watchdog_ok()
tot = sum(ms for _, ms in rep.values()) / args.steps
out = {"workload": args.workload, "ms_per_step_untraced": plain_ms, "ms_per_step_traced": traced_ms,
       "sum_of_ops_ms": tot, "remainder_ms": traced_ms - tot,
       "ops": {k: [n // args.steps, round(ms / args.steps, 3), round(ms / args.steps / traced_ms, 4)]
               for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}}
print(json.dumps(out, indent=1))
if args.out:
    json.dump(out, open(args.out, "w"), indent=1)
