"""Eager vs CUDA-graph replay of the denoise loop: ms/step at C1 (tiny, launch-bound) and C2 (10B, GPU-bound)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, flite_b200
from flite_b200 import _lib
dev = torch.device("cuda", 0)
OUT = {}
for wl, steps in (("c1", 50), ("c2", 5)):
    cfg, H, W, Lc, images = bench.WORKLOADS[wl]
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev): model = flite_b200.DiT(**cfg)
    torch.set_default_dtype(prev)
    bench.random_init_(model, 0); model.eval()
    g = torch.Generator(device=dev).manual_seed(1)
    lat = torch.randn((images, 16, H // 8, W // 8), device=dev, generator=g).bfloat16()
    pos = torch.randn((images, Lc, cfg["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
    neg = torch.zeros_like(pos)
    for graph in (False, True):
        run = lambda n: flite_b200.denoise(model, lat, neg, pos, None, n, 6.0, cuda_graph=graph)
        run(3); torch.cuda.synchronize()
        # time n and 2n steps: the difference removes the one-off capture / warm-up cost
        ts = []
        for n in (steps, 2 * steps):
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                torch.cuda.synchronize(); e0.record(); run(n); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ts.append(best)
        OUT[f"{wl}_{'graph' if graph else 'eager'}_total_ms_{steps}_{2 * steps}_steps"] = ts
        OUT[f"{wl}_{'graph' if graph else 'eager'}_ms_per_step"] = (ts[1] - ts[0]) / steps
        print(wl, "graph" if graph else "eager", OUT[f"{wl}_{'graph' if graph else 'eager'}_ms_per_step"], flush=True)
    del model; torch.cuda.empty_cache()
_lib.watchdog_ok()
json.dump(OUT, open("gpurun_out/probe7_graphs.json", "w"), indent=1)
