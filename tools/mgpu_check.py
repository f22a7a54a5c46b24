"""Multi-GPU parity + timing (run under torchrun): Ulysses sequence parallelism and CFG-split vs the 1-GPU path.

  torchrun --nproc-per-node N tools/mgpu_check.py [--cfg-ranks C] [--sp-ranks S] [--big]
"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flite_b200
from flite_b200 import parallel, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--cfg-ranks", type=int, default=1)
ap.add_argument("--sp-ranks", type=int, default=1)
ap.add_argument("--big", action="store_true", help="time the 10B architecture at 2048^2 (config C4)")
ap.add_argument("--res", type=int, default=2048)
ap.add_argument("--trace", action="store_true", help="per-op CUDA-event breakdown of one step (rank 0 prints it)")
ap.add_argument("--fused", action="store_true", help="Ulysses exchanges fused into the kernels over NVLink peer memory")
args = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", 0)); rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sp_group, cfg_group, rep, nrep = parallel.make_groups(args.cfg_ranks, args.sp_ranks)

def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
def build(cfg, seed=0):
    prev = torch.get_default_dtype(); torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev): m = flite_b200.DiT(**cfg)
    torch.set_default_dtype(prev)
    g = torch.Generator(device=dev).manual_seed(seed)
    for n, p in m.named_parameters():
        if "norm" in n: p.data.fill_(1.0)
        else: p.data.copy_(torch.randn(p.shape, device=dev, generator=g) * (0.02 if p.dim() < 2 or n.startswith(("adaLN", "final", "register")) else p.shape[-1] ** -0.5))
    return m.eval()

out = {"world": world, "cfg_ranks": args.cfg_ranks, "sp_ranks": args.sp_ranks, "fused": args.fused}
# ---------------- parity on a small model (4 heads so that P = 2 or 4 divides)
cfg = dict(in_channels=16, patch_size=2, hidden_size=1024, depth=3, num_heads=4, mlp_ratio=4.0, cross_attn_input_size=512)
m = build(cfg)
g = torch.Generator(device=dev).manual_seed(7)
b = 1
lat = torch.randn(b, 16, 32, 32, device=dev, generator=g).bfloat16()
pos = torch.randn(b, 24, 512, device=dev, generator=g).bfloat16(); neg = torch.zeros_like(pos)
mask = torch.ones(2 * b, 24, device=dev); mask[b:, 17:] = 0
t = torch.full((2 * b,), 0.7, device=dev).bfloat16()
ref = m(torch.cat([lat, lat]), torch.cat([neg, pos]), mask, t)                      # single-GPU path
if sp_group is not None:
    m.enable_sequence_parallel(sp_group)
    got = m(torch.cat([lat, lat]), torch.cat([neg, pos]), mask, t)
    out["sp_forward_rel"] = rel(got, ref)
    if args.fused:
        m.enable_sequence_parallel(sp_group, fused=True)
        for it in range(3):   # repeated calls exercise the monotonic flag epochs / buffer reuse
            got = m(torch.cat([lat, lat]), torch.cat([neg, pos]), mask, t)
        out["sp_fused_forward_rel"] = rel(got, ref)
        _lib.watchdog_ok()
ref_lat = flite_b200.denoise(build(cfg), lat, neg, pos, mask, 3, 6.0)
got_lat = flite_b200.denoise(m, lat, neg, pos, mask, 3, 6.0, cfg_group=cfg_group)
out["denoise_rel"] = rel(got_lat, ref_lat)
_lib.watchdog_ok()

# ---------------- timing at config C4: 10B architecture, one 2048^2 image, CFG 6
if args.big:
    del m
    torch.cuda.empty_cache()
    cfg = dict(in_channels=16, patch_size=2, hidden_size=3072, depth=40, num_heads=12, mlp_ratio=4.0, cross_attn_input_size=4096)
    m = build(cfg)
    m.hoist_context = False
    if sp_group is not None: m.enable_sequence_parallel(sp_group, fused=args.fused)
    R = args.res // 8
    lat = torch.randn(1, 16, R, R, device=dev, generator=g).bfloat16(); acc = lat.clone()
    pos = torch.randn(1, 256, 4096, device=dev, generator=g).bfloat16(); neg = torch.zeros_like(pos)
    if cfg_group is not None:
        half = dist.get_rank(cfg_group)
        ctx = (neg, pos)[half]; msk = torch.ones(1, 256, device=dev); tt = torch.full((1,), 0.9, device=dev).bfloat16()
    else:
        ctx = torch.cat([neg, pos]); msk = torch.ones(2, 256, device=dev); tt = torch.full((2,), 0.9, device=dev).bfloat16()
    step = lambda: flite_b200.denoise_step(m, lat, acc, ctx, msk, tt, 0.01, 6.0, True, cfg_group=cfg_group)
    for _ in range(2): step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 4
    e0.record()
    for _ in range(n): step()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[f"c4_{args.res}_ms_per_step"] = ms.item()
    if args.trace:
        from flite_b200 import ops
        ops.TRACE = []
        e0.record(); step(); e1.record()
        rep = ops.trace_report(); ops.TRACE = None
        out["trace_step_ms"] = e0.elapsed_time(e1)
        out["trace"] = {k: [n, round(ms, 3)] for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    _lib.watchdog_ok()
if rank == 0:
    print("MGPU", json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/mgpu_w{world}_c{args.cfg_ranks}_s{args.sp_ranks}{'_fused' if args.fused else ''}.json", "w"))
dist.destroy_process_group()
