"""BASELINE.json configs[4] (C5): end-to-end FLitePipeline.__call__ -- 30 Euler steps, CFG 6, 1024^2, VAE decode and
uint8 post-process included -- images/s over all ranks (data parallel over prompts, weights replicated, no collective).

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pipeline_bench.py [--images-per-gpu 4]

Synthetic text embeddings (prompt_embeds=), random-init 10B-architecture DiT and FLUX-architecture VAE decoder
(flite_b200.vae, torch/cuDNN -- library code, see its docstring).  Timed on the device (CUDA events) as the max over
ranks, barrier + synchronize on both sides; one short warm-up call first."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import flite_b200
from flite_b200 import _lib, vae as flite_vae

ap = argparse.ArgumentParser()
ap.add_argument("--images-per-gpu", type=int, default=4)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--repeats", type=int, default=1)
ap.add_argument("--check-full", action="store_true", help="re-run all steps (latent output) for the finite check")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = bench.ARCH_10B
prev = torch.get_default_dtype(); torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = flite_b200.DiT(**cfg)
    vae = flite_vae.AutoencoderKL()
torch.set_default_dtype(prev)
bench.random_init_(model, 0); model.eval()
g = torch.Generator(device=dev).manual_seed(5)
for p in vae.parameters():
    if p.dim() > 1: p.data.copy_((torch.rand(p.shape, device=dev, generator=g) * 2 - 1) * (3.0 / p.shape[1:].numel()) ** 0.5)
vae = vae.to(memory_format=torch.channels_last).eval()
pipe = flite_b200.FLitePipeline(model, vae, None, None)
pipe.set_progress_bar_config(disable=True)
b = args.images_per_gpu
emb = torch.randn((b, 256, cfg["cross_attn_input_size"]), device=dev, generator=g).bfloat16()
call = lambda steps, seed: pipe(prompt=None, height=args.res, width=args.res, num_inference_steps=steps, guidance_scale=6.0,
                                generator=torch.Generator(device=dev).manual_seed(seed + rank), prompt_embeds=emb,
                                output_type="pt").images

call_latent = lambda steps, seed: pipe(prompt=None, height=args.res, width=args.res, num_inference_steps=steps,
                                       guidance_scale=6.0, generator=torch.Generator(device=dev).manual_seed(seed + rank),
                                       prompt_embeds=emb, output_type="latent").images

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

call(2, 0)                                   # warm-up: kernels configured, cuDNN algorithms chosen, workspaces allocated
barrier()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
times, decode_ms = [], []
for r in range(args.repeats):
    barrier()
    t0 = time.perf_counter(); e0.record()
    imgs = call(args.steps, 100 + r)         # returns uint8 images on the host (the pipeline's .cpu() synchronises)
    e1.record(); barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    times.append(ms.item())
# decode + post-process alone (same latents shape), for the breakdown
lat = torch.randn((b, 16, args.res // 8, args.res // 8), device=dev, generator=g).bfloat16()
from flite_b200 import ops
barrier(); e0.record()
z = ops.latent_unscale(lat, vae.config.scaling_factor, vae.config.shift_factor)
u8 = ops.image_to_uint8(vae.decode(z).sample).cpu()
e1.record(); barrier()
_lib.watchdog_ok()
assert imgs.shape == (b, 3, args.res, args.res) and imgs.dtype == torch.uint8
# sanity of what was produced (per rank): the timed call's images and a latent-only rerun of the same seed
lat_chk = call_latent(args.steps if args.check_full else 2, 100)
stats = torch.tensor([imgs.float().mean().item(), float(torch.isfinite(lat_chk.float()).all().item()),
                      lat_chk.float().std().item(), u8.float().mean().item()], device=dev, dtype=torch.float64)
all_stats = [torch.zeros_like(stats) for _ in range(world)]
if world > 1: dist.all_gather(all_stats, stats)
else: all_stats = [stats]
best = min(times)
out = {"metric": "end-to-end FLitePipeline images/s (C5)", "value": world * b / (best / 1e3), "unit": "images/s", "n_gpus": world,
       "images_per_gpu": b, "steps": args.steps, "res": args.res, "s_per_call": best / 1e3, "calls_timed": times,
       "decode_postprocess_ms": e0.elapsed_time(e1), "per_rank_[image_mean, latents_finite, latents_std, decode_only_image_mean]": [[round(v, 3) for v in st.tolist()] for st in all_stats],
       "includes": "latent init, 30 CFG-batched denoise steps (context K/V hoisted out of the step loop), latent unscale, "
                   "VAE decode (torch/cuDNN), uint8 post-process, D2H of the images"}
if rank == 0:
    print("PIPE", json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/pipeline_c5_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
