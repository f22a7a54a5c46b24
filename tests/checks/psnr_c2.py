"""north_star image criterion at config C2: 10B architecture, 1024x1024, 30 Euler steps, CFG 6, batch 1.
New path trajectory vs the reference bf16 trajectory (oracle on the same GPU), both decoded by the same decoder."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flite_b200
from oracle import dit_oracle, sampler_oracle, synth, vae_decoder
dev = "cuda"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cfg = dict(synth.ARCH_10B)
sd = synth.make_state_dict(cfg, 0, device=dev, dtype=torch.bfloat16)
m = flite_b200.DiT(**cfg).to(torch.bfloat16); m.load_state_dict(sd); m = m.to(dev).eval()
x, ctx, mask = synth.make_inputs(cfg, 1, 1024, 1024, 256, valid_len=[200], device=dev)
xb, cb, mb = x.bfloat16(), ctx.bfloat16(), mask.bfloat16()
t0 = time.time(); trace = []
lat = flite_b200.denoise(m, xb, cb[:1], cb[1:], mask, steps, 6.0, trace=trace); torch.cuda.synchronize(); t_mine = time.time() - t0
t0 = time.time(); otrace = []
fn = lambda *a: dit_oracle.dit_forward(sd, cfg, *a)
olat = sampler_oracle.sample_pipeline(fn, xb, cb[:1], cb[1:], mb, steps, 6.0, trace=otrace); torch.cuda.synchronize(); t_ref = time.time() - t0
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
dec = vae_decoder.make_decoder(0, dev)
img, ref = vae_decoder.decode_to_image(dec, lat), vae_decoder.decode_to_image(dec, olat)
out = {"steps": steps, "final_latent_rel_l2": rel(lat, olat), "psnr_db": vae_decoder.psnr(img, ref), "image_std": ref.std().item(),
       "seconds_new_path": t_mine, "seconds_oracle_gpu": t_ref,
       "step0_velocity_rel_l2": rel(trace[0][:1] + 6.0 * (trace[0][1:] - trace[0][:1]), otrace[0])}
print(json.dumps(out)); os.makedirs("gpurun_out", exist_ok=True); json.dump(out, open("gpurun_out/psnr_c2.json", "w"), indent=1)
