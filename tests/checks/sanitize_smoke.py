"""Small end-to-end run for compute-sanitizer (memcheck / initcheck): tiny DiT, 2 denoise steps (CFG and APG), toy
decode tail, plus the peer-memory kernels in single-GPU loopback."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flite_b200
from flite_b200 import ops, _lib
from oracle import synth, vae_decoder
dev = "cuda"
cfg = dict(synth.TINY)
sd = synth.make_state_dict(cfg, 0, device=dev)
m = flite_b200.DiT(**cfg); m.load_state_dict(sd); m = m.to(dev, torch.bfloat16).eval()
x, ctx, mask = synth.make_inputs(cfg, 2, 128, 192, 24, [17, 24], 1234, device=dev)
b = 2
for hoist in (True, False):
    m.hoist_context = hoist
    lat = flite_b200.denoise(m, x.bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, 2, 6.0)
    lat2 = flite_b200.denoise(m, x.bfloat16(), ctx[:b].bfloat16(), ctx[b:].bfloat16(), mask, 2, 6.0,
                              apg_config=flite_b200.APGConfig(True, 0.03))
z = ops.latent_unscale(lat.contiguous(), 0.3611, 0.1159)
u8 = ops.image_to_uint8(vae_decoder.toy_decode(z).contiguous())
# loopback peer kernels
P, B, L, H = 2, 2, 272, 4
d, Hp, Lq = H * 256, H // P, L // P
dq = d // P
g = torch.Generator(device=dev).manual_seed(1)
xx = (torch.randn(B * L, d, device=dev, generator=g) * 0.5).bfloat16()
w = (torch.randn(3 * d, d, device=dev, generator=g) * 0.05).bfloat16(); bias = torch.randn(3 * d, device=dev, generator=g).bfloat16()
ang = torch.rand(L, 128, device=dev, generator=g) * 6.28
cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
cu = (torch.arange(0, B + 1, dtype=torch.int32) * L).to(dev)
recv = [torch.zeros(B * L, 3 * dq, device=dev, dtype=torch.bfloat16) for _ in range(P)]
ao = [torch.zeros(B * Lq, d, device=dev, dtype=torch.bfloat16) for _ in range(P)]
tab = lambda bufs: (ctypes.c_void_p * 8)(*([t.data_ptr() for t in bufs] + [None] * (8 - len(bufs))))
for r in range(P):
    xr = xx.view(B, L, d)[:, r * Lq:(r + 1) * Lq].reshape(B * Lq, d).contiguous()
    ops.gemm_qkv_p2p(xr, w, bias, cos[r * Lq:(r + 1) * Lq].contiguous(), sin[r * Lq:(r + 1) * Lq].contiguous(), Lq, P, Hp, r, L, tab(recv))
for r in range(P):
    rr = recv[r]
    ops.attention_varlen_p2p(rr[:, :dq], rr[:, dq:2 * dq], rr[:, 2 * dq:], cu, cu, Hp, L, 256 ** -0.5, tab(ao), P, Lq, r * Hp, d)
torch.cuda.synchronize()
_lib.watchdog_ok()
print("finite", torch.isfinite(lat.float()).all().item(), torch.isfinite(lat2.float()).all().item(), "u8 mean", u8.float().mean().item())
