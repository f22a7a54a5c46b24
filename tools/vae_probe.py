"""Where a VAE decode goes (4 images, 1024^2, bf16 channels-last): per-kernel CUDA time of flite_b200.vae.AutoencoderKL.decode
(torch profiler), with the native GroupNorm+SiLU path and, for comparison, torch's own (FLITE_VAE_TORCH_NORM=1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import vae as V
from torch.profiler import profile, ProfilerActivity
dev = "cuda"
torch.manual_seed(0)
m = V.AutoencoderKL().to(dev, torch.bfloat16).to(memory_format=torch.channels_last).eval()
z = torch.randn(4, 16, 128, 128, device=dev, dtype=torch.bfloat16)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    print("decode of 4 images, ms:", timeit(lambda: m.decode(z)))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        m.decode(z); torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=48,
                                                             max_shapes_column_width=70))
