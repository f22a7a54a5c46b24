import sys, time, torch
sys.path.insert(0, "/root/repo")
from flite_b200 import vae as V
dev = "cuda"
torch.manual_seed(0)
m = V.AutoencoderKL().to(dev, torch.bfloat16).eval()
z = torch.randn(4, 16, 128, 128, device=dev, dtype=torch.bfloat16)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    ref = m.decode(z).sample.float()
    print("baseline ms", timeit(lambda: m.decode(z)))
    m.enable_slicing(); print("slicing ms", timeit(lambda: m.decode(z))); m.disable_slicing()
    torch.backends.cudnn.benchmark = True
    print("cudnn.benchmark ms", timeit(lambda: m.decode(z)))
    m.decoder.to(memory_format=torch.channels_last)
    out = m.decode(z).sample.float()
    print("channels_last weights ms", timeit(lambda: m.decode(z)), "rel", ((out - ref).norm() / ref.norm()).item())
    torch.backends.cudnn.benchmark = False
    print("channels_last weights, no benchmark ms", timeit(lambda: m.decode(z)))
    # per-op profile of the fastest
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.decode(z); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
