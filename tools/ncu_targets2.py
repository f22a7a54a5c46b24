"""ncu target: the persistent attention kernel (round-robin schedule) at C2 next to the one-cluster-per-unit launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flite_b200 import _lib, ops
dev = "cuda"
lib = _lib.load()
_lib.check(lib.flite_check_device(), "flite_check_device")
g = torch.Generator(device=dev).manual_seed(0)
B, H, L = 2, 12, 4112
d = H * 256
qkv = torch.randn(B * L, 3 * d, device=dev, generator=g).bfloat16()
q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
cu = (torch.arange(B + 1, dtype=torch.int32) * L).to(dev)
o = torch.empty(B * L, d, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    ops.attention_varlen(q, k, v, cu, cu, H, L, 256 ** -0.5, out=o, variant=5)
modes = [int(m) for m in os.environ.get("SK_MODES", "1").split(",")]
for mode in modes:
    lib.flite_set_tuning(15, mode)
    for _ in range(2):
        ops.attention_streamk(q, k, v, cu, cu, H, L, L, 256 ** -0.5, out=o)
torch.cuda.synchronize()
_lib.watchdog_ok()
print("ok")
