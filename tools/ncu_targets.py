"""A short program for `ncu --set full`: the self-attention kernel (default and K3 variants) at the C2 shape, and the
down-projection / projection / QKV GEMM shapes through the own kernel and through torch.matmul (cuBLASLt), two launches each.
   ncu --set full --import-source on -k regex:'attn_fwd|gemm_bf16|nvjet|cutlass|xmma' -o gpurun_out/r2p python tools/ncu_targets.py"""
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
for variant in (5, 10):
    for _ in range(2):
        ops.attention_varlen(q, k, v, cu, cu, H, L, 256 ** -0.5, out=o, variant=variant)
M = 8224
for name, (n, kk) in {"down": (3072, 12288), "proj": (3072, 3072), "qkv": (9216, 3072)}.items():
    a = (torch.randn(M, kk, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(n, kk, device=dev, generator=g) * 0.02).bfloat16()
    c = torch.empty(M, n, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        ops.gemm(a, w, None, out=c)
    for _ in range(2):
        torch.matmul(a, w.t(), out=c)
torch.cuda.synchronize()
_lib.watchdog_ok()
print("ok")
