"""Full-size parity: 10B-architecture DiT at 1024^2 (config C2), one CFG-batched forward, against the oracle
restatement of the reference run on the same GPU in bf16 ("the reference bf16 path") and in fp32."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flite_b200
from flite_b200 import _lib
from oracle import dit_oracle, synth

dev = "cuda"
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 40
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = dict(synth.ARCH_10B, depth=depth)
t0 = time.time()
sd = synth.make_state_dict(cfg, 0, device=dev, dtype=torch.bfloat16)
print("weights", time.time() - t0, "s", flush=True)
m = flite_b200.DiT(**cfg)
m.to_empty(device=dev) if False else None
m = m.to(torch.bfloat16)
m.load_state_dict(sd)
m = m.to(dev).eval()
x, ctx, mask = synth.make_inputs(cfg, 1, res, res, 256, valid_len=[200], device=dev)
xb, cb, mb = torch.cat([x, x]).bfloat16(), ctx.bfloat16(), mask.bfloat16()
out = {}
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
for tval in (0.95, 0.5, 0.05):
    t = torch.tensor([tval, tval], device=dev).bfloat16()
    v = m(xb, cb, mb, t)
    _lib.watchdog_ok()
    v_or = dit_oracle.dit_forward(sd, cfg, xb, cb, mb, t)
    r = rel(v, v_or)
    print(f"t={tval}: mine vs oracle bf16 rel-L2 {r:.3e}  std {v_or.float().std().item():.3f}", flush=True)
    out[f"t{tval}_vs_bf16"] = r
# fp32 oracle (weights up-cast from the same bf16 values)
try:
    sd32 = {k: w.float() for k, w in sd.items()}
    t = torch.tensor([0.5, 0.5], device=dev).bfloat16()
    v32 = dit_oracle.dit_forward(sd32, cfg, xb.float(), cb.float(), mb.float(), t, rope_dtype=torch.bfloat16)
    v = m(xb, cb, mb, t); v_or = dit_oracle.dit_forward(sd, cfg, xb, cb, mb, t)
    out["mine_vs_fp32"] = rel(v, v32); out["refbf16_vs_fp32"] = rel(v_or, v32)
    print("vs fp32 oracle: mine", out["mine_vs_fp32"], " reference-bf16", out["refbf16_vs_fp32"])
except Exception as e:
    print("fp32 oracle skipped:", repr(e)[:200])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/parity_c2_depth{depth}_{res}.json", "w"), indent=1)
print(json.dumps(out))
