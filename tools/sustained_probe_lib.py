"""Shared helper of the sustained-rate probes: run `fn` back to back for `seconds`, sampling nvidia-smi."""
import subprocess
import time

import torch


def sample_loop(fn, seconds, gpu_index=0):
    """Returns (ms per call, median SM MHz, median board W) over the second half of the loop (steady state)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=clocks.sm,power.draw",
                             "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                            stderr=subprocess.DEVNULL, text=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    proc.terminate()
    out, _ = proc.communicate(timeout=5)
    ms = e0.elapsed_time(e1) / n
    lines = out.strip().splitlines()
    clk, pw = [], []
    for line in lines[len(lines) // 2:]:
        f = [v.strip() for v in line.split(",")]
        try:
            clk.append(float(f[0])); pw.append(float(f[1]))
        except Exception:
            pass
    clk.sort(); pw.sort()
    return ms, (clk[len(clk) // 2] if clk else None), (pw[len(pw) // 2] if pw else None)
