"""Multi-GPU parity of the sequence-parallel / CFG-split paths, run through torchrun when the box has at
least two GPUs (skipped on a single-GPU box; the same checks were run by hand with tools/mgpu_check.py on 2 / 4 / 8 GPUs,
profiles/mgpu_w*.json).  Every layout must reproduce the single-GPU result bit for bit."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, *args):
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), *args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU ")][-1]
    return json.loads(line[5:])


@pytest.mark.parametrize("fused", [False, True])
def test_ulysses_two_gpus_bit_identical(fused):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, "--sp-ranks", "2", *(["--fused"] if fused else []))
    assert out["sp_forward_rel"] == 0.0 and out["denoise_rel"] == 0.0
    if fused:
        assert out["sp_fused_forward_rel"] == 0.0


def test_cfg_split_two_gpus_bit_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, "--cfg-ranks", "2")
    assert out["denoise_rel"] == 0.0


def test_cfg_split_times_ulysses_four_gpus_bit_identical():
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    out = _run(4, "--cfg-ranks", "2", "--sp-ranks", "2", "--fused")
    assert out["sp_fused_forward_rel"] == 0.0 and out["denoise_rel"] == 0.0
