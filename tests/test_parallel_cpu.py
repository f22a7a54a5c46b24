"""CPU (gloo, world_size 2): the data-parallel prompt sharding used for configs C3 / C5."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flite_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_denoise(latents, neg, pos, mask, scale=1.0):
    # stands in for flite_b200.denoise: per-image function of (latent, its two context rows, its two mask rows)
    b = latents.shape[0]
    m = torch.ones(2 * b, 1) if mask is None else mask.float().sum(1, keepdim=True)
    s = (pos.float().mean((1, 2)) - neg.float().mean((1, 2)) + m[:b, 0] + 2 * m[b:, 0]).view(b, 1, 1, 1)
    return latents * scale + s


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(B, 4, 3, 3, generator=g)
    neg, pos = torch.randn(B, 5, 6, generator=g), torch.randn(B, 5, 6, generator=g)
    mask = (torch.rand(2 * B, 5, generator=g) > 0.4).float()
    full = _fake_denoise(lat, neg, pos, mask, scale=0.5)
    got = parallel.dp_denoise(_fake_denoise, lat, neg, pos, mask, gather=True, scale=0.5)
    local = parallel.dp_denoise(_fake_denoise, lat, neg, pos, mask, gather=False, scale=0.5)
    lo, hi = parallel.shard_range(B, rank, world)
    ok = torch.allclose(got, full) and torch.allclose(local, full[lo:hi])
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [1, 4, 5])
def test_dp_denoise_matches_single_process(B):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
