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


def _ulysses_worker(rank, world, port, out):
    """The layout contract of the sequence-parallel path (model.py, DiT.forward with sp_group): the QKV epilogue writes
    [sample][dest rank][local token][q|k|v of that rank's heads]; one all_to_all_single per sample must then leave every
    rank with [full sequence][q|k|v of MY heads]; the return exchange + permute must give back [local token][all heads]."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp_group, cfg_group, rep, nrep = parallel.make_groups(cfg_ranks=1, sp_ranks=world)
    P, B, L, H, hd = world, 2, 8, 4, 3                    # tiny stand-ins: 4 heads of width 3
    Lq, hq = L // P, H // P
    dq = hq * hd
    g = torch.Generator().manual_seed(0)
    qkv_full = torch.randn(B, L, 3, H, hd, generator=g)   # what a single GPU would hold
    mine = qkv_full[:, rank * Lq:(rank + 1) * Lq]          # this rank's tokens, all heads
    # send layout written by the EPI_QKV_ROPE epilogue with sp_ranks > 0 (tests/test_kernels_gpu.py checks the kernel
    # against exactly this permutation)
    send = mine.reshape(B, Lq, 3, P, dq).permute(0, 3, 1, 2, 4).contiguous()          # [B, P, Lq, 3, dq]
    recv = torch.empty(B, L, 3, dq)
    for b in range(B):
        dist.all_to_all_single(recv[b].view(P, Lq * 3 * dq), send[b].view(P, Lq * 3 * dq), group=sp_group)
    want = qkv_full[:, :, :, rank * hq:(rank + 1) * hq].reshape(B, L, 3, dq)          # full sequence, my heads
    ok = torch.equal(recv, want)
    # "attention" = identity on q; return path: [full sequence][my heads] -> token owners, then [P, Lq, dq] -> [Lq, P*dq]
    ao_full = recv[:, :, 0].contiguous()                                              # [B, L, dq]
    ao_recv = torch.empty(B, P, Lq * dq)
    for b in range(B):
        dist.all_to_all_single(ao_recv[b], ao_full[b].view(P, Lq * dq), group=sp_group)
    back = ao_recv.view(B, P, Lq, dq).permute(0, 2, 1, 3).reshape(B, Lq, P * dq)       # permute_021 per sample
    ok = ok and torch.equal(back, mine[:, :, 0].reshape(B, Lq, H * hd))
    ok = ok and cfg_group is None and nrep == 1 and dist.get_world_size(sp_group) == world
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_ulysses_exchange_layout_contract_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ulysses_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _groups_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp_group, cfg_group, rep, nrep = parallel.make_groups(cfg_ranks=2, sp_ranks=1)
    ok = sp_group is None and cfg_group is not None and dist.get_world_size(cfg_group) == 2
    ok = ok and dist.get_rank(cfg_group) == rank and (rep, nrep) == (0, 1)
    try:
        parallel.make_groups(cfg_ranks=2, sp_ranks=2)      # 4 ranks per replica do not divide a world of 2
        ok = False
    except ValueError:
        pass
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_make_groups_cfg_split_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_groups_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
