"""GPU parity tests, kernel by kernel, through the C ABI, against torch restatements of the reference ops."""
import math

import numpy as np

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    from flite_b200 import _lib
    assert torch.cuda.is_available()
    _lib.check(_lib.load().flite_check_device(), "flite_check_device")
    yield
    _lib.watchdog_ok()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).bfloat16()


# ----------------------------------------------------------------------------- sampler kernel
@pytest.mark.parametrize("numel", [8, 16 * 32 * 32, 16 * 128 * 128, 3 * 16 * 112 * 168])
def test_cfg_euler_bit_exact_bf16(numel):
    from flite_b200 import ops
    u, c, acc = rnd(numel, seed=1), rnd(numel, seed=2), rnd(numel, seed=3)
    acc0 = acc.clone()
    lat = torch.empty_like(acc)
    ops.cfg_euler(acc, u, c, 6.0, 0.0123, lat)
    ref = acc0 + 0.0123 * (u + 6.0 * (c - u))          # pipeline.py:290,296 in bf16
    assert torch.equal(acc, ref) and torch.equal(lat, ref)
    # no-CFG branch (guidance < 1, pipeline.py:291-293)
    acc = acc0.clone()
    ops.cfg_euler(acc, None, c, 0.5, 0.05, lat, do_cfg=False)
    assert torch.equal(acc, acc0 + 0.05 * c)


def test_cfg_euler_fp32_accumulator():
    from flite_b200 import ops
    n = 16 * 128 * 128
    u, c = rnd(n, seed=1), rnd(n, seed=2)
    acc = torch.randn(n, device=DEV)
    acc0 = acc.clone()
    lat = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ops.cfg_euler(acc, u, c, 6.0, 0.0123, lat)
    v = u + 6.0 * (c - u)                                # train.py:596 (bf16 tensors)
    ref = acc0 + 0.0123 * v.to(torch.float32)            # train.py:599
    assert (acc - ref).abs().max().item() <= 1e-6
    assert torch.equal(lat, acc.bfloat16())


# ----------------------------------------------------------------------------- norm + modulate
@pytest.mark.parametrize("rows,d,B", [(2 * 272, 512, 2), (2 * 4112, 3072, 2), (5, 256, 1), (37, 4096, 1), (9, 4608, 3)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_rmsnorm_modulate(rows, d, B, mode):
    from flite_b200 import ops
    x = rnd(rows, d, seed=1)
    w = (1 + 0.1 * torch.randn(d, device=DEV)).bfloat16()
    mod = rnd(B, 9 * d, scale=0.5, seed=2)
    sc, sh = mod[:, d:2 * d], mod[:, :d]
    L = (rows + B - 1) // B
    y = ops.rmsnorm_modulate(x, w if mode else None, mode, sc, sh, rows_per_sample=L)
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
    if mode == 1:
        n = (xf * rstd).bfloat16() * w
    elif mode == 2:
        n = (xf * rstd * w).bfloat16()
    else:
        n = (xf * rstd).bfloat16()
    idx = torch.arange(rows, device=DEV) // L
    ref = n * (1 + sc[idx]) + sh[idx]
    assert rel(y, ref) <= 2e-4
    assert (y == ref).float().mean().item() > 0.999       # only rstd ulp differences flip a rounding
    y2 = ops.rmsnorm_modulate(x, w if mode else None, mode)      # no modulation (context_norm)
    assert rel(y2, n) <= 2e-4


# ----------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 128, 64), (1, 128, 64), (2, 9216, 512), (2, 27648, 3072), (8, 3072, 12288), (3, 192, 1024), (200, 64, 512), (256, 512, 512), (333, 768, 192),
               (2 * 272, 1536, 512), (2 * 4112, 3072, 3072), (16 * 1181, 768, 4096)]   # last: multi-band rasterisation


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_gemm_store_bias(M, N, K, variant):
    from flite_b200 import ops
    need = {1: 256, 2: 256, 3: 128, 4: 64}.get(variant, 64)
    if N % need:
        pytest.skip("N not a multiple of this variant's tile")
    if variant == 5 and M > 8:
        pytest.skip("the GEMV variant handles M <= 8")
    a, w, b = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
    out = ops.gemm(a, w, b, variant=variant)
    assert rel(out, F.linear(a, w, b)) <= 1e-3
    out = ops.gemm(a, w, None, act=1, variant=variant)
    assert rel(out, F.silu(F.linear(a, w))) <= 2e-3


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("M,B", [(2 * 272, 2), (2 * 4112, 2), (3 * 100, 3)])
def test_gemm_gated_residual_in_place(variant, M, B):
    from flite_b200 import ops
    N = K = 512 if M < 4000 else 3072
    a, w = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2)
    resid, gate = rnd(M, N, seed=3), rnd(B, 9 * N, seed=4)[:, 2 * N:3 * N]
    ref = resid + F.linear(a, w) * gate.repeat_interleave(M // B, 0)      # model.py:289
    x = resid.clone()
    ops.gemm(a, w, None, epilogue=ops.EPI_GATED_RES, resid=x, gate=gate, rows_per_sample=M // B, variant=variant, out=x)
    assert rel(x, ref) <= 1e-3


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_gemm_swiglu(variant):
    from flite_b200 import ops
    M, K, inter = 2 * 272, 512, 2048
    a = rnd(M, K, scale=0.5, seed=1)
    wg, wu = rnd(inter, K, scale=0.05, seed=2), rnd(inter, K, scale=0.05, seed=3)
    out = ops.gemm(a, ops.interleave_gate_up(wg, wu), None, epilogue=ops.EPI_SWIGLU, variant=variant)
    g, u = F.linear(a, wg), F.linear(a, wu)
    ref = F.silu(g.float()).bfloat16() * u                                 # liger swiglu
    assert out.shape == (M, inter) and rel(out, ref) <= 2e-3


def _rope_norm_ref(qkv, cos, sin, n_heads_rot, B):
    M = qkv.shape[0]
    x = qkv[:, :n_heads_rot * 256].reshape(M, n_heads_rot, 256).float()
    if cos is not None:
        c, s = cos.float().repeat(B, 1)[:, None, :], sin.float().repeat(B, 1)[:, None, :]
        x1, x2 = x[..., :128], x[..., 128:]
        x = torch.cat([x1 * c + x2 * s, x1 * (-s) + x2 * c], -1).bfloat16().float()     # model.py:403-414
    y = (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16()            # model.py:101-108
    return torch.cat([y.reshape(M, -1), qkv[:, n_heads_rot * 256:]], 1)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("with_rope", [True, False])
def test_gemm_qkv_rope_qknorm_epilogue(variant, with_rope):
    from flite_b200 import ops
    B, L, d = 2, 272, 512
    M = B * L
    a, w, b = rnd(M, d, scale=0.5, seed=1), rnd(3 * d, d, scale=0.05, seed=2), rnd(3 * d, seed=3)
    ang = torch.rand(L, 128, device=DEV) * 6.28
    cos, sin = (ang.cos().bfloat16(), ang.sin().bfloat16()) if with_rope else (None, None)
    out = ops.gemm(a, w, b, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d,
                   rows_per_sample=L, variant=variant)
    ref = _rope_norm_ref(F.linear(a, w, b), cos, sin, 2 * d // 256, B)
    assert rel(out, ref) <= 2e-4
    # the unfused kernel gives the same thing
    plain = ops.gemm(a, w, b, variant=variant)
    ops.rope_qknorm_(plain, 2 * d // 256, cos, sin, rows_per_sample=L)
    assert rel(plain, ref) <= 2e-4


# ----------------------------------------------------------------------------- attention
ATT_CASES = [(1, 1, 128, 128), (1, 2, 256, 384), (2, 2, 272, 272), (2, 2, 272, 17), (3, 1, 100, 1), (1, 3, 500, 129),
             (2, 12, 4112, 256)]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 10])
@pytest.mark.parametrize("B,H,Lq,Lk", ATT_CASES)
def test_attention_uniform(B, H, Lq, Lk, variant):
    from flite_b200 import ops
    from oracle.dit_oracle import flash_attn_varlen
    d = H * 256
    q = rnd(B * Lq, d, seed=1)
    kv = rnd(B * Lk, 2 * d, seed=2)
    cu_q = torch.arange(B + 1, device=DEV, dtype=torch.int32) * Lq
    cu_k = torch.arange(B + 1, device=DEV, dtype=torch.int32) * Lk
    out = ops.attention_varlen(q, kv[:, :d], kv[:, d:], cu_q, cu_k, H, Lq, 256 ** -0.5, variant=variant)
    ref = flash_attn_varlen(q.view(-1, H, 256), kv[:, :d].reshape(-1, H, 256), kv[:, d:].reshape(-1, H, 256),
                            cu_q, cu_k, 256 ** -0.5)
    assert rel(out, ref.reshape(-1, d)) <= 5e-3          # bf16 output + bf16 P, same as FA2's own error


@pytest.mark.parametrize("variant", [1, 3, 4, 5, 6, 7, 8, 10])
def test_attention_ragged_and_empty_keys(variant):
    from flite_b200 import ops
    from oracle.dit_oracle import flash_attn_varlen
    H, d = 2, 512
    lens_q, lens_k = [130, 272, 5], [17, 0, 300]           # an EMPTY key sequence -> zeros, like flash-attn
    cu_q = torch.tensor([0] + list(torch.tensor(lens_q).cumsum(0)), device=DEV, dtype=torch.int32)
    cu_k = torch.tensor([0] + list(torch.tensor(lens_k).cumsum(0)), device=DEV, dtype=torch.int32)
    q, kv = rnd(sum(lens_q), d, seed=1), rnd(sum(lens_k), 2 * d, seed=2)
    out = ops.attention_varlen(q, kv[:, :d], kv[:, d:], cu_q, cu_k, H, max(lens_q), 256 ** -0.5, variant=variant)
    ok = torch.cat([torch.arange(0, 130), torch.arange(402, 407)]).to(DEV)
    cu_q2 = torch.tensor([0, 130, 135], device=DEV, dtype=torch.int32)
    cu_k2 = torch.tensor([0, 17, 317], device=DEV, dtype=torch.int32)
    ref = flash_attn_varlen(q[ok].view(-1, H, 256), kv[:, :d].reshape(-1, H, 256), kv[:, d:].reshape(-1, H, 256),
                            cu_q2, cu_k2, 256 ** -0.5)
    assert rel(out[ok], ref.reshape(-1, d)) <= 5e-3
    assert out[130:402].abs().max().item() == 0


@pytest.mark.parametrize("q_lens,k_lens,H", [([4112, 4112], [256, 256], 12), ([4112, 4112], [77, 256], 12),
                                             ([300, 1, 272, 512, 4112], [129, 0, 272, 40, 1], 3),
                                             ([1000] * 9, [256, 1, 0, 512, 300, 17, 128, 129, 255], 12)])
def test_attention_persistent_ragged_bit_equal_to_per_unit_kernel(q_lens, k_lens, H):
    """FLITE_ATTN_PERSISTENT (variant 10): one wave of clusters walks whole (sequence, head, 256-query tile) units
    round-robin with per-sequence query / key lengths (the cross-attention over the packed text context,
    model.py:188-210).  Same arithmetic per unit as variant 5, so the output must be bit-identical -- ragged query
    tails, sequences shorter than a tile, an empty key sequence (zeros) and more units than clusters included -- and
    nothing may be written outside the rows of the call."""
    from flite_b200 import _lib, ops
    cu_q = torch.tensor([0] + list(np.cumsum(q_lens)), dtype=torch.int32, device=DEV)
    cu_k = torch.tensor([0] + list(np.cumsum(k_lens)), dtype=torch.int32, device=DEV)
    nq, nk, d = sum(q_lens), sum(k_lens), H * 256
    q, k, v = rnd(nq, d, seed=1), rnd(max(nk, 1), d, seed=2)[:nk], rnd(max(nk, 1), d, seed=3)[:nk]
    base = ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=5)
    out = torch.full((nq + 4, d), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=10, out=out[:nq])
    again = ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=10)
    _lib.watchdog_ok()
    assert torch.equal(out[:nq], base) and torch.equal(again, base)
    assert bool((out[nq:] == 7.0).all())


def test_attention_self_from_qkv_buffer_full_size():
    """C2 size (2 x 12 heads x 4112^2 x 256) -- checked through properties instead of an fp32 reference:
    rows of softmax sum to one (V = ones -> output ones) and permuting keys leaves the output unchanged."""
    from flite_b200 import ops
    B, H, L, d = 2, 12, 4112, 3072
    qkv = rnd(B * L, 3 * d, seed=1)
    cu = torch.arange(B + 1, device=DEV, dtype=torch.int32) * L
    out = ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, 256 ** -0.5)
    qkv1 = qkv.clone()
    qkv1[:, 2 * d:] = 1.0
    ones = ops.attention_varlen(qkv1[:, :d], qkv1[:, d:2 * d], qkv1[:, 2 * d:], cu, cu, H, L, 256 ** -0.5)
    assert (ones.float() - 1).abs().max().item() <= 1e-2
    perm = torch.cat([torch.randperm(L, device=DEV) + b * L for b in range(B)])
    qkv2 = qkv.clone()
    qkv2[:, d:] = qkv[perm][:, d:]
    out2 = ops.attention_varlen(qkv2[:, :d], qkv2[:, d:2 * d], qkv2[:, 2 * d:], cu, cu, H, L, 256 ** -0.5)
    assert rel(out2, out) <= 5e-3


# ----------------------------------------------------------------------------- small ops
@pytest.mark.parametrize("B,H,W,d", [(2, 64, 96, 512), (1, 32, 32, 512), (2, 128, 128, 3072)])
def test_patch_embed_and_unpatchify(B, H, W, d):
    from flite_b200 import ops
    C, P = 16, 2
    x = rnd(B, C, H, W, seed=1)
    w, b, reg = rnd(d, C, P, P, scale=0.1, seed=2), rnd(d, seed=3), rnd(1, 16, d, seed=4)
    tok = ops.patch_embed(x, w, b, reg, P)                       # im2col gather + tcgen05 GEMM (C*P*P = 64)
    ref = F.conv2d(x.float(), w.float(), b.float(), stride=P).flatten(2).transpose(1, 2)   # model.py:324-328
    ref = torch.cat([reg.float().repeat(B, 1, 1), ref], 1).reshape(-1, d)
    assert rel(tok, ref) <= 3e-3
    from flite_b200 import _lib
    lib = _lib.load()
    lib.flite_set_tuning(11, 1)                                  # the CUDA-core kernel gives the same tokens
    try:
        tok_old = ops.patch_embed(x, w, b, reg, P)
    finally:
        lib.flite_set_tuning(11, 0)
    assert rel(tok_old, ref) <= 3e-3 and rel(tok, tok_old) <= 3e-3
    L = 16 + (H // P) * (W // P)
    assert torch.equal(tok.view(B, L, d)[:, :16], reg.repeat(B, 1, 1))
    tk = rnd(B * L, 64, seed=5)
    o = ops.unpatchify(tk, B, C, H, W, P, 16)
    r = tk.view(B, L, 64)[:, 16:].view(B, H // P, W // P, P, P, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)
    assert torch.equal(o, r)                                                               # model.py:583-590


@pytest.mark.parametrize("tdtype", [torch.bfloat16, torch.float32, torch.float16])
def test_timestep_embedding(tdtype):
    from flite_b200 import ops
    d = 3072
    t = torch.tensor([0.9333, 0.25, 1.0, 0.0333], device=DEV).to(tdtype)
    freqs = torch.exp(-math.log(10000) * torch.arange(0, d // 2, dtype=torch.float32) / (d // 2)).to(DEV)
    if tdtype == torch.float16:
        e = ops.timestep_embed((t * 1000).float(), 2, freqs, d)
    else:
        e = ops.timestep_embed(t.float(), int(tdtype == torch.bfloat16), freqs, d)
    args = (t * 1000)[:, None].float() * freqs[None]                                       # model.py:25,551
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1).bfloat16()
    assert (e.float() - ref.float()).abs().max().item() <= 2 ** -8


def test_pack_context_matches_nonzero_index_select():
    from flite_b200 import ops
    B, Lc, d = 3, 300, 512
    ctx = rnd(B, Lc, d, seed=1)
    mask = (torch.rand(B, Lc, device=DEV) > 0.3).float()
    mask[1] = 0                                                                             # empty sequence
    packed, cu = ops.pack_context(ctx, mask)
    idx = mask.reshape(-1).nonzero().flatten()                                              # model.py:61-62
    assert torch.equal(packed[: idx.numel()], ctx.reshape(-1, d)[idx])
    assert packed[idx.numel():].abs().max().item() == 0
    assert cu.tolist() == [0] + mask.sum(1).cumsum(0).int().tolist()                        # model.py:50-56


# ----------------------------------------------------------------------------- sequence-parallel plumbing kernels
@pytest.mark.parametrize("P", [2, 4])
def test_qkv_epilogue_head_scatter_matches_all_to_all_layout(P):
    """Ulysses: the QKV epilogue writes [sample][dest rank][local token][q|k|v][head % Hp][256] directly."""
    from flite_b200 import ops
    B, Lq, H = 2, 136, 4
    d, Hp = H * 256, H // P
    dq = d // P
    M = B * Lq
    a, w, b = rnd(M, d, scale=0.5, seed=1), rnd(3 * d, d, scale=0.05, seed=2), rnd(3 * d, seed=3)
    ang = torch.rand(Lq, 128, device=DEV) * 6.28
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    plain = ops.gemm(a, w, b, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=Lq)
    send = torch.zeros(B * P * Lq, 3 * dq, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=Lq,
             out=send, sp_ranks=P, sp_heads_per_rank=Hp)
    # reference layout: plain [B, Lq, 3, P, Hp*256] -> [B, P, Lq, 3, Hp*256]
    ref = plain.view(B, Lq, 3, P, dq).permute(0, 3, 1, 2, 4).reshape(B * P * Lq, 3 * dq)
    assert torch.equal(send, ref)


def test_patch_embed_token_slice_and_permute():
    from flite_b200 import ops
    B, C, H, W, P, d = 2, 16, 32, 64, 2, 512
    x = rnd(B, C, H, W, seed=1)
    w, b, reg = rnd(d, C, P, P, scale=0.1, seed=2), rnd(d, seed=3), rnd(1, 16, d, seed=4)
    L = 16 + (H // P) * (W // P)
    full = ops.patch_embed(x, w, b, reg, P).view(B, L, d)
    for ranks in (2, 4):
        Lq = L // ranks
        for r in range(ranks):
            part = ops.patch_embed(x, w, b, reg, P, tok_offset=r * Lq, tok_count=Lq).view(B, Lq, d)
            assert torch.equal(part, full[:, r * Lq:(r + 1) * Lq])
    t = rnd(3, 5, 64, seed=5)
    assert torch.equal(ops.permute_021(t), t.permute(1, 0, 2).contiguous())


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("M", [2 * 272, 300, 2 * 4112])
def test_qkv_epilogue_staged_stores_bit_equal(variant, M):
    """FLITE_TUNE_QKV_STAGED_STORES: the shared-memory transposed store path (used for NVLink peer stores) writes
    exactly what the one-row-per-thread path writes, ragged last M tile included."""
    from flite_b200 import _lib, ops
    d, L = 512, M // 2 if M % 2 == 0 else M
    B = M // L
    a, w, b = rnd(M, d, scale=0.5, seed=1), rnd(3 * d, d, scale=0.05, seed=2), rnd(3 * d, seed=3)
    ang = torch.rand(L, 128, device=DEV) * 6.28
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    kw = dict(epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=L, variant=variant)
    plain = ops.gemm(a, w, b, **kw)
    lib = _lib.load()
    lib.flite_set_tuning(6, 1)
    try:
        staged = torch.full((M + 8, 3 * d), 7.0, device=DEV, dtype=torch.bfloat16)   # guard rows catch overruns
        ops.gemm(a, w, b, out=staged[:M], **kw)
    finally:
        lib.flite_set_tuning(6, 0)
    assert torch.equal(staged[:M], plain)
    assert bool((staged[M:] == 7.0).all())


def _peer_table(bufs):
    import ctypes
    return (ctypes.c_void_p * 8)(*([t.data_ptr() for t in bufs] + [None] * (8 - len(bufs))))


@pytest.mark.parametrize("P", [2, 4])
def test_fused_ulysses_kernels_single_gpu_loopback(P):
    """The peer-memory kernels with every "peer" buffer on this GPU: P emulated ranks run flite_gemm_qkv_p2p /
    flite_attention_varlen_p2p in turn; the receive buffers must equal what the all-to-all path produces and the
    returned attention rows must equal single-GPU attention (SURVEY.md 8e, C4 layout)."""
    from flite_b200 import ops
    B, L, H = 2, 272, 4
    d, Hp, Lq = H * 256, H // P, L // P
    dq = d // P
    x = rnd(B * L, d, scale=0.5, seed=1)
    w, b = rnd(3 * d, d, scale=0.05, seed=2), rnd(3 * d, seed=3)
    ang = torch.rand(L, 128, device=DEV) * 6.28
    cos, sin = ang.cos().bfloat16(), ang.sin().bfloat16()
    scale = 256 ** -0.5
    cu = (torch.arange(0, B + 1, dtype=torch.int32) * L).to(DEV)
    # single-GPU truth
    qkv = ops.gemm(x, w, b, epilogue=ops.EPI_QKV_ROPE, rope_cos=cos, rope_sin=sin, qk_cols=2 * d, rows_per_sample=L)
    att = ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, scale)
    # emulated ranks
    recv = [torch.zeros(B * L, 3 * dq, device=DEV, dtype=torch.bfloat16) for _ in range(P)]
    ao = [torch.zeros(B * Lq, d, device=DEV, dtype=torch.bfloat16) for _ in range(P)]
    recv_tab, ao_tab = _peer_table(recv), _peer_table(ao)
    xv = x.view(B, L, d)
    for r in range(P):
        xr = xv[:, r * Lq:(r + 1) * Lq].reshape(B * Lq, d).contiguous()
        ops.gemm_qkv_p2p(xr, w, b, cos[r * Lq:(r + 1) * Lq].contiguous(), sin[r * Lq:(r + 1) * Lq].contiguous(),
                         Lq, P, Hp, r, L, recv_tab)
    q4 = qkv.view(B * L, 3, P, dq)
    for r in range(P):
        assert torch.equal(recv[r].view(B * L, 3, dq), q4[:, :, r])
    for r in range(P):
        rr = recv[r]
        ops.attention_varlen_p2p(rr[:, :dq], rr[:, dq:2 * dq], rr[:, 2 * dq:], cu, cu, Hp, L, scale, ao_tab, P, Lq,
                                 r * Hp, d)
    got = torch.stack([a_.view(B, Lq, d) for a_ in ao], 1)          # [B, P, Lq, d]
    got = got.permute(0, 1, 2, 3).reshape(B, P * Lq, d).reshape(B * L, d)
    assert torch.equal(got, att)
    # the stream-K attention with the same fused return path (too few units here for a split: same bits)
    ao2 = [torch.zeros(B * Lq, d, device=DEV, dtype=torch.bfloat16) for _ in range(P)]
    ao2_tab = _peer_table(ao2)
    for r in range(P):
        rr = recv[r]
        ops.attention_streamk_p2p(rr[:, :dq], rr[:, dq:2 * dq], rr[:, 2 * dq:], cu, cu, Hp, L, L, scale, ao2_tab, P, Lq,
                                  r * Hp, d)
    for a_, b_ in zip(ao, ao2):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("sk_mode", [0, 1, 2], indirect=True)
def test_attention_streamk_peer_epilogue_with_split_units(sk_mode):
    """Stream-K attention with the staged peer-store epilogue at a size where units ARE split between clusters
    (2 x 4 heads x 5 query tiles = 40 units... 12 heads: 120 units > 74 clusters): the rows written through the peer table
    (all "peers" on this GPU) must equal the locally stored result of the same kernel."""
    from flite_b200 import ops
    B, H, L, P = 2, 12, 1280, 4
    d, Lq = H * 256, L // P
    qkv = rnd(B * L, 3 * d, scale=1.0, seed=11)
    cu = (torch.arange(0, B + 1, dtype=torch.int32) * L).to(DEV)
    scale = 256 ** -0.5
    local = ops.attention_streamk(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, L, scale)
    ao = [torch.zeros(B * Lq, d, device=DEV, dtype=torch.bfloat16) for _ in range(P)]
    ops.attention_streamk_p2p(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, L, scale, _peer_table(ao), P, Lq, 0, d)
    got = torch.stack([a_.view(B, Lq, d) for a_ in ao], 1).reshape(B * L, d)
    assert torch.equal(got, local)
    base = ops.attention_varlen(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], cu, cu, H, L, scale)
    assert rel(local, base) <= 2e-3


def test_p2p_flags_signal_then_wait_loopback():
    """flite_p2p_signal / flite_p2p_wait on one GPU: two emulated ranks publish into the same flag array, the wait
    returns once both slots carry the epoch (monotonic compare) and the watchdog stays clean."""
    import ctypes
    from flite_b200 import _lib
    lib = _lib.load()
    flags = torch.zeros(64, device=DEV, dtype=torch.int32)
    tab = (ctypes.c_void_p * 8)(flags.data_ptr(), *([None] * 7))
    s = torch.cuda.current_stream().cuda_stream
    for epoch in (1, 2, 3):
        for slot in (0, 1):
            _lib.check(lib.flite_p2p_signal(tab, 1, slot, epoch, s), "signal")
        _lib.check(lib.flite_p2p_wait(flags.data_ptr(), 2, epoch, s), "wait")
    _lib.watchdog_ok()
    assert flags[:2].tolist() == [3, 3]


@pytest.mark.parametrize("variant", [3, 4, 5, 6])
def test_attention_tma_store_epilogue_bit_equal(variant):
    """FLITE_TUNE_ATTN_TMA_OUT: whole 128-row output tiles leave through shared memory + TMA bulk stores (default); ragged
    last tiles, single-row and empty-key sequences keep the per-thread path.  Same bits as with the TMA path switched off,
    nothing written outside the rows of the call, and a column-offset / strided output view works."""
    from flite_b200 import _lib, ops
    H = 2
    q_lens, k_lens = [300, 1, 272, 512], [129, 0, 272, 40]
    cu_q = torch.tensor([0] + list(np.cumsum(q_lens)), dtype=torch.int32, device=DEV)
    cu_k = torch.tensor([0] + list(np.cumsum(k_lens)), dtype=torch.int32, device=DEV)
    nq = sum(q_lens)
    q, k, v = rnd(nq, H * 256, seed=1), rnd(sum(k_lens), H * 256, seed=2), rnd(sum(k_lens), H * 256, seed=3)
    lib = _lib.load()
    wide = torch.full((nq + 4, H * 256 + 64), 7.0, device=DEV, dtype=torch.bfloat16)     # strided view, 16-byte aligned
    ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=variant, out=wide[:nq, 64:])
    lib.flite_set_tuning(16, 1)
    try:
        rows = ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=variant)
    finally:
        lib.flite_set_tuning(16, 0)
    _lib.watchdog_ok()
    assert torch.equal(wide[:nq, 64:], rows)
    assert bool((wide[nq:] == 7.0).all()) and bool((wide[:, :64] == 7.0).all())


@pytest.mark.parametrize("variant", [3, 4, 5, 6])
def test_attention_staged_output_stores_bit_equal(variant):
    """FLITE_TUNE_ATTN_STAGED_STORES: whole-row output stores through the dead Q tile give the same bits as the
    one-row-per-thread stores; ragged query tails and an empty key sequence included."""
    from flite_b200 import _lib, ops
    H = 2
    q_lens, k_lens = [300, 1, 272], [129, 0, 272]
    cu_q = torch.tensor([0] + list(np.cumsum(q_lens)), dtype=torch.int32, device=DEV)
    cu_k = torch.tensor([0] + list(np.cumsum(k_lens)), dtype=torch.int32, device=DEV)
    q, k, v = rnd(sum(q_lens), H * 256, seed=1), rnd(sum(k_lens), H * 256, seed=2), rnd(sum(k_lens), H * 256, seed=3)
    plain = ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=variant)
    lib = _lib.load()
    lib.flite_set_tuning(7, 1)
    try:
        staged = torch.full((sum(q_lens) + 4, H * 256), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=variant, out=staged[:sum(q_lens)])
    finally:
        lib.flite_set_tuning(7, 0)
    assert torch.equal(staged[:sum(q_lens)], plain)
    assert bool((staged[sum(q_lens):] == 7.0).all())


# ----------------------------------------------------------------------------- sampler / pipeline tail
def _apg_torch(acc, u, c, g, dt, thr):
    """f_lite/pipeline.py:276-287,296 with torch ops in the tensors' dtype (the reference, run on this device)."""
    from oracle import sampler_oracle
    v = sampler_oracle.apg_combine(u, c, g, thr)
    return acc + dt * v


@pytest.mark.parametrize("shape", [(1, 16, 32, 32), (1, 16, 128, 128), (3, 16, 112, 168)])
@pytest.mark.parametrize("thr", [0.03, 5.0])
def test_apg_euler_matches_reference_ops(shape, thr):
    """flite_apg_euler vs the reference's torch op sequence.  The three global reductions are summed in a different
    order (fp64 block partials vs torch's fp32 tree), so a bf16 scalar may differ by one ulp: tolerance 1e-2 rel-L2 on
    the update, and the update must be far closer to the reference than plain CFG is (the kernel implements APG)."""
    from flite_b200 import ops
    u, c = rnd(*shape, seed=1), rnd(*shape, scale=1.0, seed=2) * 0.7 + rnd(*shape, seed=1) * 0.5
    acc0 = rnd(*shape, seed=3)
    g, dt = 6.0, 0.0625
    ref = _apg_torch(acc0, u, c, g, dt, thr)
    acc, lat = acc0.clone(), torch.empty_like(acc0)
    ops.apg_euler(acc, u, c, g, dt, thr, lat)
    assert torch.equal(acc, lat)
    d_ref = (ref - acc0).float()
    r = ((acc - acc0).float() - d_ref).norm() / d_ref.norm()
    plain = (dt * (u + g * (c - u))).float()
    assert r.item() <= 1e-2, r.item()
    assert (plain - d_ref).norm() / d_ref.norm() > 10 * max(r.item(), 1e-4)
    # fp32 accumulator (train.py sample_images semantics)
    acc32, lat32 = acc0.float(), torch.empty_like(acc0)
    ops.apg_euler(acc32, u, c, g, dt, thr, lat32)
    from oracle import sampler_oracle
    ref32 = acc0.float() + dt * sampler_oracle.apg_combine(u, c, g, thr).float()        # train.py:599 accumulation
    assert rel(acc32 - acc0.float(), ref32 - acc0.float()) <= 1e-2 and torch.equal(lat32, acc32.bfloat16())


def test_apg_euler_degenerate_inputs():
    """cond == uncond: dd = 0, orth = -coef*dy = 0, std = 0 -> threshold/std = inf -> scale 1; v = dy (no NaNs)."""
    from flite_b200 import ops
    c = rnd(1, 16, 32, 32, seed=4)
    acc, lat = torch.zeros_like(c), torch.empty_like(c)
    ops.apg_euler(acc, c.clone(), c, 6.0, 1.0, 0.03, lat)
    assert torch.isfinite(acc.float()).all() and torch.equal(acc, c)


@pytest.mark.parametrize("shape", [(1, 16, 32, 32), (2, 16, 128, 128)])
def test_latent_unscale_bit_exact(shape):
    from flite_b200 import ops
    lat = rnd(*shape, seed=5)
    assert torch.equal(ops.latent_unscale(lat, 0.3611, 0.1159), lat / 0.3611 + 0.1159)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(1, 3, 256, 256), (2, 3, 64, 96), (1, 1, 8, 8), (1, 4, 17, 5)])
def test_image_to_uint8_bit_exact(dtype, shape):
    """(x/2+0.5).clamp(0,1)*255 -> round -> uint8 (pipeline.py:324-326) + the NCHW->NHWC permute of pipeline.py:327."""
    from flite_b200 import ops
    x = (torch.randn(shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6)) * 0.8).to(dtype)
    x.view(-1)[:6] = torch.tensor([-1.0, 1.0, 0.0, -3.0, 3.0, 0.00196], device=DEV).to(dtype)[: min(6, x.numel())]
    ref = ((x / 2 + 0.5).clamp(0, 1) * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.image_to_uint8(x), ref)


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("rows,d,B", [(2 * 272, 512, 2), (2 * 4112, 3072, 2), (5, 256, 1), (37, 4096, 1), (3 * 5000, 1024, 3)])
def test_rmsnorm_kernel_variants_bit_equal(kernel, rows, d, B):
    """FLITE_TUNE_RMSNORM_KERNEL 1 (two-pass) / 2 (register-resident) / 3 (streaming, persistent warps with next-row
    prefetch) must produce identical bits (same arithmetic, different schedule)."""
    from flite_b200 import _lib, ops
    x = rnd(rows, d, seed=1)
    w = (1 + 0.1 * torch.randn(d, device=DEV)).bfloat16()
    mod = rnd(B, 9 * d, scale=0.5, seed=2)
    L = rows // B
    base = ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=L)
    lib = _lib.load()
    lib.flite_set_tuning(0, kernel)
    try:
        y = torch.full((rows + 3, d), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.rmsnorm_modulate(x, w, 1, mod[:, d:2 * d], mod[:, :d], rows_per_sample=L, out=y[:rows])
        y2 = ops.rmsnorm_modulate(x, w, 2)
    finally:
        lib.flite_set_tuning(0, 0)
    assert torch.equal(y[:rows], base) and bool((y[rows:] == 7.0).all())
    assert torch.equal(y2, ops.rmsnorm_modulate(x, w, 2))


def test_gemv_skinny_gemm_matches_tensor_core_variant():
    """M <= 8 (timestep MLP / adaLN modulation): the weight-streaming GEMV that FLITE_GEMM_AUTO picks and the tcgen05
    tile variant agree to fp32-accumulation-order noise (<= 1 bf16 ulp on a few outputs), bias + SiLU included."""
    from flite_b200 import ops
    for (M, N, K) in [(2, 12288, 3072), (2, 3072, 12288), (2, 6144, 3072), (5, 640, 512)]:
        a, w, b = rnd(M, K, scale=0.5, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(N, seed=3)
        for act in (0, 1):
            g = ops.gemm(a, w, b, act=act, variant=5)
            t = ops.gemm(a, w, b, act=act, variant=3)
            auto = ops.gemm(a, w, b, act=act)
            assert torch.equal(auto, g)                       # AUTO routes skinny shapes to the GEMV
            assert rel(g, t) <= 2e-3 and (g == t).float().mean().item() > 0.98


XRES_CASES = [
    # (q_lens, k_lens, H)
    ([272, 272], [256, 256], 2),          # full tiles of keys
    ([272, 300, 1], [17, 129, 200], 2),   # ragged keys on both sides of the 128-key half, ragged queries
    ([4112, 4112], [256, 77], 12),        # config C2 cross-attention shape
    ([500, 40], [0, 256], 1),             # an empty key sequence -> zeros
    ([256], [1], 3),                      # a single key
]


@pytest.mark.parametrize("q_lens,k_lens,H", XRES_CASES)
def test_attention_resident_kv_cross_attention(q_lens, k_lens, H):
    """FLITE_ATTN_XRES (persistent cross-attention with K/V resident in shared memory, <= 256 keys per sequence)
    against the fp32 restatement of flash_attn_varlen_func and against the general 2-CTA kernel."""
    from flite_b200 import ops
    from oracle.dit_oracle import flash_attn_varlen
    cu_q = torch.tensor([0] + list(np.cumsum(q_lens)), dtype=torch.int32, device=DEV)
    cu_k = torch.tensor([0] + list(np.cumsum(k_lens)), dtype=torch.int32, device=DEV)
    Tq, Tk = sum(q_lens), max(sum(k_lens), 1)
    q = rnd(Tq, H * 256, seed=1)
    kv = rnd(Tk, 2 * H * 256, seed=2)                      # K | V column halves of one buffer (strided views)
    k, v = kv[:, :H * 256], kv[:, H * 256:]
    out = torch.full((Tq + 2, H * 256), 7.0, device=DEV, dtype=torch.bfloat16)
    ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, out=out[:Tq], variant=9)
    ref = flash_attn_varlen(q.view(Tq, H, 256), k.reshape(Tk, H, 256), v.reshape(Tk, H, 256), cu_q, cu_k, 1 / 16)
    ref = ref.reshape(Tq, H * 256)
    for b, kl in enumerate(k_lens):                        # torch softmax over zero keys gives NaN; flash-attn gives 0
        if kl == 0:
            ref[cu_q[b]:cu_q[b + 1]] = 0
    general = ops.attention_varlen(q, k, v, cu_q, cu_k, H, max(q_lens), 1 / 16, variant=5)
    assert torch.isfinite(out[:Tq].float()).all() and bool((out[Tq:] == 7.0).all())
    assert rel(out[:Tq], ref) <= 4e-3 and rel(out[:Tq], general) <= 4e-3


def test_attention_resident_kv_rejects_long_key_sequences():
    """> 256 keys in a sequence: the kernel refuses (watchdog word, tag 96) instead of truncating the context."""
    from flite_b200 import _lib, ops
    H = 1
    cu_q = torch.tensor([0, 128], dtype=torch.int32, device=DEV)
    cu_k = torch.tensor([0, 300], dtype=torch.int32, device=DEV)
    q, k, v = rnd(128, 256, seed=1), rnd(300, 256, seed=2), rnd(300, 256, seed=3)
    ops.attention_varlen(q, k, v, cu_q, cu_k, H, 128, 1 / 16, variant=9)
    with pytest.raises(_lib.FliteError, match="precondition"):
        _lib.watchdog_ok()
    _lib.watchdog_ok()                                      # the word is cleared after it has been reported


@pytest.mark.parametrize("M", [8224, 356, 640, 300, 4100])
def test_gemm_narrow_last_m_tile_bit_equal(M):
    """FLITE_TUNE_GEMM_NARROW_M: a ragged last M-tile with <= 128 valid rows runs as tcgen05 M = 128 MMAs (64 rows per
    CTA, TMEM "layout B") instead of a padded 256-row tile.  Same bits as the padded path for the three epilogues that
    use it, rows beyond M untouched (M = 300 has a 44-row tail, 640 a 128-row one, 356 = 256 + 100, 8224 = C2, 4100 = C4
    per sequence-parallel rank)."""
    from flite_b200 import _lib, ops
    lib = _lib.load()
    d, inter, B = 512, 1024, 2
    g = torch.Generator(device=DEV).manual_seed(M)
    a = (torch.randn(M, d, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(d, d, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.randn(d, device=DEV, generator=g).bfloat16()
    gate = torch.randn(B, d, device=DEV, generator=g).bfloat16()
    x0 = torch.randn(M + 8, d, device=DEV, generator=g).bfloat16()
    wgu = ops.interleave_gate_up((torch.randn(inter, d, device=DEV, generator=g) * 0.05).bfloat16(),
                                 (torch.randn(inter, d, device=DEV, generator=g) * 0.05).bfloat16())
    a2 = (torch.randn(M, inter, device=DEV, generator=g) * 0.5).bfloat16()
    wd = (torch.randn(d, inter, device=DEV, generator=g) * 0.05).bfloat16()
    rps = (M + B - 1) // B

    def run_all():
        outs = []
        o = torch.full((M + 8, d), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a, w, bias, variant=_lib.GEMM_2CTA_N256, out=o[:M])
        outs.append(o)
        x = x0.clone()
        ops.gemm(a, w, bias, epilogue=ops.EPI_GATED_RES, resid=x[:M], gate=gate, rows_per_sample=rps,
                 variant=_lib.GEMM_2CTA_N256, out=x[:M])
        outs.append(x)
        h = torch.full((M + 8, inter), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.gemm(a, wgu, None, epilogue=ops.EPI_SWIGLU, variant=_lib.GEMM_2CTA_N256, out=h[:M])
        outs.append(h)
        y = x0.clone()
        ops.gemm(a2, wd, None, epilogue=ops.EPI_GATED_RES, resid=y[:M], gate=gate, rows_per_sample=rps,
                 variant=_lib.GEMM_2CTA_N256, out=y[:M])
        outs.append(y)
        return outs

    narrow = run_all()
    lib.flite_set_tuning(14, 1)
    try:
        padded = run_all()
    finally:
        lib.flite_set_tuning(14, 0)
    for i, (n_, p_) in enumerate(zip(narrow, padded)):
        assert torch.equal(n_, p_), f"output {i} differs between the narrow and the padded last M-tile"
    assert bool((narrow[0][M:] == 7.0).all()) and bool((narrow[2][M:] == 7.0).all())
    assert torch.equal(narrow[1][M:], x0[M:])
    assert rel(narrow[0][:M], F.linear(a, w, bias)) <= 1e-3


@pytest.fixture
def sk_mode(request):
    """schedule of the persistent attention kernel (FLITE_TUNE_ATTN_SK_MODE): 0 stream-K, 1 whole units round-robin, 2 hybrid"""
    from flite_b200 import _lib
    lib = _lib.load()
    old = lib.flite_get_tuning(15)
    lib.flite_set_tuning(15, request.param)
    yield request.param
    lib.flite_set_tuning(15, old)


@pytest.mark.parametrize("sk_mode", [0, 1, 2], indirect=True)
@pytest.mark.parametrize("B,H,Lq,Lk", [(2, 12, 4112, 4112), (4, 12, 600, 600), (2, 12, 1024, 300), (2, 2, 272, 272),
                                       (3, 4, 1000, 129), (1, 3, 16400, 16400), (8, 12, 4112, 4112)])
def test_attention_streamk_matches_general_kernel(B, H, Lq, Lk, sk_mode):
    """flite_attention_streamk (persistent wave of clusters, uniform lengths; unit read-out on the epilogue warps) vs
    flite_attention_varlen (one cluster per 256-query tile): units that one cluster computes alone give the same bits --
    in the round-robin schedule (mode 1) that is every unit; a unit split between two clusters (stream-K shares, modes 0
    and 2) is merged in fp32, so it may differ by a bf16 ulp.  Also vs the fp32 oracle, and a second launch must
    reproduce the first bit for bit (the merge flags are reset by their reader)."""
    from flite_b200 import ops
    from oracle import dit_oracle
    d = H * 256
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Lq)
    q = torch.randn(B * Lq, H, 256, device=DEV, generator=g)
    k = torch.randn(B * Lk, H, 256, device=DEV, generator=g)
    q = (q * torch.rsqrt(q.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().view(B * Lq, d)
    k = (k * torch.rsqrt(k.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().view(B * Lk, d)
    v = torch.randn(B * Lk, d, device=DEV, generator=g).bfloat16()
    cu_q = (torch.arange(B + 1, dtype=torch.int32) * Lq).to(DEV)
    cu_k = (torch.arange(B + 1, dtype=torch.int32) * Lk).to(DEV)
    scale = 256 ** -0.5
    base = ops.attention_varlen(q, k, v, cu_q, cu_k, H, Lq, scale)
    out = torch.full((B * Lq + 4, d), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.attention_streamk(q, k, v, cu_q, cu_k, H, Lq, Lk, scale, out=out[:B * Lq])
    again = ops.attention_streamk(q, k, v, cu_q, cu_k, H, Lq, Lk, scale)
    from flite_b200 import _lib
    _lib.watchdog_ok()
    assert bool((out[B * Lq:] == 7.0).all())
    assert torch.equal(out[:B * Lq], again)
    eq = (out[:B * Lq] == base).float().mean().item()
    r = rel(out[:B * Lq], base)
    print(f"streamk (mode {sk_mode}) vs general: bit-equal {eq:.4f} rel {r:.2e}")
    assert r <= 2e-3 and eq > 0.5
    if sk_mode == 1:
        assert torch.equal(out[:B * Lq], base)
    if B * Lq * Lk * H <= 2 * 12 * 4112 * 4112:
        ref = dit_oracle.flash_attn_varlen(q.view(-1, H, 256).float(), k.view(-1, H, 256).float(),
                                           v.view(-1, H, 256).float(), cu_q, cu_k, scale).reshape(-1, d)
        assert rel(out[:B * Lq], ref) <= 5e-3


def test_attention_streamk_rejects_ragged_lengths():
    from flite_b200 import _lib, ops
    H, d = 2, 512
    q = rnd(600, d); k = rnd(600, d); v = rnd(600, d)
    cu = torch.tensor([0, 280, 600], dtype=torch.int32, device=DEV)      # 280 + 320, not 2 x 300
    ops.attention_streamk(q, k, v, cu, cu, H, 300, 300, 256 ** -0.5)
    with pytest.raises(_lib.FliteError):
        _lib.watchdog_ok()


# ----------------------------------------------------------------------------- VAE decoder norm (next row: pipeline tail)
@pytest.mark.parametrize("N,C,H,W,silu", [(2, 128, 64, 64, True), (1, 256, 32, 48, True), (3, 512, 16, 16, True),
                                          (1, 512, 24, 24, False), (1, 128, 1024, 1024, True)])
def test_groupnorm_silu_nhwc_vs_torch(N, C, H, W, silu):
    """flite_groupnorm_silu_nhwc (the VAE decoder's GroupNorm -> SiLU pairs, diffusers ResnetBlock2D behind
    f_lite/pipeline.py:299-307) against torch's group_norm (+ silu) on the same channels-last bf16 tensor -- the kernel
    reproduces torch's bf16 rounding points (mean / rstd stored in bf16, folded affine in fp32, norm output rounded to
    bf16 before the activation), so all but a few near-tie elements are bit-equal -- and against the fp32 op on the same
    input (bounded by the bf16 rstd: 2^-9).  Deterministic: a second launch reproduces the first."""
    import torch.nn.functional as F
    from flite_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(N * 1000 + C)
    x = (torch.randn(N, C, H, W, device=DEV, generator=g) * 1.7 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, device=DEV, generator=g) * 0.5 + 1.0).bfloat16()
    b = (torch.randn(C, device=DEV, generator=g) * 0.2).bfloat16()
    got = ops.groupnorm_silu(x, w, b, 32, 1e-6, silu=silu)
    again = ops.groupnorm_silu(x, w, b, 32, 1e-6, silu=silu)
    assert got.is_contiguous(memory_format=torch.channels_last) and torch.equal(got, again)
    ref = F.group_norm(x, 32, w, b, 1e-6)
    ref32 = F.group_norm(x.float(), 32, w.float(), b.float(), 1e-6)
    if silu:
        ref, ref32 = F.silu(ref), F.silu(ref32)
    assert rel(got, ref32) <= 6e-3                       # bf16 rstd / mean + bf16 outputs, same as torch's own bf16 path
    assert rel(ref, ref32) <= 6e-3
    differ = (got != ref).float().mean().item()
    print(f"groupnorm vs torch bf16: {differ:.2e} of the elements differ, rel {rel(got, ref):.2e}")
    assert differ <= 1e-4 and rel(got, ref) <= 1e-4       # measured: bit-equal on every case


@pytest.mark.parametrize("with_residual", [False, True])
def test_bias_residual_add_nhwc_bit_equal_to_torch(with_residual):
    """flite_bias_residual_add_nhwc: the conv bias (torch: a broadcast add_ on the cuDNN output) and the ResnetBlock2D skip
    connection in one in-place channels-last pass, same two bf16 roundings as the torch ops."""
    from flite_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(5)
    cl = torch.channels_last
    y = torch.randn(2, 128, 40, 56, device=DEV, generator=g).bfloat16().contiguous(memory_format=cl)
    bias = torch.randn(128, device=DEV, generator=g).bfloat16()
    res = torch.randn(2, 128, 40, 56, device=DEV, generator=g).bfloat16().contiguous(memory_format=cl) if with_residual else None
    ref = y + bias.view(1, -1, 1, 1)
    if with_residual:
        ref = res + ref
    got = ops.bias_residual_add_(y.clone(memory_format=cl), bias, res)
    assert got.is_contiguous(memory_format=cl) and torch.equal(got, ref)


def test_upsample_nearest2x_nhwc_bit_equal_to_torch():
    import torch.nn.functional as F
    from flite_b200 import ops
    x = torch.randn(2, 256, 24, 40, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    got = ops.upsample_nearest2x(x)
    assert got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, F.interpolate(x, scale_factor=2.0, mode="nearest"))


def test_vae_decode_native_passes_vs_torch_passes():
    """flite_b200.vae decoder on a CUDA bf16 channels-last latent with the sm_100a norm / bias / skip passes against the
    same module with those passes run by torch (vae.USE_NATIVE_KERNELS = False).  The passes are bit-equal op by op (tests
    above) and so is everything up to the mid-block attention; from there the two runs hand cuBLAS / SDPA differently
    strided views of the same values (torch's GroupNorm returns NCHW, the kernel keeps channels-last), so library kernels
    with another summation order take over and the images agree to bf16 accuracy, not to the bit."""
    from flite_b200 import vae
    torch.manual_seed(0)
    m = vae.AutoencoderKL().to(DEV)
    for p_ in m.parameters():
        if p_.dim() > 1:
            p_.data.uniform_(-1, 1).mul_((3.0 / p_.shape[1:].numel()) ** 0.5)
        else:
            p_.data.normal_(0.0, 0.2).add_(1.0 if p_.shape[0] > 3 and "norm" in "" else 0.0)
    m = m.to(torch.bfloat16).to(memory_format=torch.channels_last).eval()
    z = torch.randn(2, 16, 32, 32, device=DEV).bfloat16()
    native = m.decode(z).sample
    vae.USE_NATIVE_KERNELS = False
    try:
        plain = m.decode(z).sample
    finally:
        vae.USE_NATIVE_KERNELS = True
    r = rel(native, plain)
    differ = (native != plain).float().mean().item()
    print(f"vae decode native vs torch passes: rel {r:.2e}, {differ:.2e} of the pixels differ")
    assert torch.isfinite(native.float()).all() and r <= 2e-2
