// HBM-bound kernels of the F Lite denoise step: RMSNorm + adaLN modulate, standalone RoPE + QK-norm,
// patch embedding, timestep embedding, unpatchify, varlen context packing, and the sampler's fused
// CFG-combine + Euler update.  All vectorised 16-byte accesses, one warp per token row where a row
// reduction is needed.  bf16 rounding points follow the reference op sequence.
#pragma once

#include "common.cuh"

namespace flite {

FLITE_DEVICE void unpack8(const uint4& v, float* f) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
FLITE_DEVICE uint4 pack8(const float* f) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                      pack_bf16x2(f[6], f[7]));
}
FLITE_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// One 16-byte chunk (8 bf16) of  y = RMSNorm_w(x) [* (1 + scale) + shift]  with the reference's bf16 rounding points.
// The bf16 x bf16 products / sums are done with packed mul.rn / add.rn.bf16x2 (never contracted into an FMA:
// exact product, one rounding) which is bit-identical to "compute in fp32, round to bf16" of the torch bf16 elementwise kernels.
FLITE_DEVICE uint4 norm_mod_chunk(const uint4& xv, float rstd, const uint4& wv, int weight_mode, bool mod,
                                  const uint4& scv, const uint4& shv) {
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
    const uint32_t sc[4] = {scv.x, scv.y, scv.z, scv.w}, sh[4] = {shv.x, shv.y, shv.z, shv.w};
    uint32_t out[4];
    const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float f0 = bf16_lo(xw[j]) * rstd, f1 = bf16_hi(xw[j]) * rstd;
        __nv_bfloat162 n2;
        if (weight_mode == 2) {
            n2 = __floats2bfloat162_rn(f0 * bf16_lo(ww[j]), f1 * bf16_hi(ww[j]));       // bf16(x*rstd*w), fp32 math
        } else {
            n2 = __floats2bfloat162_rn(f0, f1);                                          // bf16(x*rstd)
            if (weight_mode == 1) n2 = __hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&ww[j]));
        }
        if (mod) {
            const __nv_bfloat162 op = __hadd2_rn(one2, *reinterpret_cast<const __nv_bfloat162*>(&sc[j]));
            n2 = __hadd2_rn(__hmul2_rn(n2, op), *reinterpret_cast<const __nv_bfloat162*>(&sh[j]));
        }
        out[j] = *reinterpret_cast<uint32_t*>(&n2);
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}
FLITE_DEVICE float ssq_chunk(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
        s = fmaf(a, a, s);
        s = fmaf(b, b, s);
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// y = RMSNorm_w(x) [* (1 + scale[s]) + shift[s]]           one warp per row, d % 256 == 0
//   weight_mode 0: no weight            (f_lite/model.py:101-108, QK-norm style)
//   weight_mode 1: Liger "llama" cast   bf16(x*rstd) * w in bf16      (model.py:238,283,292,299,437)
//   weight_mode 2: reference RMSNorm    bf16(x*rstd*w) in fp32        (model.py:104-106, final_norm)
//   modulate: n*(1+scale)+shift with a bf16 rounding after each op    (model.py:284,293,300,580)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rmsnorm_modulate_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                        long long ldy, const __nv_bfloat16* __restrict__ w, int weight_mode,
                        const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift,
                        long long ld_mod, int rows_per_sample, int rows, int d, float eps) {
    pdl_launch_dependents();
    pdl_wait();                 // x is the previous kernel's output (PDL launch)
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const int nchunk = d >> 3;  // 16-byte chunks per row
    float ssq = 0.f;
    for (int c = lane; c < nchunk; c += 32) ssq += ssq_chunk(__ldcg(xr + c));
    ssq = warp_sum(ssq);
    const float rstd = rsqrtf(ssq / (float)d + eps);
    const bool mod = scale != nullptr;
    const long long s = (long long)(row / rows_per_sample) * ld_mod;
    uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = lane; c < nchunk; c += 32) {
        const uint4 wv = weight_mode != 0 ? __ldg(reinterpret_cast<const uint4*>(w) + c) : z;
        const uint4 scv = mod ? __ldg(reinterpret_cast<const uint4*>(scale + s) + c) : z;
        const uint4 shv = mod ? __ldg(reinterpret_cast<const uint4*>(shift + s) + c) : z;
        yr[c] = norm_mod_chunk(__ldcg(xr + c), rstd, wv, weight_mode, mod, scv, shv);
    }
}

// Register-resident variant for d <= 8 * 32 * MAXC: the whole row is loaded once (MAXC independent 16-byte
// loads in flight per lane), reduced, normalised and written -- one HBM read + one write per element.
template <int MAXC>
__global__ void __launch_bounds__(256, (MAXC <= 12) ? 2 : 1)
rmsnorm_modulate_reg_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                            long long ldy, const __nv_bfloat16* __restrict__ w, int weight_mode,
                            const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift,
                            long long ld_mod, int rows_per_sample, int rows, int d, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const int nchunk = d >> 3;
    uint4 v[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        v[i] = (c < nchunk) ? __ldcg(xr + c) : make_uint4(0, 0, 0, 0);
    }
    float ssq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) ssq += ssq_chunk(v[i]);
    ssq = warp_sum(ssq);
    const float rstd = rsqrtf(ssq / (float)d + eps);
    const bool mod = scale != nullptr;
    const long long s = (long long)(row / rows_per_sample) * ld_mod;
    uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunk) {
            const uint4 wv = weight_mode != 0 ? __ldg(reinterpret_cast<const uint4*>(w) + c) : z;
            const uint4 scv = mod ? __ldg(reinterpret_cast<const uint4*>(scale + s) + c) : z;
            const uint4 shv = mod ? __ldg(reinterpret_cast<const uint4*>(shift + s) + c) : z;
            yr[c] = norm_mod_chunk(v[i], rstd, wv, weight_mode, mod, scv, shv);
        }
    }
}

// Streaming variant (FLITE_TUNE_RMSNORM_KERNEL = 3): persistent warps, one row per warp per iteration, the row is read
// ONCE into registers and the NEXT row's loads are issued before the current row is reduced and written, so loads,
// math and stores of neighbouring rows overlap inside every warp (the one-shot kernels above run load / math / store
// phases in lock step across the whole grid).  The host sizes the grid so that every warp gets the same row count.
template <int MAXC>
__global__ void __launch_bounds__(256, (MAXC <= 4) ? 2 : 1)
rmsnorm_modulate_stream_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y,
                               long long ldy, const __nv_bfloat16* __restrict__ w, int weight_mode,
                               const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift,
                               long long ld_mod, int rows_per_sample, int rows, int d, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 8;
    const int nchunk = d >> 3;
    const bool mod = scale != nullptr;
    const uint4 z = make_uint4(0, 0, 0, 0);
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint4 cur[MAXC], nxt[MAXC];
    if (row < rows) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            cur[i] = (c < nchunk) ? __ldcg(xr + c) : z;
        }
    }
    for (; row < rows; row += nwarps) {
        const int nrow = row + nwarps;
        if (nrow < rows) {
            const uint4* xn = reinterpret_cast<const uint4*>(x + (long long)nrow * ldx);
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const int c = lane + 32 * i;
                nxt[i] = (c < nchunk) ? __ldcg(xn + c) : z;
            }
        }
        float ssq = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) ssq += ssq_chunk(cur[i]);
        ssq = warp_sum(ssq);
        const float rstd = rsqrtf(ssq / (float)d + eps);
        const long long s = (long long)(row / rows_per_sample) * ld_mod;
        uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * ldy);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = lane + 32 * i;
            if (c < nchunk) {
                const uint4 wv = weight_mode != 0 ? __ldg(reinterpret_cast<const uint4*>(w) + c) : z;
                const uint4 scv = mod ? __ldg(reinterpret_cast<const uint4*>(scale + s) + c) : z;
                const uint4 shv = mod ? __ldg(reinterpret_cast<const uint4*>(shift + s) + c) : z;
                yr[c] = norm_mod_chunk(cur[i], rstd, wv, weight_mode, mod, scv, shv);
            }
        }
#pragma unroll
        for (int i = 0; i < MAXC; ++i) cur[i] = nxt[i];
    }
}

// ------------------------------------------------------------------------------------------
// Standalone RoPE + QK-RMSNorm, in place on packed projections [T, ld] (head_dim 256).
// One warp per (token, head-slot); slots [0, n_rope_norm) are rotated + normalised, e.g. q and k heads
// of a qkv buffer.  Kept as the unfused alternative to the GEMM's EPI_QKV_ROPE epilogue.
// f_lite/model.py:166-180,403-414,92-108
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rope_qknorm_kernel(__nv_bfloat16* __restrict__ buf, long long ld, int rows, int n_slots,
                   const __nv_bfloat16* __restrict__ cos_t, const __nv_bfloat16* __restrict__ sin_t,
                   int rows_per_sample, float eps) {
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= (long long)rows * n_slots) return;
    const int row = (int)(w / n_slots), slot = (int)(w % n_slots);
    const int lane = threadIdx.x & 31;
    __nv_bfloat16* p = buf + (long long)row * ld + slot * 256;
    // lane handles elements [4*lane, 4*lane+4) of the first half and the matching second-half elements
    uint2 a = *reinterpret_cast<const uint2*>(p + 4 * lane);
    uint2 b = *reinterpret_cast<const uint2*>(p + 128 + 4 * lane);
    float x1[4] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y)};
    float x2[4] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y)};
    if (cos_t != nullptr) {
        const int pos = row % rows_per_sample;
        const uint2 cv = *reinterpret_cast<const uint2*>(cos_t + (long long)pos * 128 + 4 * lane);
        const uint2 sv = *reinterpret_cast<const uint2*>(sin_t + (long long)pos * 128 + 4 * lane);
        const float cs[4] = {bf16_lo(cv.x), bf16_hi(cv.x), bf16_lo(cv.y), bf16_hi(cv.y)};
        const float sn[4] = {bf16_lo(sv.x), bf16_hi(sv.x), bf16_lo(sv.y), bf16_hi(sv.y)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y1 = bf16_round(x1[j] * cs[j] + x2[j] * sn[j]);
            const float y2 = bf16_round(x1[j] * (-sn[j]) + x2[j] * cs[j]);
            x1[j] = y1; x2[j] = y2;
        }
    }
    float ssq = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) ssq += x1[j] * x1[j] + x2[j] * x2[j];
    ssq = warp_sum(ssq);
    const float rstd = rsqrtf(ssq * (1.0f / 256.0f) + eps);
    a = make_uint2(pack_bf16x2(x1[0] * rstd, x1[1] * rstd), pack_bf16x2(x1[2] * rstd, x1[3] * rstd));
    b = make_uint2(pack_bf16x2(x2[0] * rstd, x2[1] * rstd), pack_bf16x2(x2[2] * rstd, x2[3] * rstd));
    *reinterpret_cast<uint2*>(p + 4 * lane) = a;
    *reinterpret_cast<uint2*>(p + 128 + 4 * lane) = b;
}

// ------------------------------------------------------------------------------------------
// Patch embedding (Conv2d k = s = patch, f_lite/model.py:318-328) + register-token rows (model.py:535).
// tokens[b*L + 16 + (hy*wp + wx), n] = bf16(sum_k patch[k] * W[n, k] + bias[n]),  k order (c, p1, p2)
// tokens[b*L + r, n]                 = register_tokens[r, n]           r < n_reg
// Block = 256 threads, PE_TOK tokens; each thread owns output columns tid, tid+256, ...
// ------------------------------------------------------------------------------------------
constexpr int PE_TOK = 32;
template <int KDIM>
__global__ void __launch_bounds__(256)
patch_embed_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ reg_tokens,
                   __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int P, int d, int n_reg,
                   int tok_offset, int tok_count) {
    // rows of `out` are (sample, local token); local token t is sequence position tok_offset + t (sequence-parallel
    // ranks embed only their own slice; tok_offset = 0, tok_count = L otherwise)
    __shared__ float patch[PE_TOK][KDIM + 1];
    const int wp = W / P, L = tok_count;
    const int tok0 = blockIdx.x * PE_TOK;           // index over B * tok_count rows
    for (int i = threadIdx.x; i < PE_TOK * KDIM; i += 256) {
        const int t = i / KDIM, k = i % KDIM;
        const int r = tok0 + t;
        float v = 0.f;
        if (r < B * L) {
            const int b = r / L, l = tok_offset + r % L;
            if (l >= n_reg) {
                const int pi = l - n_reg, hy = pi / wp, wx = pi % wp;
                const int c = k / (P * P), p1 = (k / P) % P, p2 = k % P;
                v = __bfloat162float(x[(((long long)b * C + c) * H + hy * P + p1) * W + wx * P + p2]);
            }
        }
        patch[t][k] = v;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < d; n += 256) {
        float wr[KDIM];
        const uint4* wp4 = reinterpret_cast<const uint4*>(w + (long long)n * KDIM);
#pragma unroll
        for (int j = 0; j < KDIM / 8; ++j) unpack8(__ldg(wp4 + j), wr + 8 * j);
        const float bv = __bfloat162float(bias[n]);
#pragma unroll 4
        for (int t = 0; t < PE_TOK; ++t) {
            const int r = tok0 + t;
            if (r >= B * L) break;
            const int l = tok_offset + r % L;
            if (l < n_reg) {
                out[(long long)r * d + n] = reg_tokens[(long long)l * d + n];
            } else {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < KDIM; ++k) acc += patch[t][k] * wr[k];
                out[(long long)r * d + n] = __float2bfloat16_rn(acc + bv);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Sinusoidal timestep embedding, f_lite/model.py:20-28,551:  arg = float(t*1000 [in t's dtype]) * freqs
// emb = bf16(cat[cos(arg), sin(arg)]).  freqs[half] is computed by the host with the reference formula.
// ------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, int t_is_bf16, const float* __restrict__ freqs,
                                      __nv_bfloat16* __restrict__ out, int B, int d) {
    const int half = d / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, j = i % half;
    float tv = t[b];
    // t_is_bf16: 0 = fp32 timesteps, 1 = bf16 timesteps (t*1000 rounded to bf16), 2 = caller already scaled
    if (t_is_bf16 == 1) tv = bf16_round(bf16_round(tv) * 1000.0f);
    else if (t_is_bf16 == 0) tv = tv * 1000.0f;
    const float arg = tv * freqs[j];
    out[(long long)b * d + j] = __float2bfloat16_rn(cosf(arg));
    out[(long long)b * d + half + j] = __float2bfloat16_rn(sinf(arg));
}

// ------------------------------------------------------------------------------------------
// Unpatchify, f_lite/model.py:577,583-590: "b (h w) (p1 p2 c) -> b c (h p1) (w p2)", register rows dropped.
// ------------------------------------------------------------------------------------------
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ tok, long long ldt,
                                  __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int P, int n_reg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * C * H * W;
    if (i >= total) return;
    const int wx_full = (int)(i % W), hy_full = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)(i / ((long long)W * H * C));
    const int hp = H / P, wp = W / P, L = n_reg + hp * wp;
    const int hy = hy_full / P, p1 = hy_full % P, wx = wx_full / P, p2 = wx_full % P;
    const long long row = (long long)b * L + n_reg + hy * wp + wx;
    out[i] = tok[row * ldt + (p1 * P + p2) * C + c];
}

// ------------------------------------------------------------------------------------------
// Varlen context packing without a host sync, f_lite/model.py:31-64 (mask.sum -> cu_seqlens; nonzero ->
// index_select).  One block per sequence computes the packed position of every valid token.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_scan_kernel(const float* __restrict__ mask, int Lc, int* __restrict__ pos, int* __restrict__ seqlens) {
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < Lc; base += 256) {
        const int j = base + threadIdx.x;
        const int valid = (j < Lc && mask[(long long)b * Lc + j] != 0.f) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int within = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int off = carry;
        for (int k = 0; k < wid; ++k) off += warp_tot[k];
        if (j < Lc) pos[(long long)b * Lc + j] = valid ? off + within : -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < 8; ++k) tot += warp_tot[k];
            carry += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) seqlens[b] = carry;
}
__global__ void cu_seqlens_kernel(const int* __restrict__ seqlens, int B, int* __restrict__ cu) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        cu[0] = 0;
        for (int b = 0; b < B; ++b) { acc += seqlens[b]; cu[b + 1] = acc; }
    }
}
__global__ void __launch_bounds__(128)
pack_rows_kernel(const __nv_bfloat16* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
                 const int* __restrict__ pos, const int* __restrict__ cu, int B, int Lc, int d) {
    const int r = blockIdx.x;  // source row in [0, B*Lc)
    const int pp = pos[r];
    if (pp < 0) return;
    const int b = r / Lc;
    const uint4* s = reinterpret_cast<const uint4*>(src + (long long)r * lds);
    uint4* t = reinterpret_cast<uint4*>(dst + (long long)(cu[b] + pp) * ldd);
    for (int c = threadIdx.x; c < (d >> 3); c += 128) t[c] = __ldg(s + c);
}

// ------------------------------------------------------------------------------------------
// [n0, n1, n2] -> [n1, n0, n2] copy of bf16 rows (n2 % 8 == 0): layout transform around the Ulysses all-to-alls.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
permute_021_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n0, int n1, int n2v) {
    const long long total = (long long)n0 * n1 * n2v;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % n2v);
        const long long t = i / n2v;
        const int b = (int)(t % n0), a = (int)(t / n0);      // dst index (a in n1, b in n0, c)
        dst[i] = __ldg(src + ((long long)b * n1 + a) * n2v + c);
    }
}

// ------------------------------------------------------------------------------------------
// Cross-GPU completion flags for the fused (peer-memory) Ulysses exchange.  signal: after all prior work of this
// stream (kernel boundary) publish `value` into slot `my_slot` of every peer's flag array; wait: spin until every
// slot of the local flag array has reached `value` (monotonic epochs, never reset).  One warp, bounded spin.
// ------------------------------------------------------------------------------------------
struct PeerFlagPtrs { unsigned int* p[8]; };
__global__ void p2p_signal_kernel(PeerFlagPtrs peers, int n, int my_slot, unsigned int value) {
    const int g = threadIdx.x;
    if (g < n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[g] + my_slot), "r"(value) : "memory");
    }
}
// Cross-RANK wait: the peers are paced by their hosts, so the policy is NCCL's (minutes, configurable through
// FLITE_TUNE_P2P_TIMEOUT_S), not the 2 s of the intra-kernel barrier watchdog; host-level skew is absorbed before the first
// peer store by one NCCL all-reduce per forward (model.py), so this spin normally only sees kernel-level skew.  On a
// time-out the sticky abort word is set and poison_on_abort_kernel turns the forward's output into NaNs (fail closed).
__global__ void p2p_wait_kernel(const unsigned int* flags, int n, unsigned int value, unsigned long long timeout_ns) {
    const int s = threadIdx.x;
    if (s < n) {
        const uint64_t t0 = globaltimer_ns();
        unsigned int v;
        while (true) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + s) : "memory");
            if ((int)(v - value) >= 0) break;
            if (*(volatile unsigned int*)&g_flite_abort != 0) break;
            if (globaltimer_ns() - t0 > timeout_ns) {
                atomicCAS(&g_flite_abort, 0u, (77u << 16) | (unsigned)s | 0x80000000u);
                break;
            }
        }
    }
    __threadfence_system();
}

// Fail closed: if any barrier / peer wait of this process has timed out (sticky g_flite_abort), overwrite `buf` with
// bf16 NaNs so that a caller who never checks flite_watchdog_status cannot consume a half-exchanged result.
__global__ void poison_on_abort_kernel(uint4* buf, long long n16) {
    if (*(volatile unsigned int*)&g_flite_abort == 0) return;
    const uint4 nan4 = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        buf[i] = nan4;
}

// ------------------------------------------------------------------------------------------
// Skinny GEMM (M <= 8 rows) for the timestep path: time_embed MLP, adaLN_modulation, final_modulation
// (f_lite/model.py:448-454,472,553-556,578) -- 358 MB of weights per step at the 10B architecture against a few KB of
// activations: pure weight streaming, HBM-bound.  A 128-row tensor-core tile would leave most SMs idle (N / 128 CTAs,
// each pulling its whole K extent alone), so: every W row is read exactly once with coalesced 16-byte streaming loads
// spread over all SMs, the M activation rows come from L1, fp32 accumulation, warp + block reduction,
// out[m, n] = act(bf16(acc + bias[n])) with the same rounding points as the EPI_STORE epilogue.
// Algorithmic bytes per launch: N*K*2 (weights) + M*(K + N)*2.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int GEMV_ROWS(int maxm) { return 4; }   // 8 rows for M <= 2 measured slower (53 -> 67 us at 27648 x 3072)
template <int MAXM>
__global__ void __launch_bounds__(256)
gemv_small_m_kernel(const __nv_bfloat16* __restrict__ A, long long lda, const __nv_bfloat16* __restrict__ Wt, long long ldw,
                    __nv_bfloat16* __restrict__ C, long long ldc, const __nv_bfloat16* __restrict__ bias, int act, int M,
                    int N, int K) {
    // A block owns GEMV_ROWS consecutive output columns (= rows of W) at a time; its 256 threads split the K extent, so
    // every thread has GEMV_ROWS independent 16-byte weight loads in flight per step and the activation chunk it
    // fetched from L1 is reused for all of them.
    constexpr int ROWS = GEMV_ROWS(MAXM);
    __shared__ float part[8][ROWS][MAXM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunk = K >> 3;                       // 16-byte chunks per row
    for (int n0 = blockIdx.x * ROWS; n0 < N; n0 += gridDim.x * ROWS) {
        float acc[ROWS][MAXM];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int m = 0; m < MAXM; ++m) acc[r][m] = 0.f;
        for (int c = threadIdx.x; c < nchunk; c += 256) {
            uint4 wv[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                wv[r] = (n0 + r < N) ? ld_nc_stream(reinterpret_cast<const uint4*>(Wt + (long long)(n0 + r) * ldw) + c)
                                     : make_uint4(0, 0, 0, 0);
            float af[MAXM][8];
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                if (m < M) unpack8(__ldg(reinterpret_cast<const uint4*>(A + (long long)m * lda) + c), af[m]);
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float wf[8];
                unpack8(wv[r], wf);
#pragma unroll
                for (int m = 0; m < MAXM; ++m) {
                    if (m < M) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[r][m] = fmaf(af[m][j], wf[j], acc[r][m]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int m = 0; m < MAXM; ++m) {
                const float v = (m < M) ? warp_sum(acc[r][m]) : 0.f;
                if (lane == 0) part[warp][r][m] = v;
            }
        __syncthreads();
        if (threadIdx.x < ROWS * MAXM) {
            const int r = threadIdx.x / MAXM, m = threadIdx.x % MAXM, n = n0 + r;
            if (m < M && n < N) {
                float v = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) v += part[w8][r][m];
                v = bf16_round(v + (bias != nullptr ? __bfloat162float(bias[n]) : 0.f));
                if (act == 1) v = silu_f(v);
                C[(long long)m * ldc + n] = __float2bfloat16_rn(v);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Patchify as a tensor-core GEMM (f_lite/model.py:318-328,535): this kernel only builds the im2col rows
//   A[b * n_img + i, k] = x[b, c, hy*P + p1, wx*P + p2],  k = (c, p1, p2)   (i-th IMAGE token of this rank's slice)
// and copies the learned register-token rows into the token matrix; the projection itself (K = C*P*P) then runs on
// gemm_bf16_kernel (+bias) straight into the image rows of the token matrix, one launch per sample.
// One block per local token row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
patch_gather_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ reg_tokens,
                    __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ out, int B, int C, int H, int W, int P,
                    int d, int n_reg, int tok_offset, int tok_count, int n_reg_local) {
    const int r = blockIdx.x;                       // local row: (sample, local token)
    const int b = r / tok_count, t = r - b * tok_count, l = tok_offset + t;
    if (l < n_reg) {
        const uint4* src = reinterpret_cast<const uint4*>(reg_tokens + (long long)l * d);
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)r * d);
        for (int i = threadIdx.x; i < d / 8; i += blockDim.x) dst[i] = __ldg(src + i);
        return;
    }
    const int kdim = C * P * P, wp = W / P;
    const int pi = l - n_reg, hy = pi / wp, wx = pi - hy * wp;
    const long long arow = (long long)b * (tok_count - n_reg_local) + (t - n_reg_local);
    for (int k = threadIdx.x; k < kdim; k += blockDim.x) {
        const int c = k / (P * P), p1 = (k / P) % P, p2 = k % P;
        A[arow * kdim + k] = x[(((long long)b * C + c) * H + hy * P + p1) * W + wx * P + p2];
    }
}

// ------------------------------------------------------------------------------------------
// Sampler: fused CFG combine + Euler update (f_lite/pipeline.py:290,296-297; f_lite/train.py:596,599).
//   v   = bf16(u + bf16(g * bf16(c - u)))                     (tensor ops in the model dtype)
//   acc = acc + dt * v      bf16 accumulate (pipeline)  |  fp32 accumulate (train.py sample_images)
//   lat = bf16(acc)                                            (next model input)
// One pass, 8 elements per thread, 16-byte accesses.  acc_is_fp32 selects the accumulator dtype.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cfg_euler_kernel(void* __restrict__ acc, int acc_is_fp32, const __nv_bfloat16* __restrict__ v_uncond,
                 const __nv_bfloat16* __restrict__ v_cond, float g, float dt, int do_cfg,
                 __nv_bfloat16* __restrict__ lat_out, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
         i += (long long)gridDim.x * blockDim.x) {
        float c[8], u[8], v[8], a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_cond) + i), c);
        if (do_cfg) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(v_uncond) + i), u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = bf16_round(u[j] + bf16_round(g * bf16_round(c[j] - u[j])));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = c[j];
        }
        if (acc_is_fp32) {
            float4* ap = reinterpret_cast<float4*>(acc) + 2 * i;
            float4 a0 = ap[0], a1 = ap[1];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] + dt * v[j];
            ap[0] = make_float4(a[0], a[1], a[2], a[3]);
            ap[1] = make_float4(a[4], a[5], a[6], a[7]);
        } else {
            uint4* ap = reinterpret_cast<uint4*>(acc) + i;
            unpack8(*ap, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + bf16_round(dt * v[j]));
            *ap = pack8(a);
        }
        reinterpret_cast<uint4*>(lat_out)[i] = pack8(a);
    }
}


// ------------------------------------------------------------------------------------------
// Sampler, Augmented Parallel Guidance (f_lite/pipeline.py:276-287) fused with the Euler update.  The reference
// evaluates three GLOBAL reductions over the whole batch tensor with torch ops in the model dtype:
//   dy = c; dd = bf16(c - u); coef = bf16(bf16(sum(bf16(dy*dd))) / bf16(sum(bf16(dy*dy)))); par = bf16(coef*dy);
//   orth = bf16(dd - par); std = bf16(unbiased std(orth)); scale = min(1, bf16(thr/std)); orth = bf16(orth*scale);
//   v = bf16(dy + bf16((g-1)*orth));  acc += dt*v;  lat = bf16(acc)
// Three stream-ordered launches share per-block fp64 partial sums in a caller-provided workspace (no host sync,
// no atomics: every block re-reduces the partials in the same order, so the scalars are deterministic).
// ------------------------------------------------------------------------------------------
constexpr int APG_BLOCKS = 128;
constexpr int APG_THREADS = 256;

FLITE_DEVICE double apg_block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = (threadIdx.x < APG_THREADS / 32) ? sh[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}
// sum of the APG_BLOCKS partials at ws[0..APG_BLOCKS), same order in every block
FLITE_DEVICE double apg_total(const double* ws, double* sh) {
    return apg_block_sum(threadIdx.x < APG_BLOCKS ? ws[threadIdx.x] : 0.0, sh);
}

__global__ void __launch_bounds__(APG_THREADS)
apg_dot_kernel(const __nv_bfloat16* __restrict__ v_uncond, const __nv_bfloat16* __restrict__ v_cond, long long n8,
               double* __restrict__ ws) {
    __shared__ double sh[APG_THREADS / 32];
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float c[8], u[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_cond) + i), c);
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_uncond) + i), u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dd = bf16_round(c[j] - u[j]);
            s1 += (double)bf16_round(c[j] * dd);
            s2 += (double)bf16_round(c[j] * c[j]);
        }
    }
    const double t1 = apg_block_sum(s1, sh);
    const double t2 = apg_block_sum(s2, sh);
    if (threadIdx.x == 0) { ws[blockIdx.x] = t1; ws[APG_BLOCKS + blockIdx.x] = t2; }
}

FLITE_DEVICE float apg_coef(const double* ws, double* sh) {
    const float s1 = bf16_round((float)apg_total(ws, sh));
    const float s2 = bf16_round((float)apg_total(ws + APG_BLOCKS, sh));
    return bf16_round(s1 / s2);
}

__global__ void __launch_bounds__(APG_THREADS)
apg_orth_stats_kernel(const __nv_bfloat16* __restrict__ v_uncond, const __nv_bfloat16* __restrict__ v_cond,
                      long long n8, double* __restrict__ ws) {
    __shared__ double sh[APG_THREADS / 32];
    const float coef = apg_coef(ws, sh);
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float c[8], u[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_cond) + i), c);
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_uncond) + i), u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float orth = bf16_round(bf16_round(c[j] - u[j]) - bf16_round(coef * c[j]));
            s1 += (double)orth;
            s2 += (double)orth * (double)orth;
        }
    }
    const double t1 = apg_block_sum(s1, sh);
    const double t2 = apg_block_sum(s2, sh);
    if (threadIdx.x == 0) { ws[2 * APG_BLOCKS + blockIdx.x] = t1; ws[3 * APG_BLOCKS + blockIdx.x] = t2; }
}

__global__ void __launch_bounds__(APG_THREADS)
apg_euler_kernel(void* __restrict__ acc, int acc_is_fp32, const __nv_bfloat16* __restrict__ v_uncond,
                 const __nv_bfloat16* __restrict__ v_cond, float gm1, float dt, float threshold,
                 __nv_bfloat16* __restrict__ lat_out, long long n8, const double* __restrict__ ws) {
    __shared__ double sh[APG_THREADS / 32];
    const float coef = apg_coef(ws, sh);
    const double n = (double)n8 * 8.0;
    const double sum = apg_total(ws + 2 * APG_BLOCKS, sh), sq = apg_total(ws + 3 * APG_BLOCKS, sh);
    const double mean = sum / n;
    double var = (sq - n * mean * mean) / (n - 1.0);          // torch.std default: unbiased
    if (var < 0.0) var = 0.0;
    const float stdv = bf16_round((float)sqrt(var));
    const float ratio = bf16_round(threshold / stdv);          // python float / 0-dim bf16 tensor -> bf16
    const float scale = ratio < 1.0f ? ratio : 1.0f;           // min(1, ratio); NaN (std == 0) -> 1 like python's min
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float c[8], u[8], a[8], v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_cond) + i), c);
        unpack8(__ldg(reinterpret_cast<const uint4*>(v_uncond) + i), u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float orth = bf16_round(bf16_round(c[j] - u[j]) - bf16_round(coef * c[j]));
            orth = bf16_round(orth * scale);
            v[j] = bf16_round(c[j] + bf16_round(gm1 * orth));
        }
        if (acc_is_fp32) {
            float4* ap = reinterpret_cast<float4*>(acc) + 2 * i;
            float4 a0 = ap[0], a1 = ap[1];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] + dt * v[j];
            ap[0] = make_float4(a[0], a[1], a[2], a[3]);
            ap[1] = make_float4(a[4], a[5], a[6], a[7]);
        } else {
            uint4* ap = reinterpret_cast<uint4*>(acc) + i;
            unpack8(*ap, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = bf16_round(a[j] + bf16_round(dt * v[j]));
            *ap = pack8(a);
        }
        reinterpret_cast<uint4*>(lat_out)[i] = pack8(a);
    }
}

// ------------------------------------------------------------------------------------------
// Pipeline tail (f_lite/pipeline.py:299-327).
//   latent_unscale : lat / scaling_factor + shift_factor   (pipeline.py:304), bf16 rounding points; like torch's CUDA
//                    kernel for division by a host scalar the quotient is lat * (1 / scaling_factor) in fp32
//   image_to_uint8 : (x / 2 + 0.5).clamp(0,1) * 255 -> round (half to even) -> uint8, NCHW -> NHWC
//                    (pipeline.py:324-327: the reference keeps NCHW and permutes per image on the host)
// in_is_fp32 selects the decoder dtype: bf16 reproduces the rounding after every torch op, fp32 computes in fp32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
latent_unscale_kernel(const __nv_bfloat16* __restrict__ lat, __nv_bfloat16* __restrict__ out, float inv_scaling,
                      float shift, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(lat) + i), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = bf16_round(bf16_round(a[j] * inv_scaling) + shift);
        reinterpret_cast<uint4*>(out)[i] = pack8(a);
    }
}

__global__ void __launch_bounds__(256)
image_to_uint8_kernel(const void* __restrict__ img, int in_is_fp32, uint8_t* __restrict__ out, int C, long long hw,
                      long long total_px) {
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < total_px;
         px += (long long)gridDim.x * blockDim.x) {
        const long long b = px / hw, r = px - b * hw;
        for (int c = 0; c < C; ++c) {
            const long long src = (b * C + c) * hw + r;
            float x;
            if (in_is_fp32) {
                x = __ldg(reinterpret_cast<const float*>(img) + src);
                x = fminf(fmaxf(x / 2.0f + 0.5f, 0.0f), 1.0f);
                x = fminf(fmaxf(rintf(x * 255.0f), 0.0f), 255.0f);
            } else {
                x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(img)[src]);
                x = bf16_round(bf16_round(x / 2.0f) + 0.5f);
                x = fminf(fmaxf(x, 0.0f), 1.0f);
                x = fminf(fmaxf(rintf(bf16_round(x * 255.0f)), 0.0f), 255.0f);
            }
            out[px * C + c] = (uint8_t)x;
        }
    }
}

}  // namespace flite
