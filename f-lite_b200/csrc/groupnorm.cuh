// GroupNorm (+ SiLU) for channels-last activations: the VAE decoder's norm -> activation pairs
// (diffusers ResnetBlock2D / Attention.group_norm / conv_norm_out behind f_lite/pipeline.py:299-307).  HBM-bound:
// x [N, HW, C] bf16 is read twice (statistics, apply) and written once; torch runs the same thing as a row-moments kernel
// plus unvectorised elementwise kernels on the channels-last tensor (85 % of the decode time at 1024^2, tools/vae_probe.py).
//   pass 1  groupnorm_stats_kernel : per (image, pixel slice) partial sum / sum of squares of every group, fixed order
//   pass 2  groupnorm_apply_kernel : mean / rstd from the partials (double, fixed order), y = bf16(x*a + b) [, silu -> bf16]
// Deterministic (no atomics).  Rounding points as in torch's bf16 path (aten group_norm_kernel.cu, what diffusers' VAE runs
// in bf16): mean and rstd are STORED in the activation dtype (bf16) before use, the affine is folded as a = rstd * gamma,
// b = beta - a * mean in fp32, y = bf16(a * x + b); SiLU is evaluated in fp32 on that bf16 value and rounded again.
#pragma once

#include "common.cuh"

namespace flite {

constexpr int GN_THREADS = 256;

// grid (splits, N).  Requires C % 8 == 0, (C / groups) % 4 == 0, C / 8 <= GN_THREADS, groups <= 64.
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, long long hw, int C, int groups, float2* __restrict__ partials) {
    extern __shared__ float gn_smem[];                       // [pixel lanes][C / 4 halves][2]
    const int n = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
    const int chunks = C >> 3;                               // 16-byte chunks per pixel
    const int lanes = GN_THREADS / chunks;                   // pixels processed per iteration
    const int c = threadIdx.x % chunks, pl = threadIdx.x / chunks;
    const long long p0 = hw * split / splits, p1 = hw * (split + 1) / splits;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;            // channels [8c, 8c+4) and [8c+4, 8c+8)
    if (pl < lanes) {
        const uint4* base = reinterpret_cast<const uint4*>(x + ((long long)n * hw) * C) + c;
        auto acc = [&](const uint4& v) {
            const float a0 = bf16_lo(v.x), a1 = bf16_hi(v.x), a2 = bf16_lo(v.y), a3 = bf16_hi(v.y);
            const float b0 = bf16_lo(v.z), b1 = bf16_hi(v.z), b2 = bf16_lo(v.w), b3 = bf16_hi(v.w);
            s0 += (a0 + a1) + (a2 + a3);
            q0 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            s1 += (b0 + b1) + (b2 + b3);
            q1 += (b0 * b0 + b1 * b1) + (b2 * b2 + b3 * b3);
        };
        long long p = p0 + pl;
        for (; p + 3LL * lanes < p1; p += 4LL * lanes) {      // four independent 16-byte loads in flight per thread
            const uint4 v0 = __ldcg(base + p * chunks), v1 = __ldcg(base + (p + lanes) * chunks);
            const uint4 v2 = __ldcg(base + (p + 2LL * lanes) * chunks), v3 = __ldcg(base + (p + 3LL * lanes) * chunks);
            acc(v0); acc(v1); acc(v2); acc(v3);
        }
        for (; p < p1; p += lanes) acc(__ldcg(base + p * chunks));
        float* slot = gn_smem + ((size_t)pl * (2 * chunks) + 2 * c) * 2;
        slot[0] = s0; slot[1] = q0; slot[2] = s1; slot[3] = q1;
    }
    __syncthreads();
    // one thread per group: its halves (4 channels each) over all pixel lanes, fixed order
    if (threadIdx.x < groups) {
        const int halves_per_group = (C / groups) >> 2;
        float s = 0.f, q = 0.f;
        for (int l = 0; l < lanes; ++l)
            for (int hh = 0; hh < halves_per_group; ++hh) {
                const float* slot = gn_smem + ((size_t)l * (2 * chunks) + threadIdx.x * halves_per_group + hh) * 2;
                s += slot[0]; q += slot[1];
            }
        partials[((long long)n * splits + split) * groups + threadIdx.x] = make_float2(s, q);
    }
}

// grid (blocks, N)
__global__ void __launch_bounds__(GN_THREADS)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                       const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, long long hw, int C,
                       int groups, float eps, int apply_silu, const float2* __restrict__ partials, int splits) {
    __shared__ float s_mean[64], s_rstd[64];
    const int n = blockIdx.y;
    if (threadIdx.x < groups) {
        double s = 0.0, q = 0.0;
        for (int i = 0; i < splits; ++i) {
            const float2 v = partials[((long long)n * splits + i) * groups + threadIdx.x];
            s += (double)v.x; q += (double)v.y;
        }
        const double cnt = (double)hw * (double)(C / groups);
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = bf16_round((float)mean);                       // torch keeps mean / rstd in the input dtype
        s_rstd[threadIdx.x] = bf16_round(rsqrtf((float)var + eps));
    }
    __syncthreads();
    const int chunks = C >> 3, cpg = C / groups;
    const long long total = hw * chunks;
    const uint4* xin = reinterpret_cast<const uint4*>(x + ((long long)n * hw) * C);
    uint4* yout = reinterpret_cast<uint4*>(y + ((long long)n * hw) * C);
    auto norm_chunk = [&](long long i, const uint4& v) {
        const int c = (int)(i % chunks);
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma) + c);
        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta) + c);
        const int g0 = (8 * c) / cpg, g1 = (8 * c + 4) / cpg;
        const float m0 = s_mean[g0], r0 = s_rstd[g0], m1 = s_mean[g1], r1 = s_rstd[g1];
        const uint32_t xw[4] = {v.x, v.y, v.z, v.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float m = j < 2 ? m0 : m1, r = j < 2 ? r0 : r1;
            const float ca = r * bf16_lo(gw[j]), cb = r * bf16_hi(gw[j]);          // a = rstd * gamma
            float a = bf16_round(fmaf(bf16_lo(xw[j]), ca, bf16_lo(bw[j]) - ca * m));   // y = a * x + (beta - a * mean)
            float b = bf16_round(fmaf(bf16_hi(xw[j]), cb, bf16_hi(bw[j]) - cb * m));
            if (apply_silu) { a = silu_f(a); b = silu_f(b); }
            o[j] = pack_bf16x2(a, b);
        }
        __stcs(yout + i, make_uint4(o[0], o[1], o[2], o[3]));       // written once, read next by a cuDNN kernel
    };
    const long long stride = (long long)gridDim.x * GN_THREADS;
    long long i = (long long)blockIdx.x * GN_THREADS + threadIdx.x;
    for (; i + 3 * stride < total; i += 4 * stride) {           // four independent 16-byte loads in flight per thread
        const uint4 v0 = __ldcs(xin + i), v1 = __ldcs(xin + i + stride), v2 = __ldcs(xin + i + 2 * stride),
                    v3 = __ldcs(xin + i + 3 * stride);
        norm_chunk(i, v0); norm_chunk(i + stride, v1); norm_chunk(i + 2 * stride, v2); norm_chunk(i + 3 * stride, v3);
    }
    for (; i < total; i += stride) norm_chunk(i, __ldcs(xin + i));
}

// y[N*HW, C] (channels-last, bf16) += bias[C]  [then y = residual + y], in place, 16 bytes per thread per step.
// torch runs a convolution's bias as a broadcast add_ on the cuDNN output (an unvectorised kernel on channels-last tensors:
// 19.5 ms of an 87 ms decode); rounding points kept: y = bf16(y + b), then bf16(residual + y) as the separate torch add.
__global__ void __launch_bounds__(256)
bias_residual_add_nhwc_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ bias,
                              const __nv_bfloat16* __restrict__ residual, long long rows, int C) {
    const int chunks = C >> 3;
    const long long total = rows * chunks;
    uint4* yv = reinterpret_cast<uint4*>(y);
    const uint4* rv = reinterpret_cast<const uint4*>(residual);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % chunks);
        const uint4 v = yv[i];
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(bias) + c);
        uint4 r = make_uint4(0, 0, 0, 0);
        if (residual != nullptr) r = __ldcg(rv + i);
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, bw[4] = {b.x, b.y, b.z, b.w}, rw[4] = {r.x, r.y, r.z, r.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = bf16_round(bf16_lo(vw[j]) + bf16_lo(bw[j])), a1 = bf16_round(bf16_hi(vw[j]) + bf16_hi(bw[j]));
            if (residual != nullptr) { a0 = bf16_lo(rw[j]) + a0; a1 = bf16_hi(rw[j]) + a1; }
            o[j] = pack_bf16x2(a0, a1);
        }
        yv[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Nearest-neighbour 2x upsampling of a channels-last activation: out[n, 2h+dy, 2w+dx, :] = in[n, h, w, :] (diffusers
// Upsample2D: F.interpolate(scale_factor=2, mode="nearest") in front of its convolution).  One 16-byte chunk of an input
// pixel is read once and written to its four output pixels.
__global__ void __launch_bounds__(256)
upsample_nearest2x_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H, int W, int C) {
    const int chunks = C >> 3;
    const long long total = (long long)N * H * W * chunks;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* yv = reinterpret_cast<uint4*>(y);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i % chunks);
        const long long pix = i / chunks;
        const int w = (int)(pix % W);
        const long long nh = pix / W;                     // n * H + h
        const uint4 v = __ldcg(xv + i);
        const long long o = ((nh * 2) * (2LL * W) + 2 * w) * chunks + c;       // output pixel (n, 2h, 2w)
        yv[o] = v;
        yv[o + chunks] = v;
        yv[o + 2LL * W * chunks] = v;
        yv[o + 2LL * W * chunks + chunks] = v;
    }
}

}  // namespace flite
