// Varlen non-causal flash attention forward for head_dim 256 on sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces flash_attn_interface.flash_attn_varlen_func at f_lite/model.py:203-210 (self-attention over
// the packed image tokens and cross-attention over the packed text tokens).  Q/K/V are read in place from
// the projection buffers ([tokens, ld] with head h at columns col0 + 256*h), so the reference's
// "l (k h d) -> k h l d" / "h l d -> l h d" rearranges (model.py:163,181-183) cost nothing.
//
// One CTA = one (128-query tile, head, sequence).  Warp roles:
//   warp 0     : TMA producer  (Q once, then K_j / V_j tiles of 128 keys, 128B-swizzled)
//   warp 1     : tcgen05.mma issuer:  S_j = Q K_j^T  (128x128, fp32 in TMEM, double-buffered)
//                                      O  += P_j V_j  (128x256, fp32 in TMEM; V consumed MN-major)
//   warps 2..  : softmax warpgroups (kWG = 1 or 2), one query row per thread; with kWG = 2 the two
//                warpgroups split the 128 key columns of every S tile (and the 256 O columns) in halves and
//                exchange the partial row maxima through smem.  Running max / sum in registers, lazy O
//                rescale (only when the row max grows by > 2^8), P_j written as bf16 into swizzled smem,
//                final O / l -> bf16 -> global.
// TMEM: S0 [0,128) | S1 [128,256) | O [256,512).   SMEM: Q 64K | K 64K | V 64K | P 32K | barriers | exchange.
#pragma once

#include "common.cuh"

namespace flite {

struct AttnParams {
    const int* cu_q;        // [B+1] cumulative query lengths (rows into the Q tensor)
    const int* cu_k;        // [B+1] cumulative key lengths   (rows into the K/V tensors)
    __nv_bfloat16* out;     // out[row, h*256 + c]
    long long ldo;
    int q_col0, k_col0, v_col0;
    float scale_log2;       // softmax_scale * log2(e)
    int debug;              // profiling experiments only (0 in production): bit0 skip softmax math, bit1 skip K/V reloads
    // Fused Ulysses return path (attn_fwd_cg2_kernel): when out_peer[0] != nullptr query row l of this rank's heads is
    // stored into the token owner's buffer over NVLink peer memory: out_peer[l / sp_lq][(b*sp_lq + l % sp_lq), sp_head0 + h]
    __nv_bfloat16* out_peer[8];
    int sp_lq;
    int sp_head0;
    int stage_out;          // attn_fwd_cg2_kernel: 1 = store whole output rows via a shared-memory transpose
    // attn_fwd_cg2_kernel: 1 = a CTA whose 128 query rows are all valid writes its O tile as bf16 into the (dead) Q region
    // of shared memory, 128B-swizzled, and one thread sends it out with four TMA box stores: the per-thread row stores of
    // the plain epilogue are 32 different cache lines per instruction and cost ~8k cycles per unit (ncu: 6 % of the launch)
    int tma_out;
};

constexpr int ATT_SQ = 0, ATT_SK = 65536, ATT_SV = 131072, ATT_SP = 196608, ATT_BAR = 229376;
constexpr int ATT_XCH = ATT_BAR + 128;          // float [2 parity][2 half][128 rows]
constexpr int ATT_SMEM_USED = ATT_XCH + 2048;
constexpr int ATT_SMEM = 232448;                // 227 KB: everything the SM gives one CTA

FLITE_DEVICE float fast_exp2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int kWG>
__global__ void __launch_bounds__(64 + 128 * kWG, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    const int q_beg = p.cu_q[b], q_len = p.cu_q[b + 1] - q_beg;
    if (qt * 128 >= q_len) return;  // uniform for the whole CTA, before any barrier / TMEM allocation
    const int k_beg = p.cu_k[b], k_len = p.cu_k[b + 1] - k_beg;
    const int n_tiles = (k_len + 127) / 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    if ((smem - smem_raw) + ATT_SMEM_USED > ATT_SMEM) {   // dynamic smem base not aligned as expected
        if (threadIdx.x == 0) atomicCAS(&g_flite_abort, 0u, (99u << 16) | 0x80000000u);
        return;
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR);
    float* xch = reinterpret_cast<float*>(smem + ATT_XCH);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 2;
    uint64_t* v_full = bars + 3;
    uint64_t* v_empty = bars + 4;
    uint64_t* s_full = bars + 5;   // [2]
    uint64_t* p_full = bars + 7;
    uint64_t* pv_done = bars + 8;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp_idx = threadIdx.x >> 5;
    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            mbar_init(q_full, 1);
            mbar_init(k_full, 1);
            mbar_init(k_empty, 1);
            mbar_init(v_full, 1);
            mbar_init(v_empty, 1);
            mbar_init(&s_full[0], 1);
            mbar_init(&s_full[1], 1);
            mbar_init(p_full, 128 * kWG);
            mbar_init(pv_done, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(tmem_ptr_smem, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_o = tmem_base + 256;

    const int q_row0 = q_beg + qt * 128;

    if (warp_idx == 0) {
        // ================================ TMA producer ================================
        if (elect_one() && n_tiles > 0) {
            mbar_arrive_expect_tx(q_full, 65536);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tma_load_2d(smem + ATT_SQ + c * 16384, &tmap_q, q_full, p.q_col0 + h * 256 + c * 64, q_row0);
            for (int j = 0; j < n_tiles; ++j) {
                const int krow = k_beg + j * 128;
                mbar_wait(k_empty, (j & 1) ^ 1, 11);
                mbar_arrive_expect_tx(k_full, 65536);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tma_load_2d(smem + ATT_SK + c * 16384, &tmap_k, k_full, p.k_col0 + h * 256 + c * 64, krow);
                mbar_wait(v_empty, (j & 1) ^ 1, 12);
                mbar_arrive_expect_tx(v_full, 65536);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tma_load_2d(smem + ATT_SV + c * 16384, &tmap_v, v_full, p.v_col0 + h * 256 + c * 64, krow);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer ================================
        if (elect_one() && n_tiles > 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
            constexpr uint32_t idesc_o = make_idesc_bf16(128, 256, 0, 1);   // P (K-major) x V (MN-major)
            const uint32_t sq = smem_u32(smem + ATT_SQ), sk = smem_u32(smem + ATT_SK);
            const uint32_t sv = smem_u32(smem + ATT_SV), sp = smem_u32(smem + ATT_SP);
            auto issue_s = [&](int j) {
                mbar_wait(k_full, j & 1, 13);
                tc_fence_after();
                const uint32_t d = tmem_base + (j & 1) * 128;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
                    umma_ss<1>(d, make_smem_desc_sw128(sq + off, 16, 1024), make_smem_desc_sw128(sk + off, 16, 1024),
                               idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit(k_empty);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0, 14);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_s(j + 1);
                mbar_wait(p_full, j & 1, 15);
                mbar_wait(v_full, j & 1, 16);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t da = make_smem_desc_sw128(sp + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sv + k * 2048, 16384, 1024);
                    umma_ss<1>(tmem_o, da, db, idesc_o, (j | k) != 0 ? 1u : 0u);
                }
                umma_commit(v_empty);
                umma_commit(pv_done);
            }
        }
        __syncwarp();
    } else {
        // ================================ softmax / correction / epilogue ================================
        constexpr int NC = 128 / kWG;                        // S columns (keys) per thread per tile
        constexpr int OC = 256 / kWG;                        // O columns per thread
        const int q = warp_idx & 3;                          // TMEM lane quarter of this warp
        const int half = (kWG == 2) ? ((warp_idx - 2) >> 2) : 0;
        const int lane = (int)lane_id();
        const int r = q * 32 + lane;                         // row inside the 128-query tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1, 17);
            tc_fence_after();
            const uint32_t ts = tmem_base + lane_off + (j & 1) * 128 + half * NC;
            const int kv_valid = min(128, k_len - j * 128) - half * NC;   // valid columns of this thread's slice
            const bool full = kv_valid >= NC;
            uint32_t s[NC];
#pragma unroll
            for (int c = 0; c < NC / 32; ++c) tmem_ld_x32(ts + c * 32, s + c * 32);
            tmem_ld_wait();
            float mx = -INFINITY;
            if (full) {
#pragma unroll
                for (int i = 0; i < NC; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
            } else {
#pragma unroll
                for (int i = 0; i < NC; ++i)
                    if (i < kv_valid) mx = fmaxf(mx, __uint_as_float(s[i]));
            }
            if constexpr (kWG == 2) {
                float* slot = xch + (j & 1) * 256;
                slot[half * 128 + r] = mx;
                named_bar_sync(1 + q, 64);
                mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
            }
            const float m_new = fmaxf(m_used, mx * p.scale_log2);
            const bool need = (j > 0) && (m_new - m_used > 8.0f);
            const bool need_any = __any_sync(0xffffffffu, need);
            float corr = 1.0f;
            if (j == 0) {
                m_used = m_new;
            } else if (need_any) {
                corr = fast_exp2(m_used - m_new);
                m_used = m_new;
            }
            // p = 2^(s*scale - m), packed to bf16 in place
            uint32_t pk[NC / 2];
            float rs0 = 0.f, rs1 = 0.f;
            const float neg_m = -m_used;
            if (full) {
#pragma unroll
                for (int i = 0; i < NC; i += 2) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, neg_m));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, neg_m));
                    rs0 += p0; rs1 += p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NC; i += 2) {
                    const float p0 = (i < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, neg_m)) : 0.f;
                    const float p1 = (i + 1 < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, neg_m)) : 0.f;
                    rs0 += p0; rs1 += p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
            }
            l = l * corr + (rs0 + rs1);
            // P_{j-1} V_{j-1} must be complete before P (smem) or O (TMEM) are touched
            if (j > 0) {
                mbar_wait(pv_done, (j - 1) & 1, 18);
                tc_fence_after();
                if (need_any) {
#pragma unroll 1
                    for (int c = 0; c < OC / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld_x32(tmem_o + lane_off + half * OC + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st_x32(tmem_o + lane_off + half * OC + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            // write P (bf16, K-major, 128B swizzle): 64-key chunks of [128 rows x 128 B]
            uint8_t* sp_row = smem + ATT_SP + r * 128 + ((kWG == 2) ? half * 16384 : 0);
#pragma unroll
            for (int u = 0; u < NC / 8; ++u) {
                const int chunk = u >> 3, unit = u & 7;
                uint4 v = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                *reinterpret_cast<uint4*>(sp_row + chunk * 16384 + ((unit ^ (r & 7)) << 4)) = v;
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> out[row, h*256 + c] ----
        const int row_in_seq = qt * 128 + r;
        const bool row_ok = row_in_seq < q_len;
        __nv_bfloat16* orow = p.out + (long long)(q_beg + row_in_seq) * p.ldo + h * 256 + half * OC;
        if (n_tiles > 0) {
            if constexpr (kWG == 2) {
                // the parity slot of the (non-existent) next tile is free: tile n-2's reads all precede barrier n-1
                float* slot = xch + (n_tiles & 1) * 256;
                slot[half * 128 + r] = l;
                named_bar_sync(1 + q, 64);
                l += slot[(half ^ 1) * 128 + r];
            }
            mbar_wait(pv_done, (n_tiles - 1) & 1, 19);
            tc_fence_after();
            const float inv_l = 1.0f / l;
#pragma unroll 1
            for (int c = 0; c < OC / 32; ++c) {
                uint32_t o[32];
                tmem_ld_x32(tmem_o + lane_off + half * OC + c * 32, o);
                tmem_ld_wait();
                if (row_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        dst[i] = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l));
                }
            }
        } else if (row_ok) {
            // empty key sequence: flash-attn returns zeros
            uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
            for (int i = 0; i < OC / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace flite
