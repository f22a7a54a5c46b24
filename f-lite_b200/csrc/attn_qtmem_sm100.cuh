// 2-CTA varlen flash attention forward, head_dim 256, with BOTH A operands in tensor memory (sm_100a).
//
// Shared memory bandwidth (128 B/clk/SM) paces attn_fwd_cg2_kernel: Q (64 KB) is re-read from smem for every key
// tile.  Here Q is loaded once into TMEM (128 columns, bf16 pairs) and used as the A operand of S = Q K^T
// (tcgen05.mma [d], [a_tmem], b_desc); P is written by the softmax threads into the TMEM columns of S_j and used
// as the A operand of O += P V.  Shared memory then only holds a 6-stage ring of K/V half-tiles (64 keys per tile):
//   per CTA and 64-key tile: TMA writes 32 KB + MMA reads 32 KB  vs  1024 MMA cycles  =>  64 B/clk, half the port.
//
//   TMEM / CTA (512 cols): O [0,256) | Q [256,384) | S0 [384,448) | S1 [448,512)   (P_j aliases the first 32 cols of S_j)
//   SMEM / CTA: ring of 6 x (K half-tile 32 keys x 256 | V half-tile 64 keys x 128 cols) = 192 KB | barriers | exchange
// Cluster of 2 CTAs = two adjacent 128-query tiles of one (head, sequence); MMAs are cta_group::2 (M = 256).
// Roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2.. softmax warpgroups (kWG = 1|2).
#pragma once

#include "attn_sm100.cuh"

namespace flite {

constexpr int AQ_STAGES = 6;
constexpr int AQ_STAGE_BYTES = 32768;                 // K 16 KB + V 16 KB
constexpr int AQ_BAR = AQ_STAGES * AQ_STAGE_BYTES;    // 196608
constexpr int AQ_XCH = AQ_BAR + 256;
constexpr int AQ_SMEM = AQ_XCH + 2048 + 1024;
constexpr uint32_t AQ_TM_O = 0, AQ_TM_Q = 256, AQ_TM_S = 384;

template <int kWG>
__global__ void __launch_bounds__(64 + 128 * kWG, 1)
attn_fwd_qtmem_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const __nv_bfloat16* __restrict__ q_ptr, long long ldq, const AttnParams p) {
    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    const int q_beg = p.cu_q[b], q_len = p.cu_q[b + 1] - q_beg;
    if ((qt & ~1) * 128 >= q_len) return;  // uniform for the whole cluster
    const int k_beg = p.cu_k[b], k_len = p.cu_k[b + 1] - k_beg;
    const int n_tiles = (k_len + 63) / 64;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AQ_BAR);
    float* xch = reinterpret_cast<float*>(smem + AQ_XCH);
    uint64_t* full = bars;                       // [AQ_STAGES] leader: K+V of both CTAs landed
    uint64_t* empty = bars + AQ_STAGES;          // [AQ_STAGES] each CTA: stage consumed by PV_j
    uint64_t* s_full = bars + 2 * AQ_STAGES;     // [2] each CTA
    uint64_t* p_full = s_full + 2;               // [2] leader: P_j written by every softmax warp of both CTAs
    uint64_t* pv_done = p_full + 2;              //     each CTA
    uint64_t* q_ready = pv_done + 1;             //     leader: Q of both CTAs is in TMEM
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_ready + 1);

    const int warp_idx = threadIdx.x >> 5;
    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            for (int i = 0; i < AQ_STAGES; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s_full[i], 1);
                mbar_init(&p_full[i], 2 * 4 * kWG);
            }
            mbar_init(pv_done, 1);
            mbar_init(q_ready, 2 * 4 * kWG);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2>(tmem_ptr_smem, 512);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_o = tmem_base + AQ_TM_O;

    if (warp_idx == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (elect_one()) {
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % AQ_STAGES;
                const uint32_t ph = ((j / AQ_STAGES) & 1) ^ 1;
                const int krow = k_beg + j * 64;
                mbar_wait<true>(&empty[st], ph, 31);
                if (is_leader) mbar_arrive_expect_tx(&full[st], 2 * AQ_STAGE_BYTES);
                uint8_t* sk = smem + st * AQ_STAGE_BYTES;
                uint8_t* sv = sk + 16384;
#pragma unroll
                for (int c = 0; c < 4; ++c)   // this CTA's 32 key rows, 4 chunks of 64 head-dim columns
                    tma_load_2d_cg2(sk + c * 4096, &tmap_k, &full[st], 0, p.k_col0 + h * 256 + c * 64,
                                    krow + (int)cta_rank * 32);
#pragma unroll
                for (int c = 0; c < 2; ++c)   // all 64 key rows, this CTA's 128 head-dim columns
                    tma_load_2d_cg2(sv + c * 8192, &tmap_v, &full[st], 0,
                                    p.v_col0 + h * 256 + (int)cta_rank * 128 + c * 64, krow);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer (leader CTA) ================================
        if (is_leader && elect_one() && n_tiles > 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(256, 64, 0, 0);    // Q (TMEM) x K (K-major smem)
            constexpr uint32_t idesc_o = make_idesc_bf16(256, 256, 0, 1);   // P (TMEM) x V (MN-major smem)
            const uint32_t s0 = smem_u32(smem);
            auto issue_s = [&](int j) {
                const int st = j % AQ_STAGES;
                mbar_wait<true>(&full[st], (j / AQ_STAGES) & 1, 33);
                tc_fence_after();
                const uint32_t d = tmem_base + AQ_TM_S + (j & 1) * 64;
                const uint32_t sk = s0 + st * AQ_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    umma_ts<2>(d, tmem_base + AQ_TM_Q + k * 8,
                               make_smem_desc_sw128(sk + (k >> 2) * 4096 + (k & 3) * 32, 16, 1024), idesc_s,
                               k != 0 ? 1u : 0u);
                umma_commit_cg2(&s_full[j & 1], 0x3);
            };
            mbar_wait<true>(q_ready, 0, 34);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % AQ_STAGES;
                if (j + 1 < n_tiles) issue_s(j + 1);
                mbar_wait<true>(&p_full[j & 1], (j >> 1) & 1, 35);
                tc_fence_after();
                const uint32_t sv = s0 + st * AQ_STAGE_BYTES + 16384;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts<2>(tmem_o, tmem_base + AQ_TM_S + (j & 1) * 64 + k * 8,
                               make_smem_desc_sw128(sv + k * 2048, 8192, 1024), idesc_o, (j | k) != 0 ? 1u : 0u);
                umma_commit_cg2(&empty[st], 0x3);
                umma_commit_cg2(pv_done, 0x3);
            }
        }
        __syncwarp();
    } else {
        // ================================ softmax / correction / epilogue ================================
        constexpr int NC = 64 / kWG;                         // S columns (keys) per thread per tile
        constexpr int OC = 256 / kWG;                        // O columns per thread
        constexpr int QC = 128 / kWG;                        // packed Q columns per thread
        const int q = warp_idx & 3;                          // TMEM lane quarter of this warp
        const int half = (kWG == 2) ? ((warp_idx - 2) >> 2) : 0;
        const int lane = (int)lane_id();
        const int r = q * 32 + lane;                         // row inside the 128-query tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row_in_seq = qt * 128 + r;
        const bool row_ok = row_in_seq < q_len;

        // ---- Q row -> TMEM (packed bf16 pairs), once
        if (n_tiles > 0) {
            const uint4* qrow = reinterpret_cast<const uint4*>(q_ptr + (long long)(q_beg + row_in_seq) * ldq + p.q_col0 +
                                                               h * 256 + half * (2 * QC));
#pragma unroll
            for (int c = 0; c < QC / 32; ++c) {
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 v = row_ok ? __ldg(qrow + c * 8 + i) : make_uint4(0, 0, 0, 0);
                    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
                }
                tmem_st_x32(tmem_base + lane_off + AQ_TM_Q + half * QC + c * 32, w);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_cluster(q_ready, 0);
            __syncwarp();
        }

        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait<true>(&s_full[j & 1], (j >> 1) & 1, 37);
            tc_fence_after();
            const uint32_t ts = tmem_base + lane_off + AQ_TM_S + (j & 1) * 64;
            const int kv_valid = min(64, k_len - j * 64) - half * NC;   // valid columns of this thread's slice
            const bool full_tile = kv_valid >= NC;
            uint32_t s[NC];
            tmem_ld_x32(ts + half * NC, s);
            if constexpr (NC == 64) tmem_ld_x32(ts + 32, s + 32);
            tmem_ld_wait();
            float mx = -INFINITY;
            if (full_tile) {
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
                named_bar_sync(1 + q, 64);       // also: both warps of this row group have read S_j
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
            uint32_t pk[NC / 2];
            float rs0 = 0.f, rs1 = 0.f;
            const float neg_m = -m_used;
            if (full_tile) {
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
            if (j > 0 && need_any) {             // O rescale: P_{j-1} V_{j-1} must have completed
                mbar_wait<true>(pv_done, (j - 1) & 1, 38);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < OC / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_off + half * OC + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                    tmem_st_x32(tmem_o + lane_off + half * OC + c * 32, o);
                }
            }
            // P_j (packed bf16 pairs) over the first 32 columns of S_j
            if constexpr (NC == 64) tmem_st_x32(ts, pk);
            else tmem_st_x16(ts + half * (NC / 2), pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_cluster(&p_full[j & 1], 0);
            __syncwarp();
        }
        // ---- epilogue: O / l -> bf16 -> out[row, h*256 + c] ----
        __nv_bfloat16* orow = p.out + (long long)(q_beg + row_in_seq) * p.ldo + h * 256 + half * OC;
        if (n_tiles > 0) {
            if constexpr (kWG == 2) {
                float* slot = xch + (n_tiles & 1) * 256;
                slot[half * 128 + r] = l;
                named_bar_sync(1 + q, 64);
                l += slot[(half ^ 1) * 128 + r];
            }
            mbar_wait<true>(pv_done, (n_tiles - 1) & 1, 39);
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
            uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
            for (int i = 0; i < OC / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 1) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace flite
