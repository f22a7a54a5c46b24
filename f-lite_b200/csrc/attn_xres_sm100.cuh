// Cross-attention with resident K/V: varlen non-causal attention for SHORT key sequences (<= 256 keys per sequence,
// the packed text context of f_lite/model.py:188-210), head_dim 256, sm_100a.
//
// attn_fwd_cg2_kernel spends most of a cross-attention launch in per-CTA fixed cost: with two 128-key tiles per query
// tile the prologue (barrier init, TMEM allocation, cluster syncs) and the exposed Q / K / V load latency dominate
// (0.108 ms per launch at 10B / 1024^2 for ~0.03 ms of tensor work).  This kernel is persistent instead:
//   * a cluster of two CTAs (cta_group::2) walks a contiguous range of units (sequence b, head h, 256-query tile pair);
//   * the K / V of (b, h) -- at most 256 keys -- stay resident in the pair's shared memory while the units of that
//     (b, h) are processed: K as 128 key rows per CTA, V as all 256 keys x 128 head-dim columns per CTA;
//   * the whole key extent is ONE tile: S = Q K^T is a single 256 x 256 MMA chain, the softmax is a single pass (exact
//     row max, no running rescale), O = P V a single 256 x 256 chain with P read from TMEM (packed bf16 over S);
//   * Q of unit n+1 is loaded as soon as the S MMAs of unit n have retired, under the softmax / PV of unit n.
// Roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warps 2..9 two softmax warpgroups (each row is
// split between two threads: keys [0,128) / [128,256) and O columns [0,128) / [128,256)).
//   SMEM / CTA: Q 64K | K 64K | V 64K | barriers | exchange          TMEM / CTA: S [0,256) (P aliases [0,128)) | O [256,512)
#pragma once

#include "attn_sm100.cuh"

namespace flite {

constexpr int XR_SQ = 0, XR_SK = 65536, XR_SV = 131072, XR_BAR = 196608;
constexpr int XR_XCH = XR_BAR + 128;             // float [2 parity][max | sum][2 halves][128 rows]
constexpr int XR_USED = XR_XCH + 2 * 2 * 2 * 128 * 4;
constexpr int XR_SMEM = XR_USED + 1024;
constexpr int XR_THREADS = 64 + 256;

struct XresParams {
    const int* cu_q;
    const int* cu_k;
    __nv_bfloat16* out;
    long long ldo;
    int q_col0, k_col0, v_col0;
    float scale_log2;
    int B, H, q_pairs;       // units = B * H * q_pairs, unit = ((b * H) + h) * q_pairs + qp
};

__global__ void __launch_bounds__(XR_THREADS, 1)
attn_xres_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const XresParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XR_BAR);
    float* xch = reinterpret_cast<float*>(smem + XR_XCH);
    uint64_t* q_full = bars + 0;    // leader: Q of the unit landed (both CTAs' halves)
    uint64_t* kv_full = bars + 1;   // leader: K and V of the (b, h) group landed
    uint64_t* s_full = bars + 2;    // each CTA: S MMAs retired (S readable, Q reusable)
    uint64_t* p_full = bars + 3;    // leader: P written by every softmax warp of both CTAs
    uint64_t* pv_done = bars + 4;   // each CTA: PV MMAs retired (O readable, K/V reusable)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp_idx = threadIdx.x >> 5;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    pdl_launch_dependents();
    pdl_wait();
    // precondition (same answer in every thread of every CTA, before any barrier): the whole key extent of a sequence
    // is ONE resident 256-key tile.  Fail loudly through the watchdog word instead of truncating the context.
    {
        bool bad = (smem - smem_raw) + XR_USED > XR_SMEM;
        for (int b = 0; b < p.B; ++b) bad |= (p.cu_k[b + 1] - p.cu_k[b]) > 256;
        if (bad) {
            if (threadIdx.x == 0) atomicCAS(&g_flite_abort, 0u, (96u << 16) | 0x80000000u);
            return;
        }
    }

    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            mbar_init(q_full, 1);
            mbar_init(kv_full, 1);
            mbar_init(s_full, 1);
            mbar_init(p_full, 2 * 8);
            mbar_init(pv_done, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2>(tmem_ptr_smem, 512);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_o = tmem_base + 256;

    // contiguous unit range of this cluster; consecutive units share (b, h) so K/V are reloaded only at group changes
    const int n_units = p.B * p.H * p.q_pairs;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int u0 = (int)((long long)cluster_id * n_units / n_clusters);
    const int u1 = (int)((long long)(cluster_id + 1) * n_units / n_clusters);
    // a unit is processed iff its query tile pair starts inside the sequence and the sequence has keys
    // (the cumulative lengths are re-read from global memory only when the sequence changes: their latency would
    // otherwise sit on every unit's critical path in all three roles)
    int dec_b = -1, dec_q_beg = 0, dec_q_len = 0, dec_k_beg = 0, dec_k_len = 0;
    auto decode = [&](int u, int& b, int& h, int& qp, int& q_beg, int& q_len, int& k_beg, int& k_len) {
        qp = u % p.q_pairs;
        const int g = u / p.q_pairs;
        h = g % p.H;
        b = g / p.H;
        if (b != dec_b) {
            dec_b = b;
            dec_q_beg = p.cu_q[b]; dec_q_len = p.cu_q[b + 1] - dec_q_beg;
            dec_k_beg = p.cu_k[b]; dec_k_len = p.cu_k[b + 1] - dec_k_beg;
        }
        q_beg = dec_q_beg; q_len = dec_q_len; k_beg = dec_k_beg; k_len = dec_k_len;
    };

    if (warp_idx == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (elect_one()) {
            int n = 0, last_group = -1;
            for (int u = u0; u < u1; ++u) {
                int b, h, qp, q_beg, q_len, k_beg, k_len;
                decode(u, b, h, qp, q_beg, q_len, k_beg, k_len);
                if (qp * 256 >= q_len || k_len <= 0) continue;
                const int group = u / p.q_pairs;
                // Order matters for the parity waits: once S of unit n-1 has retired, PV of unit n-2 has too (the tensor
                // pipe is in order), so pv_done is at phase n-1 or n and waiting for parity (n-1)&1 is unambiguous.
                if (n > 0) mbar_wait<true>(s_full, (n - 1) & 1, 42);           // previous unit's S MMAs have read Q
                if (group != last_group) {
                    if (n > 0) mbar_wait<true>(pv_done, (n - 1) & 1, 41);      // previous group's last PV has read K/V
                    if (is_leader) mbar_arrive_expect_tx(kv_full, 2 * 131072);
#pragma unroll
                    for (int c = 0; c < 4; ++c)   // this CTA's 128 key rows, 4 chunks of 64 head-dim columns
                        tma_load_2d_cg2(smem + XR_SK + c * 16384, &tmap_k, kv_full, 0, p.k_col0 + h * 256 + c * 64,
                                        k_beg + (int)cta_rank * 128);
#pragma unroll
                    for (int c = 0; c < 2; ++c)   // all 256 keys (two 128-row boxes), this CTA's 128 head-dim columns
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb)
                            tma_load_2d_cg2(smem + XR_SV + c * 32768 + hb * 16384, &tmap_v, kv_full, 0,
                                            p.v_col0 + h * 256 + (int)cta_rank * 128 + c * 64, k_beg + hb * 128);
                    last_group = group;
                }
                if (is_leader) mbar_arrive_expect_tx(q_full, 2 * 65536);
                const int q_row0 = q_beg + (2 * qp + (int)cta_rank) * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tma_load_2d_cg2(smem + XR_SQ + c * 16384, &tmap_q, q_full, 0, p.q_col0 + h * 256 + c * 64, q_row0);
                ++n;
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer (leader CTA) ================================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(256, 256, 0, 0);   // Q (K-major) x K (K-major), 256 keys
            constexpr uint32_t idesc_o = make_idesc_bf16(256, 256, 0, 1);   // P (TMEM)    x V (MN-major)
            const uint32_t sq = smem_u32(smem + XR_SQ), sk = smem_u32(smem + XR_SK), sv = smem_u32(smem + XR_SV);
            int n = 0, g = 0, last_group = -1;
            for (int u = u0; u < u1; ++u) {
                int b, h, qp, q_beg, q_len, k_beg, k_len;
                decode(u, b, h, qp, q_beg, q_len, k_beg, k_len);
                if (qp * 256 >= q_len || k_len <= 0) continue;
                const int group = u / p.q_pairs;
                if (group != last_group) {
                    mbar_wait<true>(kv_full, g & 1, 43);
                    ++g;
                    last_group = group;
                }
                mbar_wait<true>(q_full, n & 1, 44);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 16; ++k) {      // 16 head-dim columns per step inside the 64-column swizzle chunks
                    const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
                    umma_ss<2>(tmem_base, make_smem_desc_sw128(sq + off, 16, 1024), make_smem_desc_sw128(sk + off, 16, 1024),
                               idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit_cg2(s_full, 0x3);
                mbar_wait<true>(p_full, n & 1, 45);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 16; ++k) {      // 16 keys per step: 8 packed P columns, 16 rows of V (2048 B per chunk)
                    const uint64_t db = make_smem_desc_sw128(sv + k * 2048, 32768, 1024);
                    umma_ts<2>(tmem_o, tmem_base + k * 8, db, idesc_o, k != 0 ? 1u : 0u);
                }
                umma_commit_cg2(pv_done, 0x3);
                ++n;
            }
        }
        __syncwarp();
    } else {
        // ================================ softmax / epilogue (two warpgroups) ================================
        constexpr int NC = 128;                               // keys per thread (half of the 256-key tile)
        constexpr int OC = 128;                               // O columns per thread
        const int q = warp_idx & 3;                           // TMEM lane quarter of this warp
        const int half = (warp_idx - 2) >> 2;
        const int lane = (int)lane_id();
        const int r = q * 32 + lane;                          // row inside this CTA's 128-query tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        int n = 0;
        for (int u = u0; u < u1; ++u) {
            int b, h, qp, q_beg, q_len, k_beg, k_len;
            decode(u, b, h, qp, q_beg, q_len, k_beg, k_len);
            if (qp * 256 >= q_len) continue;
            const int row_in_seq = (2 * qp + (int)cta_rank) * 128 + r;
            const bool row_ok = row_in_seq < q_len;
            __nv_bfloat16* orow = p.out + (long long)(q_beg + row_in_seq) * p.ldo + h * 256 + half * OC;
            if (k_len <= 0) {                                  // empty key sequence: flash-attn returns zeros
                if (row_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
                    for (int i = 0; i < OC / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            float* slot = xch + (n & 1) * 512;                 // [max | sum][half][row]
            mbar_wait<true>(s_full, n & 1, 46);
            tc_fence_after();
            const int kv_valid = min(256, k_len) - half * NC;  // valid keys of this thread's half
            const uint32_t ts = tmem_base + lane_off + half * NC;
            // pass 1: row max of this half (nothing kept in registers)
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < NC / 32; ++c) {
                uint32_t s[32];
                tmem_ld_x32(ts + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(s[i]));
            }
            slot[half * 128 + r] = mx;
            named_bar_sync(1 + q, 64);
            mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);        // k_len >= 1: half 0 always has a valid key
            const float neg_m = -mx * p.scale_log2;
            // pass 2: p = 2^(s*scale - m) packed to bf16, row sum of this half
            uint32_t pk[NC / 2];
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int c = 0; c < NC / 32; ++c) {
                uint32_t s[32];
                tmem_ld_x32(ts + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const int kk = c * 32 + i;
                    const float p0 = (kk < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, neg_m)) : 0.f;
                    const float p1 = (kk + 1 < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, neg_m)) : 0.f;
                    rs0 += p0; rs1 += p1;
                    pk[kk >> 1] = pack_bf16x2(p0, p1);
                }
            }
            // row-sum exchange; the same barrier guarantees that BOTH threads of the row have finished reading S before
            // either overwrites it with P (half 1's P columns [64,128) are S columns of half 0's keys)
            slot[256 + half * 128 + r] = rs0 + rs1;
            tc_fence_before();
            named_bar_sync(1 + q, 64);
            tc_fence_after();
            const float l = (half == 0) ? (rs0 + rs1) + slot[256 + 128 + r] : slot[256 + r] + (rs0 + rs1);
            const uint32_t tp = tmem_base + lane_off + half * (NC / 2);
#pragma unroll
            for (int c = 0; c < NC / 64; ++c) tmem_st_x32(tp + c * 32, pk + c * 32);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_cluster(p_full, 0);
            __syncwarp();
            const float inv_l = 1.0f / l;
            mbar_wait<true>(pv_done, n & 1, 47);
            tc_fence_after();
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
            // the O reads above must be complete before this thread lets the next unit's PV overwrite O: it signals
            // p_full of the next unit only after them (program order + the tcgen05 fence before that arrive)
            ++n;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 1) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace flite
