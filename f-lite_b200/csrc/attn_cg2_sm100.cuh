// 2-CTA (cta_group::2) varlen flash attention forward, head_dim 256, sm_100a.
//
// A cluster of two CTAs processes two adjacent 128-query tiles of the same (head, sequence) and shares every
// K/V tile: with tcgen05.mma.cta_group::2 the pair computes S = Q K^T as a 256 x 128 MMA (each CTA stages only
// 64 of the 128 key rows) and O += P V as a 256 x 256 MMA (each CTA stages only 128 of the 256 head-dim columns
// of V), so K/V traffic from L2 per query row is halved relative to attn_fwd_kernel and the freed shared memory
// double-buffers K and V (loads of tile j+1 overlap the MMAs of tile j).
//   SMEM / CTA: Q 64K | K half-tile 32K x2 | V half-tile 32K x2 | P 32K | barriers | exchange
//   TMEM / CTA: S0 [0,128) | S1 [128,256) | O [256,512)   (own 128 query rows)
// kPTmem: P_j is written as packed bf16 into the TMEM columns of S_j (aliased; 64 columns) and fed to the PV MMA
// as the A operand straight from TMEM (tcgen05.mma [d], [a_tmem], b_desc) -- no shared-memory round trip for P.
// Roles per CTA: warp 0 TMA producer (own halves; complete_tx on the leader's barriers), warp 1 MMA issuer
// (leader CTA only; commits multicast to both CTAs), warps 2.. softmax warpgroups (as in attn_fwd_kernel).
#pragma once

#include "attn_sm100.cuh"

namespace flite {

constexpr int ATT2_SQ = 0, ATT2_SK = 65536, ATT2_SV = 131072, ATT2_SP = 196608;

template <int kWG, bool kPTmem>
__global__ void __launch_bounds__(64 + 128 * kWG, 1)
attn_fwd_cg2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                    const AttnParams p) {
    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    pdl_launch_dependents();   // PDL: the next kernel may be scheduled as SMs drain; it waits for this grid to finish
    pdl_wait();                // q/k/v (and cu_seqlens) come from earlier kernels of the stream
    const int q_beg = p.cu_q[b], q_len = p.cu_q[b + 1] - q_beg;
    if ((qt & ~1) * 128 >= q_len) return;  // uniform for the whole cluster, before any barrier / TMEM allocation
    const int k_beg = p.cu_k[b], k_len = p.cu_k[b + 1] - k_beg;
    const int n_tiles = (k_len + 127) / 128;
    // Ragged last key tile (4112 = 32*128 + 16 at C2): its MMAs only cover the valid keys rounded up to 16 -- S = Q K^T
    // with N = tail_n instead of 128 (each CTA then stages keys [rank*tail_n/2, +tail_n/2) of the tile) and O += P V over
    // tail_n/16 K-steps instead of 8.  The softmax masks columns >= the valid count as before.
    const int tail_n = (k_len > 0 && (k_len & 127)) ? (((k_len & 127) + 15) & ~15) : 128;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const bool smem_bad = (smem - smem_raw) + ATT_SMEM_USED > ATT_SMEM;   // same value in both CTAs
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR);
    float* xch = reinterpret_cast<float*>(smem + ATT_XCH);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]  leader
    uint64_t* k_empty = bars + 3;   // [2]  each CTA
    uint64_t* v_full = bars + 5;    // [2]  leader
    uint64_t* v_empty = bars + 7;   // [2]  each CTA
    uint64_t* s_full = bars + 9;    // [2]  each CTA
    uint64_t* p_full = bars + 11;   //      leader, one arrival per softmax warp of both CTAs
    uint64_t* pv_done = bars + 12;  //      each CTA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp_idx = threadIdx.x >> 5;
    if (smem_bad) {
        if (threadIdx.x == 0) atomicCAS(&g_flite_abort, 0u, (98u << 16) | 0x80000000u);
        return;
    }
    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            mbar_init(q_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 1);
                mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1);
                mbar_init(&v_empty[i], 1);
                mbar_init(&s_full[i], 1);
            }
            mbar_init(p_full, 2 * 4 * kWG);
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

    const int q_row0 = q_beg + qt * 128;

    if (warp_idx == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (elect_one() && n_tiles > 0) {
            if (is_leader) mbar_arrive_expect_tx(q_full, 2 * 65536);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tma_load_2d_cg2(smem + ATT2_SQ + c * 16384, &tmap_q, q_full, 0, p.q_col0 + h * 256 + c * 64, q_row0);
            for (int j = 0; j < n_tiles; ++j) {
                if ((p.debug & 2) && j >= 2) break;
                const int st = j & 1;
                const uint32_t ph = ((j >> 1) & 1) ^ 1;
                const int krow = k_beg + j * 128;
                mbar_wait<true>(&k_empty[st], ph, 21);
                if (is_leader) mbar_arrive_expect_tx(&k_full[st], 2 * 32768);
                const int kn_half = (j == n_tiles - 1 ? tail_n : 128) >> 1;   // keys of this tile staged per CTA
#pragma unroll
                for (int c = 0; c < 4; ++c)   // this CTA's 64 key rows, 4 chunks of 64 head-dim columns
                    tma_load_2d_cg2(smem + ATT2_SK + st * 32768 + c * 8192, &tmap_k, &k_full[st], 0,
                                    p.k_col0 + h * 256 + c * 64, krow + (int)cta_rank * kn_half);
                mbar_wait<true>(&v_empty[st], ph, 22);
                if (is_leader) mbar_arrive_expect_tx(&v_full[st], 2 * 32768);
#pragma unroll
                for (int c = 0; c < 2; ++c)   // all 128 key rows, this CTA's 128 head-dim columns
                    tma_load_2d_cg2(smem + ATT2_SV + st * 32768 + c * 16384, &tmap_v, &v_full[st], 0,
                                    p.v_col0 + h * 256 + (int)cta_rank * 128 + c * 64, krow);
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer (leader CTA) ================================
        if (is_leader && elect_one() && n_tiles > 0) {
            constexpr uint32_t idesc_s_full = make_idesc_bf16(256, 128, 0, 0);   // Q (K-major) x K (K-major)
            const uint32_t idesc_s_tail = make_idesc_bf16(256, tail_n, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(256, 256, 0, 1);   // P (K-major) x V (MN-major)
            const uint32_t sq = smem_u32(smem + ATT2_SQ), sk = smem_u32(smem + ATT2_SK);
            const uint32_t sv = smem_u32(smem + ATT2_SV), sp = smem_u32(smem + ATT2_SP);
            auto issue_s = [&](int j) {
                const int st = j & 1;
                if (!((p.debug & 2) && j >= 2)) mbar_wait<true>(&k_full[st], (j >> 1) & 1, 23);
                tc_fence_after();
                const uint32_t d = tmem_base + (j & 1) * 128;
                const uint32_t idesc_s = (j == n_tiles - 1) ? idesc_s_tail : idesc_s_full;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t offq = (k >> 2) * 16384 + (k & 3) * 32;
                    const uint32_t offk = st * 32768 + (k >> 2) * 8192 + (k & 3) * 32;
                    umma_ss<2>(d, make_smem_desc_sw128(sq + offq, 16, 1024), make_smem_desc_sw128(sk + offk, 16, 1024),
                               idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit_cg2(&k_empty[st], 0x3);
                umma_commit_cg2(&s_full[j & 1], 0x3);
            };
            mbar_wait<true>(q_full, 0, 24);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                if (j + 1 < n_tiles) issue_s(j + 1);
                mbar_wait<true>(p_full, j & 1, 25);
                if (!((p.debug & 2) && j >= 2)) mbar_wait<true>(&v_full[st], (j >> 1) & 1, 26);
                tc_fence_after();
                const int pv_steps = (j == n_tiles - 1) ? (tail_n >> 4) : 8;   // 16 keys per K-step
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k >= pv_steps) break;
                    const uint64_t db = make_smem_desc_sw128(sv + st * 32768 + k * 2048, 16384, 1024);
                    if constexpr (kPTmem) {
                        // A = P_j: rows = lanes, 16 keys = 8 packed columns per K-step, inside S_j's columns
                        umma_ts<2>(tmem_o, tmem_base + (j & 1) * 128 + k * 8, db, idesc_o, (j | k) != 0 ? 1u : 0u);
                    } else {
                        const uint64_t da = make_smem_desc_sw128(sp + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
                        umma_ss<2>(tmem_o, da, db, idesc_o, (j | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit_cg2(&v_empty[st], 0x3);
                umma_commit_cg2(pv_done, 0x3);
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
            mbar_wait<true>(&s_full[j & 1], (j >> 1) & 1, 27);
            tc_fence_after();
            if (p.debug & 1) {   // experiment: no softmax math, P left as is
                if (j > 0) mbar_wait<true>(pv_done, (j - 1) & 1, 28);
                l = 1.f;
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_cluster(p_full, 0);
                __syncwarp();
                continue;
            }
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
            // smem P: P_{j-1} V_{j-1} must be complete before P is overwritten.  TMEM P lives in S_j's own columns, so only an
            // O rescale NEEDS the previous PV -- but every phase of pv_done is still consumed, in order: a parity wait is
            // only meaningful while the barrier is in the awaited phase or the one after it, and the epilogue's wait for
            // the last phase must not be able to run two phases ahead of a skipped one.  PV_{j-1} started when S_j
            // retired, so by this point it is normally done and the wait is free.
            if (j > 0) {
                mbar_wait<true>(pv_done, (j - 1) & 1, 28);
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
            if constexpr (kPTmem) {
                // P (packed bf16 pairs) over the first 64 columns of S_j.  Every thread of this row group has
                // finished its tcgen05.ld of S_j (kWG == 2: guaranteed by the row-max exchange barrier above).
                const uint32_t tp = tmem_base + lane_off + (j & 1) * 128 + half * (NC / 2);
#pragma unroll
                for (int c = 0; c < NC / 64; ++c) tmem_st_x32(tp + c * 32, pk + c * 32);
                tmem_st_wait();
            } else {
                // write P (bf16, K-major, 128B swizzle): 64-key chunks of [128 rows x 128 B]
                uint8_t* sp_row = smem + ATT2_SP + r * 128 + ((kWG == 2) ? half * 16384 : 0);
#pragma unroll
                for (int u = 0; u < NC / 8; ++u) {
                    const int chunk = u >> 3, unit = u & 7;
                    uint4 v = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    *reinterpret_cast<uint4*>(sp_row + chunk * 16384 + ((unit ^ (r & 7)) << 4)) = v;
                }
                fence_proxy_async_smem();
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive_cluster(p_full, 0);   // P of both CTAs is consumed by the leader's MMAs
            __syncwarp();
        }
        // ---- epilogue: O / l -> bf16 -> out[row, h*256 + c] ----
        const int row_in_seq = qt * 128 + r;
        const bool row_ok = row_in_seq < q_len;
        __nv_bfloat16* orow = p.out + (long long)(q_beg + row_in_seq) * p.ldo + h * 256 + half * OC;
        if (p.out_peer[0] != nullptr && row_ok) {   // fused all-to-all: write to the rank that owns this token
            const int owner = row_in_seq / p.sp_lq, li = row_in_seq - owner * p.sp_lq;
            orow = p.out_peer[owner] + ((long long)b * p.sp_lq + li) * p.ldo + (p.sp_head0 + h) * 256 + half * OC;
        }
        if (n_tiles > 0) {
            if constexpr (kWG == 2) {
                // the parity slot of the (non-existent) next tile is free: tile n-2's reads all precede barrier n-1
                float* slot = xch + (n_tiles & 1) * 256;
                slot[half * 128 + r] = l;
                named_bar_sync(1 + q, 64);
                l += slot[(half ^ 1) * 128 + r];
            }
            mbar_wait<true>(pv_done, (n_tiles - 1) & 1, 29);
            tc_fence_after();
            const float inv_l = 1.0f / l;
            if (p.tma_out && qt * 128 + 128 <= q_len) {
                // Whole tile (uniform for the CTA): O / l -> bf16 -> the Q region (dead: every MMA has retired), laid out as
                // four [128 rows x 64 columns] boxes with the 128-byte swizzle of the tensor map; one thread issues the
                // four TMA stores.  Conflict-free st.shared (16-byte chunk index XOR row), no per-thread global stores.
#pragma unroll 1
                for (int c = 0; c < OC / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_off + half * OC + c * 32, o);
                    tmem_ld_wait();
                    const int col = half * OC + c * 32;                      // first of 32 output columns
                    uint8_t* box = smem + ATT2_SQ + (col >> 6) * 16384 + r * 128;
                    const int ch0 = (col & 63) >> 3;                         // first 16-byte chunk inside the 128-byte row
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(box + (((ch0 + i) ^ (r & 7)) << 4)) = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l));
                }
                fence_proxy_async_smem();
                named_bar_sync(5, 128 * kWG);
                if (warp_idx == 2 && elect_one()) {
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb)
                        tma_store_2d(&tmap_o, smem + ATT2_SQ + cb * 16384, h * 256 + cb * 64, q_beg + qt * 128);
                    tma_store_commit();
                    tma_store_wait_read<0>();      // shared memory must stay valid until the bulk stores have read it
                }
                __syncwarp();
            } else if (p.stage_out) {
                // Whole-row stores (NVLink peer destinations): all MMAs have completed (pv_done), so the Q tile in
                // shared memory is dead; each warp transposes its 32 rows through it -- lane = row on the way in,
                // lane = 16-byte chunk on the way out (32 / CH rows of OC*2 contiguous bytes per store instruction).
                constexpr int CH = OC / 8;                       // 16-byte chunks per row segment
                uint8_t* stg = smem + ATT2_SQ + (warp_idx - 2) * (32 * OC * 2);
                const unsigned long long my_ptr = row_ok ? (unsigned long long)orow : 0ull;
#pragma unroll 1
                for (int c = 0; c < OC / 32; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_off + half * OC + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(stg + lane * (OC * 2) + (((c * 4 + i) ^ (lane & (CH - 1))) << 4)) = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * i]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l));
                }
                __syncwarp();
                const int cc = lane & (CH - 1);
#pragma unroll 4
                for (int i2 = 0; i2 < CH; ++i2) {
                    const int rr = i2 * (32 / CH) + lane / CH;
                    const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * (OC * 2) + ((cc ^ (rr & (CH - 1))) << 4));
                    const unsigned long long rp = __shfl_sync(0xffffffffu, my_ptr, rr);
                    if (rp != 0ull) reinterpret_cast<uint4*>(rp)[cc] = v;
                }
            } else {
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
            }
        } else if (row_ok) {
            // empty key sequence: flash-attn returns zeros
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
