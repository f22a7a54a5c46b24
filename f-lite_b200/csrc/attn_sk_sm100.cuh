// Persistent stream-K form of the 2-CTA flash attention (attn_fwd_cg2_kernel<1, true>), head_dim 256, sm_100a.
//
// Why: at C2 the self-attention launch has 2 x 12 x 17 = 408 (sequence, head, 256-query tile) units of 33 key tiles for
// 74 cluster slots: 5.51 waves of work take 6 waves of time (8 %), and every CTA pays barrier init + TMEM allocation +
// an exposed Q / first-K load.  Here the grid is ONE wave of clusters; the launch's  units x key-tiles  tile-steps are
// cut into equal contiguous shares (stream-K): cluster c runs tile-steps [c T / C, (c+1) T / C) of the order
// (sequence, head, query tile, key tile).  A share begins and ends inside a unit in general, so a unit is computed in at
// most two parts by neighbouring clusters:
//   * the cluster that owns the unit's LAST key tiles reaches it FIRST (it is the first segment of its share): it stores
//     its unnormalised partial (O fp32, running max m, row sum l) into its workspace slot and raises two flags;
//   * the cluster that owns the unit's FIRST key tiles reaches it LAST (last segment of its share): it picks the partner's
//     partial up (long since written), merges  O = (a O1 + a' O2) / (a l1 + a' l2),  a = 2^(m1 - m), a' = 2^(m2 - m),
//     and writes the bf16 rows.  Flags are reset by their reader, so the kernel is CUDA-graph replayable.
// Everything else is the cg2 kernel: S = Q K^T as 256 x 128 cta_group::2 MMAs (each CTA stages 64 key rows), P written
// as packed bf16 over S in TMEM and fed to the PV MMA from there, O (256 columns) in TMEM, lazy rescale, ragged last key
// tile narrowed to the valid keys.  Barrier phases run on a launch-global tile counter; Q is single-buffered (a q_empty
// barrier lets the producer reload it as soon as a segment's last S MMA has retired).
// Requires uniform sequence lengths (the DiT's image stream: every sequence has L tokens), checked on the device.
// Replaces flash_attn_varlen_func at f_lite/model.py:203-210 for the self-attention call.
#pragma once

#include "attn_cg2_sm100.cuh"

namespace flite {

constexpr int SK_SLOT_FLOATS = 256 * 256 + 2 * 256;      // per cluster: O [256 rows][256] | m [256] | l [256]
constexpr int SK_FLAG_BYTES = 2048;                      // 2 flags (one per CTA rank) x up to 256 cluster slots
constexpr int SK_MAX_CLUSTERS = 256;

struct AttnSkParams {
    const int* cu_q;
    const int* cu_k;
    __nv_bfloat16* out;
    long long ldo;
    int q_col0, k_col0, v_col0;
    float scale_log2;
    int B, H;
    int q_len, k_len;      // uniform per-sequence lengths
    int QT, NT;            // 256-row query tiles / 128-key tiles per (sequence, head)
    long long total;       // B * H * QT * NT tile-steps
    // Schedule: units [0, rr_units) are run WHOLE, round-robin (cluster c takes units c, c + C, c + 2C, ...: neighbouring
    // clusters walk the same key tiles of neighbouring query tiles in lock step, exactly like the one-cluster-per-unit
    // launch, and no unit of this part is split); the tile-steps of the remaining units are cut into equal contiguous
    // shares (stream-K).  rr_units = 0: pure stream-K; rr_units = all units: persistent, never splits a unit.
    long long rr_units;
    int debug;             // profiling experiments only (0 in production): bit3 = skip the epilogue of whole units
    int tma_out;           // 1 = whole units with 128 valid rows per CTA leave through shared memory + TMA box stores
    // ragged = 1 (round-robin schedule only: rr_units == all units): per-sequence query / key lengths from cu_q / cu_k, QT =
    // query tiles of the LONGEST sequence (units past a sequence's end are skipped), an empty key sequence gives zeros.
    // This is the cross-attention over the packed text context (f_lite/model.py:188-210) on the persistent kernel.
    int ragged;
    unsigned int* flags;   // workspace head: [cluster slot][cta rank], zero when idle
    float* slots;          // workspace body: [cluster slot][SK_SLOT_FLOATS]
    // Fused Ulysses return path (as in attn_fwd_cg2_kernel): when out_peer[0] != nullptr query row l of this rank's heads
    // is stored into the token owner's buffer over NVLink peer memory, out_peer[l / sp_lq][(b*sp_lq + l % sp_lq), sp_head0 + h],
    // as whole 256-byte row segments (each warp transposes its 32 rows through the otherwise unused P region of smem).
    __nv_bfloat16* out_peer[8];
    int sp_lq;
    int sp_head0;
};

// The segment walk of one cluster, identical in every warp role: whole units of the round-robin part first, then the
// cluster's contiguous share [t, t_end) of the remaining tile-steps.  A segment is (unit u, key tiles [j0, j1)).
struct SkWalk {
    long long u_rr, rr_units, t, t_end;
    int C, NT;
    FLITE_DEVICE SkWalk(const AttnSkParams& p, int cluster_id, int num_clusters) {
        C = num_clusters; NT = p.NT;
        u_rr = cluster_id; rr_units = p.rr_units;
        const long long t_off = p.rr_units * NT, rem = p.total - t_off;
        t = t_off + rem * cluster_id / num_clusters;
        t_end = t_off + rem * (cluster_id + 1) / num_clusters;
    }
    FLITE_DEVICE bool next(int& u, int& j0, int& j1) {
        if (u_rr < rr_units) { u = (int)u_rr; j0 = 0; j1 = NT; u_rr += C; return true; }
        if (t >= t_end) return false;
        u = (int)(t / NT); j0 = (int)(t - (long long)u * NT);
        j1 = (int)min((long long)NT, j0 + (t_end - t));
        t += j1 - j0;
        return true;
    }
};

// What one segment works on.  Uniform mode: every sequence has p.q_len / p.k_len tokens and [j0, j1) comes from the walk;
// ragged mode: the lengths of sequence b, whole units only ([0, nt)).  active == false: no MMA / load / barrier traffic at all
// (no query rows here, or no keys: the softmax warps then write zeros).
struct SkSeg {
    int b, h, qt, q_beg, q_len, k_beg, k_len, nt, tail_n;
    bool active;
};
FLITE_DEVICE SkSeg sk_segment(const AttnSkParams& p, int u, int& j0, int& j1) {
    SkSeg s;
    s.qt = u % p.QT;
    const int bh = u / p.QT;
    s.h = bh % p.H; s.b = bh / p.H;
    s.q_beg = p.cu_q[s.b]; s.k_beg = p.cu_k[s.b];
    if (p.ragged) {
        s.q_len = p.cu_q[s.b + 1] - s.q_beg;
        s.k_len = p.cu_k[s.b + 1] - s.k_beg;
    } else {
        s.q_len = p.q_len; s.k_len = p.k_len;
    }
    s.nt = (s.k_len + 127) / 128;
    s.tail_n = (s.k_len & 127) ? (((s.k_len & 127) + 15) & ~15) : 128;   // MMA extent of a ragged last key tile
    if (p.ragged) { j0 = 0; j1 = s.nt; }
    s.active = s.qt * 256 < s.q_len && j1 > j0;
    return s;
}

constexpr int SK_THREADS = 192;

__global__ void __launch_bounds__(SK_THREADS, 1)
attn_sk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
               const AttnSkParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // preconditions, same answer in every thread of every CTA, before any barrier: smem window, uniform lengths
    {
        bool bad = (smem - smem_raw) + ATT_SMEM_USED > ATT_SMEM;
        if (!p.ragged) {
            for (int b = 0; b < p.B; ++b)
                bad |= (p.cu_q[b + 1] - p.cu_q[b] != p.q_len) || (p.cu_k[b + 1] - p.cu_k[b] != p.k_len);
        } else {
            bad |= p.rr_units * p.NT != p.total;     // ragged lengths: whole units only
        }
        if (bad) {
            if (threadIdx.x == 0) atomicCAS(&g_flite_abort, 0u, (95u << 16) | 0x80000000u);
            return;
        }
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR);
    uint64_t* q_full = bars + 0;    //      leader: Q of the segment landed (both CTAs' halves)
    uint64_t* q_empty = bars + 1;   //      each CTA: the segment's last S MMA retired, Q reusable
    uint64_t* k_full = bars + 2;    // [2]  leader
    uint64_t* k_empty = bars + 4;   // [2]  each CTA
    uint64_t* v_full = bars + 6;    // [2]  leader
    uint64_t* v_empty = bars + 8;   // [2]  each CTA
    uint64_t* s_full = bars + 10;   // [2]  each CTA
    uint64_t* p_full = bars + 12;   //      leader, one arrival per softmax warp of both CTAs
    uint64_t* pv_done = bars + 13;  //      each CTA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp_idx = threadIdx.x >> 5;
    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            mbar_init(q_full, 1);
            mbar_init(q_empty, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&k_full[i], 1);
                mbar_init(&k_empty[i], 1);
                mbar_init(&v_full[i], 1);
                mbar_init(&v_empty[i], 1);
                mbar_init(&s_full[i], 1);
            }
            mbar_init(p_full, 2 * 4);
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

    // segment walk, identical in every role: SkWalk w(p, cluster_id, num_clusters); while (w.next(u, j0, j1)) { ... }

    if (warp_idx == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (elect_one()) {
            int g = 0, seg = 0;
            SkWalk w(p, cluster_id, num_clusters);
            int u, j0, j1;
            for (; w.next(u, j0, j1);) {
                const SkSeg sg = sk_segment(p, u, j0, j1);
                if (!sg.active) continue;
                const int h = sg.h, k_beg = sg.k_beg, NT = sg.nt, tail_n = sg.tail_n;
                const int q_row0 = sg.q_beg + sg.qt * 256 + (int)cta_rank * 128;
                if (seg > 0) mbar_wait<true>(q_empty, (seg - 1) & 1, 41);
                if (is_leader) mbar_arrive_expect_tx(q_full, 2 * 65536);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tma_load_2d_cg2(smem + ATT2_SQ + c * 16384, &tmap_q, q_full, 0, p.q_col0 + h * 256 + c * 64, q_row0);
                for (int j = j0; j < j1; ++j, ++g) {
                    const int st = g & 1;
                    const uint32_t ph = ((g >> 1) & 1) ^ 1;
                    const int krow = k_beg + j * 128;
                    const int kn_half = (j == NT - 1 ? tail_n : 128) >> 1;   // keys of this tile staged per CTA
                    mbar_wait<true>(&k_empty[st], ph, 42);
                    if (is_leader) mbar_arrive_expect_tx(&k_full[st], 2 * 32768);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        tma_load_2d_cg2(smem + ATT2_SK + st * 32768 + c * 8192, &tmap_k, &k_full[st], 0,
                                        p.k_col0 + h * 256 + c * 64, krow + (int)cta_rank * kn_half);
                    mbar_wait<true>(&v_empty[st], ph, 43);
                    if (is_leader) mbar_arrive_expect_tx(&v_full[st], 2 * 32768);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_2d_cg2(smem + ATT2_SV + st * 32768 + c * 16384, &tmap_v, &v_full[st], 0,
                                        p.v_col0 + h * 256 + (int)cta_rank * 128 + c * 64, krow);
                }
                ++seg;
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer (leader CTA) ================================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc_s_full = make_idesc_bf16(256, 128, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(256, 256, 0, 1);
            const uint32_t sq = smem_u32(smem + ATT2_SQ), sk = smem_u32(smem + ATT2_SK), sv = smem_u32(smem + ATT2_SV);
            int g0 = 0, seg = 0;
            SkWalk w(p, cluster_id, num_clusters);
            int u, j0, j1;
            for (; w.next(u, j0, j1);) {
                const SkSeg sg = sk_segment(p, u, j0, j1);
                if (!sg.active) continue;
                const int n = j1 - j0, NT = sg.nt, tail_n = sg.tail_n;
                const uint32_t idesc_s_tail = make_idesc_bf16(256, tail_n, 0, 0);
                auto issue_s = [&](int i) {      // S of the segment's i-th tile
                    const int g = g0 + i, j = j0 + i, st = g & 1;
                    mbar_wait<true>(&k_full[st], (g >> 1) & 1, 44);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (g & 1) * 128;
                    const uint32_t idesc_s = (j == NT - 1) ? idesc_s_tail : idesc_s_full;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const uint32_t offq = (k >> 2) * 16384 + (k & 3) * 32;
                        const uint32_t offk = st * 32768 + (k >> 2) * 8192 + (k & 3) * 32;
                        umma_ss<2>(d, make_smem_desc_sw128(sq + offq, 16, 1024), make_smem_desc_sw128(sk + offk, 16, 1024),
                                   idesc_s, k != 0 ? 1u : 0u);
                    }
                    umma_commit_cg2(&k_empty[st], 0x3);
                    umma_commit_cg2(&s_full[g & 1], 0x3);
                    if (i == n - 1) umma_commit_cg2(q_empty, 0x3);   // last S of the segment: Q may be reloaded
                };
                mbar_wait<true>(q_full, seg & 1, 45);
                issue_s(0);
                for (int i = 0; i < n; ++i) {
                    const int g = g0 + i, j = j0 + i, st = g & 1;
                    if (i + 1 < n) issue_s(i + 1);
                    // P of tile g from every softmax warp; for i == 0 this also orders the previous segment's O read-out
                    // (the same warps arrive here only after their epilogue) before the overwrite below
                    mbar_wait<true>(p_full, g & 1, 46);
                    mbar_wait<true>(&v_full[st], (g >> 1) & 1, 47);
                    tc_fence_after();
                    const int pv_steps = (j == NT - 1) ? (tail_n >> 4) : 8;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k >= pv_steps) break;
                        const uint64_t db = make_smem_desc_sw128(sv + st * 32768 + k * 2048, 16384, 1024);
                        umma_ts<2>(tmem_o, tmem_base + (g & 1) * 128 + k * 8, db, idesc_o, (i | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_cg2(&v_empty[st], 0x3);
                    umma_commit_cg2(pv_done, 0x3);
                }
                g0 += n;
                ++seg;
            }
        }
        __syncwarp();
    } else {
        // ================================ softmax / correction / epilogue ================================
        const int q = warp_idx & 3;
        const int lane = (int)lane_id();
        const int r = q * 32 + lane;                         // row inside this CTA's 128-query tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        int g0 = 0;
        SkWalk w(p, cluster_id, num_clusters);
        int u, j0, j1;
        while (w.next(u, j0, j1)) {
            const SkSeg sg = sk_segment(p, u, j0, j1);
            const int qt = sg.qt, h = sg.h, b = sg.b, NT = sg.nt;
            if (!sg.active) {
                // no keys (ragged mode): flash-attn returns zeros for the rows that exist; no rows: nothing to do
                const int row_in_seq0 = qt * 256 + (int)cta_rank * 128 + r;
                if (row_in_seq0 < sg.q_len) {
                    uint4* dst = reinterpret_cast<uint4*>(p.out + (long long)(sg.q_beg + row_in_seq0) * p.ldo + h * 256);
#pragma unroll
                    for (int x = 0; x < 32; ++x) dst[x] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const int n = j1 - j0;
            float m_used = -INFINITY, l = 0.f;
            for (int i = 0; i < n; ++i) {
                const int g = g0 + i, j = j0 + i;
                mbar_wait<true>(&s_full[g & 1], (g >> 1) & 1, 48);
                tc_fence_after();
                const uint32_t ts = tmem_base + lane_off + (g & 1) * 128;
                const int kv_valid = min(128, sg.k_len - j * 128);
                const bool full = kv_valid >= 128;
                uint32_t s[128];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_x32(ts + c * 32, s + c * 32);
                tmem_ld_wait();
                float mx = -INFINITY;
                if (full) {
#pragma unroll
                    for (int x = 0; x < 128; ++x) mx = fmaxf(mx, __uint_as_float(s[x]));
                } else {
#pragma unroll
                    for (int x = 0; x < 128; ++x)
                        if (x < kv_valid) mx = fmaxf(mx, __uint_as_float(s[x]));
                }
                const float m_new = fmaxf(m_used, mx * p.scale_log2);
                const bool need = (i > 0) && (m_new - m_used > 8.0f);
                const bool need_any = __any_sync(0xffffffffu, need);
                float corr = 1.0f;
                if (i == 0) {
                    m_used = m_new;
                } else if (need_any) {
                    corr = fast_exp2(m_used - m_new);
                    m_used = m_new;
                }
                uint32_t pk[64];
                float rs0 = 0.f, rs1 = 0.f;
                const float neg_m = -m_used;
                if (full) {
#pragma unroll
                    for (int x = 0; x < 128; x += 2) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(s[x]), p.scale_log2, neg_m));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(s[x + 1]), p.scale_log2, neg_m));
                        rs0 += p0; rs1 += p1;
                        pk[x >> 1] = pack_bf16x2(p0, p1);
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 128; x += 2) {
                        const float p0 = (x < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[x]), p.scale_log2, neg_m)) : 0.f;
                        const float p1 = (x + 1 < kv_valid) ? fast_exp2(fmaf(__uint_as_float(s[x + 1]), p.scale_log2, neg_m)) : 0.f;
                        rs0 += p0; rs1 += p1;
                        pk[x >> 1] = pack_bf16x2(p0, p1);
                    }
                }
                l = l * corr + (rs0 + rs1);
                if (i > 0) {
                    // Every phase of pv_done is consumed, in order: a parity wait is only meaningful while the barrier is in
                    // the awaited phase or the one after it, so phases must not be skipped (PV_{g-1} started when this
                    // tile's S retired and is normally done by now -- the wait costs nothing in steady state).
                    mbar_wait<true>(pv_done, (g - 1) & 1, 49);
                    if (need_any) {              // O rescale
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < 8; ++c) {
                            uint32_t o[32];
                            tmem_ld_x32(tmem_o + lane_off + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int x = 0; x < 32; ++x) o[x] = __float_as_uint(__uint_as_float(o[x]) * corr);
                            tmem_st_x32(tmem_o + lane_off + c * 32, o);
                        }
                        tmem_st_wait();
                    }
                }
                // P (packed bf16 pairs) over the first 64 columns of S_g
                const uint32_t tp = tmem_base + lane_off + (g & 1) * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) tmem_st_x32(tp + c * 32, pk + c * 32);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (elect_one()) mbar_arrive_cluster(p_full, 0);
                __syncwarp();
            }
            // ---- segment epilogue ----
            mbar_wait<true>(pv_done, (g0 + n - 1) & 1, 50);
            tc_fence_after();
            const int row_in_seq = qt * 256 + (int)cta_rank * 128 + r;
            const bool row_ok = row_in_seq < sg.q_len;
            const bool is_writer = j0 > 0;             // this cluster owns the unit's LAST key tiles: publish a partial
            const bool is_reader = j1 < NT;            // ... the FIRST key tiles: merge the partner's partial and finish
            // final rows of this unit: out = O * w1 (+ partner partial * w2), bf16; local rows are stored directly, rows
            // that belong to a peer GPU are staged through smem and leave as whole 256-byte segments
            auto emit_rows = [&](float w1, float w2, const float4* prow) {
                __nv_bfloat16* orow = p.out + (long long)(sg.q_beg + row_in_seq) * p.ldo + h * 256;
                const bool to_peer = p.out_peer[0] != nullptr;
                if (to_peer && row_ok) {
                    const int owner = row_in_seq / p.sp_lq, li = row_in_seq - owner * p.sp_lq;
                    orow = p.out_peer[owner] + ((long long)b * p.sp_lq + li) * p.ldo + (p.sp_head0 + h) * 256;
                }
                uint8_t* stg = smem + ATT2_SP + (warp_idx - 2) * (32 * 256);      // 32 rows x 128 columns per warp
                const unsigned long long my_ptr = row_ok ? (unsigned long long)orow : 0ull;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_off + c * 32, o);
                    float4 pv[8];
                    if (prow != nullptr) {
#pragma unroll
                        for (int x = 0; x < 8; ++x) pv[x] = __ldcg(prow + c * 8 + x);
                    } else {
#pragma unroll
                        for (int x = 0; x < 8; ++x) pv[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tmem_ld_wait();
                    uint4 w[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float4 pa = pv[2 * x], pb = pv[2 * x + 1];
                        w[x] = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * x]) * w1 + pa.x * w2, __uint_as_float(o[8 * x + 1]) * w1 + pa.y * w2),
                            pack_bf16x2(__uint_as_float(o[8 * x + 2]) * w1 + pa.z * w2, __uint_as_float(o[8 * x + 3]) * w1 + pa.w * w2),
                            pack_bf16x2(__uint_as_float(o[8 * x + 4]) * w1 + pb.x * w2, __uint_as_float(o[8 * x + 5]) * w1 + pb.y * w2),
                            pack_bf16x2(__uint_as_float(o[8 * x + 6]) * w1 + pb.z * w2, __uint_as_float(o[8 * x + 7]) * w1 + pb.w * w2));
                    }
                    if (!to_peer) {
                        if (row_ok) {
                            uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                            for (int x = 0; x < 4; ++x) dst[x] = w[x];
                        }
                        continue;
                    }
                    // stage: lane = row on the way in (16-byte chunks XOR-swizzled by the row), 128 columns per pass
                    const int cc = c & 3;
#pragma unroll
                    for (int x = 0; x < 4; ++x)
                        *reinterpret_cast<uint4*>(stg + lane * 256 + (((cc * 4 + x) ^ (lane & 15)) << 4)) = w[x];
                    if (cc == 3) {
                        __syncwarp();
                        const int ch = lane & 15;                     // 16-byte chunk of the 256-byte segment
#pragma unroll 4
                        for (int i2 = 0; i2 < 16; ++i2) {
                            const int rr = 2 * i2 + (lane >> 4);
                            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 256 + ((ch ^ (rr & 15)) << 4));
                            const unsigned long long rp = __shfl_sync(0xffffffffu, my_ptr, rr);
                            if (rp != 0ull) reinterpret_cast<uint4*>(rp)[(c >> 2) * 16 + ch] = v;
                        }
                        __syncwarp();
                    }
                }
            };
            if (is_writer && is_reader) {              // a share shorter than one unit: the host never launches that
                if (lane == 0) atomicCAS(&g_flite_abort, 0u, (94u << 16) | 0x80000000u);
            }
            if (is_writer) {
                float* slot = p.slots + (long long)cluster_id * SK_SLOT_FLOATS;
                const int srow = (int)cta_rank * 128 + r;
                float4* orow = reinterpret_cast<float4*>(slot + (long long)srow * 256);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint32_t o[32];
                    tmem_ld_x32(tmem_o + lane_off + c * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int x = 0; x < 8; ++x)
                        orow[c * 8 + x] = make_float4(__uint_as_float(o[4 * x]), __uint_as_float(o[4 * x + 1]),
                                                      __uint_as_float(o[4 * x + 2]), __uint_as_float(o[4 * x + 3]));
                }
                slot[256 * 256 + srow] = m_used;
                slot[256 * 256 + 256 + srow] = l;
                __threadfence();
                named_bar_sync(1, 128);                // all 128 rows of this CTA are in the slot
                if (r == 0) {
                    unsigned int* f = p.flags + cluster_id * 2 + cta_rank;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
                }
            } else if (is_reader) {
                const int pc = cluster_id + 1;         // the next cluster's share starts with the rest of this unit
                const unsigned int* f = p.flags + pc * 2 + cta_rank;
                {
                    const uint64_t t0 = globaltimer_ns();
                    unsigned int v = 0;
                    while (true) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                        if (v != 0) break;
                        if (*(volatile unsigned int*)&g_flite_abort != 0) break;
                        if (globaltimer_ns() - t0 > 2000000000ull) {
                            atomicCAS(&g_flite_abort, 0u, (93u << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
                            break;
                        }
                    }
                }
                const float* slot = p.slots + (long long)pc * SK_SLOT_FLOATS;
                const int srow = (int)cta_rank * 128 + r;
                const float m2 = __ldcg(slot + 256 * 256 + srow), l2 = __ldcg(slot + 256 * 256 + 256 + srow);
                const float m = fmaxf(m_used, m2);
                const float a1 = fast_exp2(m_used - m), a2 = fast_exp2(m2 - m);
                const float inv = 1.0f / (a1 * l + a2 * l2);
                const float w1 = a1 * inv, w2 = a2 * inv;
                const float4* prow = reinterpret_cast<const float4*>(slot + (long long)srow * 256);
                emit_rows(w1, w2, prow);
                named_bar_sync(1, 128);                // every row of this CTA has consumed the partial
                if (r == 0) *const_cast<volatile unsigned int*>(f) = 0u;   // idle again (graph replay / next launch)
            } else if (p.debug & 8) {
            } else if (p.tma_out && p.out_peer[0] == nullptr && qt * 256 + (int)cta_rank * 128 + 128 <= sg.q_len) {
                // Whole tile: O / l -> bf16 -> the P region of shared memory (unused: P lives in TMEM) as [128 rows x 64
                // columns] boxes with the tensor map's 128-byte swizzle, two passes of 128 columns; lane 0 of warp 2 issues the
                // TMA stores (and owns their bulk groups).  The per-thread row stores of emit_rows touch 32 different cache
                // lines per instruction (~8k cycles per unit); these are conflict-free st.shared + four bulk stores.
                const float inv_l = 1.0f / l;
                const bool issuer = warp_idx == 2 && lane == 0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    if (issuer) tma_store_wait_read<0>();            // the previous stores have read the staging region
                    named_bar_sync(2, 128);
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t o[32];
                        tmem_ld_x32(tmem_o + lane_off + pass * 128 + c * 32, o);
                        tmem_ld_wait();
                        uint8_t* box = smem + ATT2_SP + (c >> 1) * 16384 + r * 128;
                        const int ch0 = (c & 1) * 4;
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            *reinterpret_cast<uint4*>(box + (((ch0 + x) ^ (r & 7)) << 4)) = make_uint4(
                                pack_bf16x2(__uint_as_float(o[8 * x]) * inv_l, __uint_as_float(o[8 * x + 1]) * inv_l),
                                pack_bf16x2(__uint_as_float(o[8 * x + 2]) * inv_l, __uint_as_float(o[8 * x + 3]) * inv_l),
                                pack_bf16x2(__uint_as_float(o[8 * x + 4]) * inv_l, __uint_as_float(o[8 * x + 5]) * inv_l),
                                pack_bf16x2(__uint_as_float(o[8 * x + 6]) * inv_l, __uint_as_float(o[8 * x + 7]) * inv_l));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(2, 128);
                    if (issuer) {
                        const int row0 = sg.q_beg + qt * 256 + (int)cta_rank * 128;
                        tma_store_2d(&tmap_o, smem + ATT2_SP, h * 256 + pass * 128, row0);
                        tma_store_2d(&tmap_o, smem + ATT2_SP + 16384, h * 256 + pass * 128 + 64, row0);
                        tma_store_commit();
                    }
                }
            } else {
                emit_rows(1.0f / l, 0.f, nullptr);
            }
            tc_fence_before();     // this segment's TMEM reads are ordered before the p_full arrival of the next one
            g0 += n;
        }
    }

    if (warp_idx == 2 && lane_id() == 0) tma_store_wait_read<0>();   // shared memory stays valid until the bulk stores have read it
    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 1) tmem_dealloc<2>(tmem_base, 512);
}

}  // namespace flite
