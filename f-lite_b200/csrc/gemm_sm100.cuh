// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   warp 0      : TMA producer (A / W tiles -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> fused math -> global), one output row per thread
//
// kCtaGroup == 2 pairs two SMs (cluster of 2) on one 256 x BLOCK_N tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and half of the W tile, the leader CTA issues the MMAs and
// multicasts the commits to both CTAs' barriers.
//
// Replaces the cuBLASLt calls behind nn.Linear in the reference (f_lite/model.py:151-156,162,212,267,
// 436,448-454,472,475) and fuses what the reference runs as separate elementwise kernels:
//   EPI_STORE     : (+bias) [SiLU]                          model.py:162,189,448-452
//   EPI_GATED_RES : x + (acc) * gate[sample]                model.py:289,294-297,301
//   EPI_SWIGLU    : silu(gate) * up on interleaved columns  liger swiglu, model.py:267
//   EPI_QKV_ROPE  : +bias, 2-D RoPE, QK-RMSNorm per head     model.py:162-183,403-414,92-108
#pragma once

#include "common.cuh"

namespace flite {

enum GemmEpilogue : int { EPI_STORE = 0, EPI_GATED_RES = 1, EPI_SWIGLU = 2, EPI_QKV_ROPE = 3 };

struct GemmParams {
    int M, N, K;
    __nv_bfloat16* C;
    long long ldc;
    const __nv_bfloat16* bias;   // [N] or nullptr
    int act;                     // EPI_STORE: 0 none, 1 SiLU after the bf16 rounding of (acc + bias)
    const __nv_bfloat16* resid;  // EPI_GATED_RES: [M, ldr]
    long long ldr;
    const __nv_bfloat16* gate;   // EPI_GATED_RES: gate[(row / rows_per_sample) * ld_gate + col]
    long long ld_gate;
    int rows_per_sample;
    // EPI_QKV_ROPE: columns are (k in {q,k,v}, head, 256); rope tables [rows_per_sample, 128] bf16
    const __nv_bfloat16* rope_cos;
    const __nv_bfloat16* rope_sin;
    int qk_cols;                 // columns [0, qk_cols) get RoPE (if tables != null) + RMSNorm; rest plain
    float eps;
    // EPI_QKV_ROPE, Ulysses sequence parallelism: write each head straight into the all-to-all send layout
    // [sample][dest rank][local token][q|k|v][head % sp_hp][256]  (sp_ranks == 0: plain row-major [M, N])
    int sp_ranks;
    int sp_hp;                   // heads per rank
    int n_heads;
    // Fused exchange: when sp_peer[0] != nullptr the head is stored straight into the DESTINATION rank's receive buffer
    // over NVLink peer memory ([sample][full sequence][q|k|v of that rank's heads], row = smp*sp_seq + sp_rank*Lq + li)
    // instead of a local all-to-all send buffer.
    __nv_bfloat16* sp_peer[8];
    int sp_rank;
    int sp_seq;
    // EPI_QKV_ROPE: 1 = transpose each warp's 32 rows through shared memory so that every store instruction writes
    // whole 256-byte row segments (needed for NVLink efficiency when the destination is a peer GPU)
    int stage_stores;
    // Tail balancing: work units [0, full_units) are whole tiles; units beyond are HALF tiles (BLOCK_N/2 columns)
    // of the remaining tiles, so that a last partial wave is spread over twice as many clusters.
    int full_units;              // == number of tiles when no split is used
    int num_units;
    // Rasterisation: tiles are ordered band by band (band_m consecutive M-tiles), N-tile by N-tile inside a band and
    // M fastest, so that a band of A (band_m x TILE_M x K) stays L2-resident while W streams past it once per band.
    int band_m;
    // Ragged M (8224 = 32*256 + 32 at C2): when the last 256-row M-tile holds <= 128 valid rows it is run as a NARROW tile,
    // tcgen05.mma cta_group::2 with M = 128 (64 rows per CTA) -- measured 64 instead of 128 cycles per K-step
    // (tools/probe_umma.cu), i.e. the padding costs half.  narrow_m = 1 enables it; the narrow tile is logical M-block 0
    // of the rasterisation (so it never lands in the half-width tail units).  TMEM layout of that mode ("layout B"):
    // lanes 0-63 = rows 0-63 x accumulator columns [0, N/2), lanes 64-127 = the same rows x columns [N/2, N), both at
    // TMEM columns [0, N/2).
    int narrow_m;
    // L2 eviction-priority hints of the A / W tile loads (0 = plain load); chosen with the band height so that the operand
    // the rasterisation keeps resident is evict_last and the one that streams past it is evict_first.
    unsigned long long hint_a, hint_b;
    int debug;                   // profiling experiments only (0 in production): bit0 = gated-residual epilogue without its global loads / stores
};

constexpr int GEMM_BLOCK_K = 64;
// warps: 0 TMA producer, 1 MMA issuer, 2.. epilogue.  The QKV epilogue (RoPE + QK-norm over a whole 256-column head per
// row) is the heaviest one: it runs on TWO warpgroups, each owning one half of the (j, j+128) column pairs of every row.
__host__ __device__ constexpr int gemm_epi_warpgroups(int epi) { return epi == 3 ? 2 : 1; }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + 128 * gemm_epi_warpgroups(epi); }

template <int kCtaGroup, int BLOCK_N, int kStages>
struct GemmSmem {
    static constexpr int BN_CTA = BLOCK_N / kCtaGroup;
    static constexpr int A_BYTES = 128 * GEMM_BLOCK_K * 2;
    static constexpr int B_BYTES = BN_CTA * GEMM_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = kStages * STAGE_BYTES;
    static constexpr int STAGING_OFFSET = BAR_OFFSET + 256;           // 4 epilogue warps x 32 rows x 256 B
    static constexpr int STAGING_BYTES = 4 * 32 * 256;
    static constexpr int XCH_OFFSET = STAGING_OFFSET + STAGING_BYTES;  // EPI_QKV_ROPE: [tile parity][column half][128 rows] fp32
    static constexpr int XCH_BYTES = 2 * 2 * 128 * 4;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
    // QKV epilogue: + staging + sum-of-squares exchange; 227 KB leaves 768 B of alignment slack (checked at run time)
    static constexpr int QKV_USED = XCH_OFFSET + XCH_BYTES;
    static constexpr int TOTAL_QKV = (QKV_USED + 1024 <= 232448) ? QKV_USED + 1024 : 232448;
};

template <int kCtaGroup, int BLOCK_N, int kStages, int kEpi>
__global__ void __launch_bounds__(gemm_threads(kEpi), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_b_half, const __grid_constant__ CUtensorMap tmap_a_half,
                 const GemmParams p) {
    using S = GemmSmem<kCtaGroup, BLOCK_N, kStages>;
    constexpr int TILE_M = 128 * kCtaGroup;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512, "accumulators are double-buffered in TMEM");
    static_assert(BLOCK_N % 32 == 0, "epilogue works on 32-column chunks");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0;
    const bool is_leader = cta_rank == 0;
    if constexpr (kEpi == EPI_QKV_ROPE) {
        // the QKV layout leaves < 1 KB of alignment slack: refuse to run (same answer in both CTAs, before any barrier)
        if ((smem - smem_raw) + S::QKV_USED > S::TOTAL_QKV) {
            if (threadIdx.x == 0) atomicCAS(&g_flite_abort, 0u, (97u << 16) | 0x80000000u);
            return;
        }
    }

    if (warp_idx == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp_idx == 1) {
        if (elect_one()) {
            for (int i = 0; i < kStages; ++i) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tmem_full_bar[i], 1);
                mbar_init(&tmem_empty_bar[i], 4 * kCtaGroup * gemm_epi_warpgroups(kEpi));
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<kCtaGroup>(tmem_ptr_smem, TMEM_COLS);
    }
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // PDL: barrier init / TMEM allocation above overlap the previous kernel's tail; no global access before this point
    pdl_launch_dependents();
    pdl_wait();

    const int num_m_tiles = (p.M + TILE_M - 1) / TILE_M;
    const int num_n_tiles = p.N / BLOCK_N;
    const int num_tiles = num_m_tiles * num_n_tiles;
    (void)num_tiles;
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        const int per_band = p.band_m * num_n_tiles;
        const int band = tile / per_band, rem = tile - band * per_band;
        const int m_first = band * p.band_m;
        const int gm = min(p.band_m, num_m_tiles - m_first);
        m_blk = m_first + rem % gm;
        n_blk = rem / gm;
        if (p.narrow_m) m_blk = (m_blk == 0) ? num_m_tiles - 1 : m_blk - 1;   // logical block 0 = the narrow (last) M-tile
    };
    auto is_narrow = [&](int m_blk) { return kCtaGroup == 2 && p.narrow_m && m_blk == num_m_tiles - 1; };
    // unit -> (tile, half): half == -1 means the whole tile, 0/1 the left/right BLOCK_N/2 columns
    auto unit_tile = [&](int u, int& tile, int& half) {
        if (u < p.full_units) { tile = u; half = -1; }
        else { tile = p.full_units + ((u - p.full_units) >> 1); half = (u - p.full_units) & 1; }
    };
    const int num_kb = p.K / GEMM_BLOCK_K;
    const int cluster_id = blockIdx.x / kCtaGroup;
    const int num_clusters = gridDim.x / kCtaGroup;

    if (warp_idx == 0) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = cluster_id; unit < p.num_units; unit += num_clusters) {
                int tile, half;
                unit_tile(unit, tile, half);
                const int bn_cta = (half < 0) ? S::BN_CTA : S::BN_CTA / 2;       // W rows this CTA stages
                int m_blk, n_blk;
                tile_coords(tile, m_blk, n_blk);
                const bool narrow = is_narrow(m_blk);
                const int m0 = m_blk * TILE_M + (int)cta_rank * (narrow ? 64 : 128);
                const int n0 = n_blk * BLOCK_N + (half > 0 ? BLOCK_N / 2 : 0) + (int)cta_rank * bn_cta;
                const CUtensorMap* tb = (half < 0) ? &tmap_b : &tmap_b_half;
                const CUtensorMap* ta = narrow ? &tmap_a_half : &tmap_a;
                const uint32_t stage_bytes = (narrow ? S::A_BYTES / 2 : S::A_BYTES) + bn_cta * GEMM_BLOCK_K * 2;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait<kCtaGroup == 2>(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    uint8_t* sb = sa + S::A_BYTES;
                    if constexpr (kCtaGroup == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
                        if (p.hint_a) tma_load_2d_hint(sa, ta, &full_bar[stage], kb * GEMM_BLOCK_K, m0, p.hint_a);
                        else tma_load_2d(sa, ta, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
                        if (p.hint_b) tma_load_2d_hint(sb, tb, &full_bar[stage], kb * GEMM_BLOCK_K, n0, p.hint_b);
                        else tma_load_2d(sb, tb, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
                    } else {
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
                        if (p.hint_a) tma_load_2d_cg2_hint(sa, ta, &full_bar[stage], 0, kb * GEMM_BLOCK_K, m0, p.hint_a);
                        else tma_load_2d_cg2(sa, ta, &full_bar[stage], 0, kb * GEMM_BLOCK_K, m0);
                        if (p.hint_b) tma_load_2d_cg2_hint(sb, tb, &full_bar[stage], 0, kb * GEMM_BLOCK_K, n0, p.hint_b);
                        else tma_load_2d_cg2(sb, tb, &full_bar[stage], 0, kb * GEMM_BLOCK_K, n0);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp_idx == 1) {
        // ================================ MMA issuer ================================
        if (is_leader && elect_one()) {
            constexpr uint32_t idesc_full = make_idesc_bf16(TILE_M, BLOCK_N, 0, 0);
            constexpr uint32_t idesc_half = make_idesc_bf16(TILE_M, BLOCK_N / 2, 0, 0);
            constexpr uint32_t idesc_narrow = make_idesc_bf16(TILE_M / 2, BLOCK_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit = cluster_id; unit < p.num_units; unit += num_clusters, ++it) {
                uint32_t idesc = (unit < p.full_units) ? idesc_full : idesc_half;
                if (p.narrow_m && unit < p.full_units) {
                    int m_blk, n_blk;
                    tile_coords(unit, m_blk, n_blk);
                    if (is_narrow(m_blk)) idesc = idesc_narrow;
                }
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait<kCtaGroup == 2>(&tmem_empty_bar[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait<kCtaGroup == 2>(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                        // +32 bytes (= 16 bf16) along K inside the 128B swizzle atom
                        umma_ss<kCtaGroup>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (kCtaGroup == 1) umma_commit(&empty_bar[stage]);
                    else umma_commit_cg2(&empty_bar[stage], 0x3);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if constexpr (kCtaGroup == 1) umma_commit(&tmem_full_bar[acc]);
                else umma_commit_cg2(&tmem_full_bar[acc], 0x3);
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue ================================
        const int q = warp_idx & 3;  // TMEM lane quarter this warp may access
        const int lane = (int)lane_id();
        int it = 0;
        for (int unit = cluster_id; unit < p.num_units; unit += num_clusters, ++it) {
            int tile, half;
            unit_tile(unit, tile, half);
            const int bn_eff = (half < 0) ? BLOCK_N : BLOCK_N / 2;     // accumulator columns of this unit
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            // narrow tile (layout B): this CTA holds 64 rows; warps 0/1 own accumulator columns [0, N/2), warps 2/3 the
            // columns [N/2, N), all at TMEM columns [0, N/2) of their own lanes
            const bool narrow = is_narrow(m_blk);
            const int m0 = m_blk * TILE_M + (int)cta_rank * (narrow ? 64 : 128);
            const int n_tile0 = n_blk * BLOCK_N + (half > 0 ? BLOCK_N / 2 : 0);
            const int col_half = narrow ? (q >> 1) : 0;
            const int n0 = n_tile0 + col_half * (BLOCK_N / 2);          // first output column this thread handles
            const int n_cols = narrow ? BLOCK_N / 2 : bn_eff;            // ... and how many (TMEM columns [0, n_cols))
            const int row = m0 + (narrow ? (q & 1) : q) * 32 + lane;
            const bool row_ok = row < p.M;
            // EPI_GATED_RES: the residual row slice, the per-sample gate and the bias do not depend on the accumulator;
            // chunk c+1's loads are issued before chunk c is combined, and the first chunk's loads are issued before
            // the wait for the accumulator itself, so their latency hides behind the mainloop / the previous chunk.
            uint4 res_n[4], gate_n[4], bias_n[4];
            const __nv_bfloat16* gate_row = nullptr;
            const __nv_bfloat16* res_row = nullptr;
            auto prefetch = [&](int col) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    res_n[j] = (row_ok && !(p.debug & 1)) ? __ldcg(reinterpret_cast<const uint4*>(res_row + col) + j) : make_uint4(0, 0, 0, 0);
                    gate_n[j] = __ldg(reinterpret_cast<const uint4*>(gate_row + col) + j);
                    bias_n[j] = p.bias != nullptr ? __ldg(reinterpret_cast<const uint4*>(p.bias + col) + j)
                                                  : make_uint4(0, 0, 0, 0);
                }
            };
            if constexpr (kEpi == EPI_GATED_RES) {
                gate_row = p.gate + (long long)((row_ok ? row : 0) / p.rows_per_sample) * p.ld_gate;
                res_row = p.resid + (long long)(row_ok ? row : 0) * p.ldr;
                prefetch(n0);
            }
            mbar_wait<kCtaGroup == 2>(&tmem_full_bar[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;

            if constexpr (kEpi == EPI_GATED_RES) {
                // x' = x + bf16(bf16(acc + bias) * gate); operands prefetched above / one chunk ahead
#pragma unroll 1
                for (int c = 0; c < n_cols / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_x32(taddr + c * 32, r);
                    uint32_t res_p[16], gate_p[16], bias_p[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        res_p[4 * j] = res_n[j].x; res_p[4 * j + 1] = res_n[j].y; res_p[4 * j + 2] = res_n[j].z; res_p[4 * j + 3] = res_n[j].w;
                        gate_p[4 * j] = gate_n[j].x; gate_p[4 * j + 1] = gate_n[j].y; gate_p[4 * j + 2] = gate_n[j].z; gate_p[4 * j + 3] = gate_n[j].w;
                        bias_p[4 * j] = bias_n[j].x; bias_p[4 * j + 1] = bias_n[j].y; bias_p[4 * j + 2] = bias_n[j].z; bias_p[4 * j + 3] = bias_n[j].w;
                    }
                    const int col = n0 + c * 32;
                    if (c + 1 < n_cols / 32) prefetch(col + 32);
                    tmem_ld_wait();
                    uint32_t outp[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // reference rounding points: y = bf16(acc+b); yg = bf16(y*gate); x' = bf16(x+yg)
                        float a = bf16_round(__uint_as_float(r[2 * j]) + bf16_lo(bias_p[j]));
                        float b = bf16_round(__uint_as_float(r[2 * j + 1]) + bf16_hi(bias_p[j]));
                        a = bf16_round(a * bf16_lo(gate_p[j]));
                        b = bf16_round(b * bf16_hi(gate_p[j]));
                        outp[j] = pack_bf16x2(bf16_lo(res_p[j]) + a, bf16_hi(res_p[j]) + b);
                    }
                    if (row_ok && !((p.debug & 1) && outp[0] != 0x12345678u)) {
                        uint4* cp = reinterpret_cast<uint4*>(p.C + (long long)row * p.ldc + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            cp[j] = make_uint4(outp[4 * j], outp[4 * j + 1], outp[4 * j + 2], outp[4 * j + 3]);
                    }
                }
            } else if constexpr (kEpi == EPI_STORE) {
#pragma unroll 1
                for (int c = 0; c < n_cols / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_x32(taddr + c * 32, r);
                    tmem_ld_wait();
                    const int col = n0 + c * 32;
                    uint32_t outp[16];
                    uint32_t bias_p[16];
                    if (p.bias != nullptr) {
                        const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 v = __ldg(bp + j);
                            bias_p[4 * j] = v.x; bias_p[4 * j + 1] = v.y; bias_p[4 * j + 2] = v.z; bias_p[4 * j + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) bias_p[j] = 0;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float a = bf16_round(__uint_as_float(r[2 * j]) + bf16_lo(bias_p[j]));
                        float b = bf16_round(__uint_as_float(r[2 * j + 1]) + bf16_hi(bias_p[j]));
                        if (p.act == 1) { a = silu_f(a); b = silu_f(b); }
                        outp[j] = pack_bf16x2(a, b);
                    }
                    if (row_ok) {
                        uint4* cp = reinterpret_cast<uint4*>(p.C + (long long)row * p.ldc + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            cp[j] = make_uint4(outp[4 * j], outp[4 * j + 1], outp[4 * j + 2], outp[4 * j + 3]);
                    }
                }
            } else if constexpr (kEpi == EPI_SWIGLU) {
                // weight rows are interleaved in groups of 64: [gate 64 | up 64] per 128 accumulator columns
                static_assert(kEpi != EPI_SWIGLU || BLOCK_N % 128 == 0, "SwiGLU epilogue needs 128-column groups");
#pragma unroll 1
                for (int c = 0; c < n_cols / 64; ++c) {
                    const int grp = c >> 1, hf = c & 1;
                    uint32_t g[32], u[32];
                    tmem_ld_x32(taddr + grp * 128 + hf * 32, g);
                    tmem_ld_x32(taddr + grp * 128 + 64 + hf * 32, u);
                    tmem_ld_wait();
                    uint32_t outp[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // liger swiglu: silu(bf16(gate).fp32).bf16 * bf16(up)
                        float g0 = bf16_round(__uint_as_float(g[2 * j])), g1 = bf16_round(__uint_as_float(g[2 * j + 1]));
                        float u0 = bf16_round(__uint_as_float(u[2 * j])), u1 = bf16_round(__uint_as_float(u[2 * j + 1]));
                        g0 = bf16_round(silu_f(g0));
                        g1 = bf16_round(silu_f(g1));
                        outp[j] = pack_bf16x2(g0 * u0, g1 * u1);
                    }
                    if (row_ok) {
                        const int col = n0 / 2 + grp * 64 + hf * 32;   // n0 already includes a narrow tile's column half
                        uint4* cp = reinterpret_cast<uint4*>(p.C + (long long)row * p.ldc + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            cp[j] = make_uint4(outp[4 * j], outp[4 * j + 1], outp[4 * j + 2], outp[4 * j + 3]);
                    }
                }
            } else if constexpr (kEpi == EPI_QKV_ROPE) {
                // One 256-column tile = one head of q, k or v; one ROW is handled by two threads (one per epilogue
                // warpgroup, same TMEM lane): warpgroup h owns columns [64h, 64h+64) and their RoPE partners
                // [128+64h, 128+64h+64), so the rotation pairs (j, j+128) stay thread-local and only the two partial sums
                // of squares are exchanged through shared memory.  The rotated bf16 values stay in registers (64 packed
                // words); TMEM is released to the MMA warp as soon as it has been read, before normalisation and stores.
                static_assert(kEpi != EPI_QKV_ROPE || BLOCK_N == 256, "QKV epilogue needs one head per tile");
                const int hw = (warp_idx - 2) >> 2;           // column half owned by this warpgroup
                const bool do_norm = n0 < p.qk_cols;
                const bool do_rope = do_norm && p.rope_cos != nullptr;
                const int pos = (row_ok ? row : 0) % p.rows_per_sample;
                const uint4* cosr = reinterpret_cast<const uint4*>(p.rope_cos + (long long)pos * 128);
                const uint4* sinr = reinterpret_cast<const uint4*>(p.rope_sin + (long long)pos * 128);
                uint32_t keep[64];   // [0,32): columns 64h..64h+63, [32,64): columns 128+64h.. (bf16 pairs)
                float ssq = 0.f;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = hw * 8 + cc;
                    uint32_t x1[8], x2[8];
                    tmem_ld_x8(taddr + c * 8, x1);
                    tmem_ld_x8(taddr + 128 + c * 8, x2);
                    tmem_ld_wait();
                    uint4 b1 = make_uint4(0, 0, 0, 0), b2 = b1, cs = b1, sn = b1;
                    if (p.bias != nullptr) {
                        b1 = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + c * 8));
                        b2 = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + 128 + c * 8));
                    }
                    if (do_rope) {
                        cs = __ldg(cosr + c);
                        sn = __ldg(sinr + c);
                    }
                    const uint32_t b1w[4] = {b1.x, b1.y, b1.z, b1.w}, b2w[4] = {b2.x, b2.y, b2.z, b2.w};
                    const uint32_t csw[4] = {cs.x, cs.y, cs.z, cs.w}, snw[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // reference rounding points: qkv = bf16(acc + bias); rope in fp32 -> bf16; norm in fp32 -> bf16
                        float a0 = bf16_round(__uint_as_float(x1[2 * j]) + bf16_lo(b1w[j]));
                        float a1 = bf16_round(__uint_as_float(x1[2 * j + 1]) + bf16_hi(b1w[j]));
                        float e0 = bf16_round(__uint_as_float(x2[2 * j]) + bf16_lo(b2w[j]));
                        float e1 = bf16_round(__uint_as_float(x2[2 * j + 1]) + bf16_hi(b2w[j]));
                        if (do_rope) {
                            // model.py:412-413: y1 = x1*cos + x2*sin ; y2 = x1*(-sin) + x2*cos
                            const float cl = bf16_lo(csw[j]), ch = bf16_hi(csw[j]), sl = bf16_lo(snw[j]), sh = bf16_hi(snw[j]);
                            const float y10 = bf16_round(a0 * cl + e0 * sl), y20 = bf16_round(a0 * (-sl) + e0 * cl);
                            const float y11 = bf16_round(a1 * ch + e1 * sh), y21 = bf16_round(a1 * (-sh) + e1 * ch);
                            a0 = y10; e0 = y20; a1 = y11; e1 = y21;
                        }
                        ssq += a0 * a0 + a1 * a1 + e0 * e0 + e1 * e1;
                        keep[cc * 4 + j] = pack_bf16x2(a0, a1);
                        keep[32 + cc * 4 + j] = pack_bf16x2(e0, e1);
                    }
                }
                // all TMEM reads of this accumulator are done: hand it back to the MMA warp now
                tc_fence_before();
                __syncwarp();
                if (elect_one()) {
                    if constexpr (kCtaGroup == 1) mbar_arrive(&tmem_empty_bar[acc]);
                    else mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
                }
                __syncwarp();
                // full-row sum of squares: exchange the two halves (double-buffered by tile parity; the two warps that
                // share TMEM lane quarter q meet on named barrier 1 + q).  NOTE the summation order (own half + other
                // half) differs between the two threads only by commutation, so both compute the same rstd bits.
                {
                    float* xch = reinterpret_cast<float*>(smem + S::XCH_OFFSET) + (it & 1) * 256;
                    const int r_in_tile = q * 32 + lane;
                    xch[hw * 128 + r_in_tile] = ssq;
                    named_bar_sync(1 + q, 64);
                    const float other = xch[(hw ^ 1) * 128 + r_in_tile];
                    ssq = (hw == 0) ? ssq + other : other + ssq;
                }
                const float rstd = do_norm ? rsqrtf(ssq * (1.0f / 256.0f) + p.eps) : 1.0f;
                long long out_off = (long long)row * p.ldc + n0;
                __nv_bfloat16* out_base = p.C;
                if (p.sp_ranks > 0) {
                    const int t = n0 >> 8, which = t / p.n_heads, head = t % p.n_heads;
                    const int rr = row_ok ? row : 0;
                    const int smp = rr / p.rows_per_sample, li = rr % p.rows_per_sample;
                    const int dst = head / p.sp_hp;
                    const long long col = (long long)which * p.sp_hp * 256 + (head % p.sp_hp) * 256;
                    if (p.sp_peer[0] != nullptr) {
                        const long long drow = (long long)smp * p.sp_seq + (long long)p.sp_rank * p.rows_per_sample + li;
                        out_off = drow * p.ldc + col;
                        out_base = p.sp_peer[dst];
                    } else {
                        const long long drow = ((long long)smp * p.sp_ranks + dst) * p.rows_per_sample + li;
                        out_off = drow * p.ldc + col;
                    }
                }
                if (p.stage_stores) {
                    // warp-local transpose through shared memory: lane = row on the way in, lane = 16-byte chunk of a
                    // row segment on the way out (four rows x 128 contiguous bytes per store instruction); chunks are
                    // XOR-swizzled by the row.  Two passes: columns [64h, 64h+64), then [128+64h, 128+64h+64).
                    uint8_t* stg = smem + S::STAGING_OFFSET + (warp_idx - 2) * (32 * 128);
                    const unsigned long long my_ptr = row_ok ? (unsigned long long)(out_base + out_off) : 0ull;
#pragma unroll
                    for (int seg = 0; seg < 2; ++seg) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            uint32_t o[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t v = keep[seg * 32 + 4 * u + j];
                                o[j] = do_norm ? pack_bf16x2(bf16_lo(v) * rstd, bf16_hi(v) * rstd) : v;
                            }
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((u ^ (lane & 7)) << 4)) =
                                make_uint4(o[0], o[1], o[2], o[3]);
                        }
                        __syncwarp();
                        const int c = lane & 7;
#pragma unroll 4
                        for (int i2 = 0; i2 < 8; ++i2) {
                            const int r = 4 * i2 + (lane >> 3);
                            const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
                            const unsigned long long rp = __shfl_sync(0xffffffffu, my_ptr, r);
                            if (rp != 0ull) reinterpret_cast<uint4*>(rp)[seg * 16 + hw * 8 + c] = v;
                        }
                        __syncwarp();
                    }
                } else if (row_ok) {
                    uint4* cp = reinterpret_cast<uint4*>(out_base + out_off);
#pragma unroll
                    for (int seg = 0; seg < 2; ++seg) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            uint32_t o[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t v = keep[seg * 32 + 4 * u + j];
                                o[j] = do_norm ? pack_bf16x2(bf16_lo(v) * rstd, bf16_hi(v) * rstd) : v;
                            }
                            cp[seg * 16 + hw * 8 + u] = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
                continue;   // barrier already signalled
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                if constexpr (kCtaGroup == 1) mbar_arrive(&tmem_empty_bar[acc]);
                else mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            }
            __syncwarp();
        }
    }

    // ================================ teardown ================================
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp_idx == 1) tmem_dealloc<kCtaGroup>(tmem_base, TMEM_COLS);
}

}  // namespace flite
