// Common sm_100a PTX wrappers for the F Lite B200 kernels: mbarrier, TMA, tcgen05 / TMEM.
// Everything here is inline PTX -- no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace flite {

#define FLITE_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// Watchdog: a barrier wait that exceeds ~2 s records who was waiting and makes every later
// wait fall through, so a protocol bug ends the kernel (with garbage output and an error code
// the host reports) instead of hanging the GPU.
// ------------------------------------------------------------------------------------------
__device__ unsigned int g_flite_abort = 0;   // 0 = ok, else (tag << 16 | blockIdx.x & 0xffff)

// Programmatic dependent launch (PDL): kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while the previous kernel of the stream is still draining; pdl_wait() blocks until that kernel has
// completed and its memory is visible (a no-op for a normal launch), pdl_launch_dependents() lets the NEXT kernel's
// CTAs be scheduled as SMs free up.  Rule used throughout: nothing produced by an earlier kernel is read (and nothing
// is written) before pdl_wait(); data rewritten inside the hot loop (residual stream) is read with ld.global.cg.
FLITE_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
FLITE_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// streaming 16-byte load: read-only path, do not keep the line in L1 (weights that are touched exactly once)
FLITE_DEVICE uint4 ld_nc_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

FLITE_DEVICE void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

FLITE_DEVICE uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

FLITE_DEVICE uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

FLITE_DEVICE uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %laneid;" : "=r"(l));
    return l;
}

FLITE_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

FLITE_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}

FLITE_DEVICE void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// map a shared::cta address of this CTA to the shared::cluster address of the same offset in CTA `rank`
FLITE_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
FLITE_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

FLITE_DEVICE void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

FLITE_DEVICE void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

FLITE_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
FLITE_DEVICE void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote = mapa_shared(smem_u32(bar), rank);
    // default semantics (release.cta), as CUTLASS's ClusterBarrier::arrive(cta_id): a .release.cluster here compiles to
    // MEMBAR.ALL.GPU + ERRBAR (measured: 21 % of the softmax warps' stall samples).  The data these arrivals publish
    // lives in TMEM / is fenced with fence.proxy.async, ordered by tcgen05.fence::before_thread_sync.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

FLITE_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

FLITE_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// cluster-scope acquire variant (needed when the arrival came from the peer CTA / its async proxy)
FLITE_DEVICE bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        // default (acquire.cta) like CUTLASS's ClusterBarrier::wait: acquire.cluster adds a CCTL.IVALL (L1 invalidate)
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

template <bool kCluster = false>
FLITE_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag) {
    if (kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t n = 0;
    while (true) {
        if (kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity)) return;
        if ((++n & 0x3ff) == 0) {
            if (*(volatile unsigned int*)&g_flite_abort != 0) return;
            if (globaltimer_ns() - t0 > 2000000000ull) {
                atomicCAS(&g_flite_abort, 0u, (tag << 16) | (blockIdx.x & 0xffffu) | 0x80000000u);
                return;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2D tiles
// ------------------------------------------------------------------------------------------
FLITE_DEVICE void tma_prefetch_desc(const CUtensorMap* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}

FLITE_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-CTA pair load: data lands in THIS CTA's smem, complete_tx is signalled on the barrier at the
// same offset in CTA `bar_rank` (the MMA leader).
FLITE_DEVICE void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, uint32_t bar_rank,
                                  int c0, int c1) {
    uint32_t bar_addr = mapa_shared(smem_u32(bar), bar_rank);
    uint32_t dst_addr = mapa_shared(smem_u32(smem_dst), cluster_ctarank());
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];"
        ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}

// Same loads with an L2 eviction-priority hint (createpolicy encodings as used by CUTLASS' TMA::CacheHintSm90):
// evict_first for an operand that streams past once, evict_last for the operand the rasterisation wants resident.
constexpr unsigned long long L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr unsigned long long L2_EVICT_LAST = 0x14F0000000000000ull;

FLITE_DEVICE void tma_load_2d_hint(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, int c0, int c1,
                                   unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
        "[%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
FLITE_DEVICE void tma_load_2d_cg2_hint(void* smem_dst, const CUtensorMap* desc, uint64_t* bar, uint32_t bar_rank,
                                       int c0, int c1, unsigned long long policy) {
    uint32_t bar_addr = mapa_shared(smem_u32(bar), bar_rank);
    uint32_t dst_addr = mapa_shared(smem_u32(smem_dst), cluster_ctarank());
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], "
        "[%1, {%3, %4}], [%2], %5;"
        ::"r"(dst_addr), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_addr), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

FLITE_DEVICE void tma_store_2d(const CUtensorMap* desc, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
FLITE_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
FLITE_DEVICE void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
FLITE_DEVICE void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, MMA, commit, TMEM load/store
// ------------------------------------------------------------------------------------------
template <int kCtaGroup>
FLITE_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    if constexpr (kCtaGroup == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                     "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                     "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}

template <int kCtaGroup>
FLITE_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

FLITE_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FLITE_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// SMEM matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46)
// | version=1 [46,48) | layout [61,64) (2 = SWIZZLE_128B).
FLITE_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32.
// a_major / b_major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_major, int b_major) {
    return (1u << 4)                      // c_format = F32
           | (1u << 7)                    // a_format = BF16
           | (1u << 10)                   // b_format = BF16
           | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

template <int kCtaGroup>
FLITE_DEVICE void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kCtaGroup == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// A operand from TMEM (bf16 packed two per 32-bit column, one row per lane), B from smem.
template <int kCtaGroup>
FLITE_DEVICE void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kCtaGroup == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}

// tcgen05.commit: the mbarrier gets one arrival when all prior MMAs of this thread have completed.
FLITE_DEVICE void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 2-CTA: arrive on the barrier at this offset in every CTA of `cta_mask`.
FLITE_DEVICE void umma_commit_cg2(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quarter) x N consecutive 32-bit columns, one row per thread.
FLITE_DEVICE void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
FLITE_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
FLITE_DEVICE void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
FLITE_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

FLITE_DEVICE void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
FLITE_DEVICE void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
FLITE_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// small numeric helpers
// ------------------------------------------------------------------------------------------
FLITE_DEVICE float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

FLITE_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
FLITE_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
FLITE_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

FLITE_DEVICE float silu_f(float x) { return x / (1.0f + __expf(-x)); }

}  // namespace flite
