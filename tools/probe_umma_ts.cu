// Probe (run on the B200 through gpurun): issue-to-completion time of the attention kernel's two MMA shapes --
// S = Q K^T (cta_group::2, M 256, N 128, both operands K-major in shared memory) and O += P V (M 256, N 256, B = V MN-major
// in shared memory, A = P either from shared memory or from TMEM) -- alone and while four warps keep reading / writing
// other TMEM columns with tcgen05.ld / tcgen05.st the way the softmax warps do.  Values are irrelevant (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_umma_ts.bin tools/probe_umma_ts.cu
#include <cstdio>
#include <cstdlib>

#include "../f-lite_b200/csrc/common.cuh"

using namespace flite;

// mode 0: S-like SS (M256 N128 K-major x K-major)      1: PV-like SS (A K-major smem, B MN-major smem, N 256)
// mode 2: PV-like TS (A from TMEM, B MN-major smem)    3: GEMM-like SS (M256 N256 K-major x K-major)
// traffic: 0 none | 1 warps 0-3 loop {4 x tcgen05.ld.x32 of 128 S columns, 2 x tcgen05.st.x32} back to back
__global__ void __launch_bounds__(160, 1) probe_kernel(long long* cycles, int mode, int traffic, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // 64 KB: A (128 rows x 256 K) as 4 chunks of [128 x 64]
    uint8_t* sb = smem + 65536;          // 32 KB: B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 98304 + 64);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 98304 + 128);
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) *stop = 0;
    if (warp == 4) {
        if (elect_one()) {
            mbar_init(bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2>(tmem_ptr, 512);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 4) {
        if (rank == 0 && elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(256, 128, 0, 0), idesc_o = make_idesc_bf16(256, 256, 0, 1),
                           idesc_g = make_idesc_bf16(256, 256, 0, 0);
            const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const int k = i & 7;
                if (mode == 0)
                    umma_ss<2>(tmem_base + 256, make_smem_desc_sw128(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                               make_smem_desc_sw128(b0 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc_s, 1u);
                else if (mode == 1)
                    umma_ss<2>(tmem_base + 256, make_smem_desc_sw128(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                               make_smem_desc_sw128(b0 + k * 2048, 16384, 1024), idesc_o, 1u);
                else if (mode == 2)
                    umma_ts<2>(tmem_base + 256, tmem_base + 128 + k * 8, make_smem_desc_sw128(b0 + k * 2048, 16384, 1024),
                               idesc_o, 1u);
                else
                    umma_ss<2>(tmem_base + 256, make_smem_desc_sw128(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                               make_smem_desc_sw128(b0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), idesc_g, 1u);
            }
            umma_commit_cg2(bar, 0x3);
            mbar_wait<true>(bar, 0, 1);
            cycles[0] = clock64() - t0;
            *stop = 1;
            uint32_t remote = mapa_shared(smem_u32((const void*)stop), 1);
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1) : "memory");
        }
        __syncwarp();
    } else if (traffic) {
        // softmax-like TMEM traffic on columns [0, 128): read 128 columns, write 64
        const uint32_t t = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t acc = 0;
        long long n = 0;
        while (*stop == 0) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_x32(t + c * 32, r);
                tmem_ld_wait();
                acc += r[c];
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = acc + i;
                tmem_st_x32(t + c * 32, r);
            }
            tmem_st_wait();
            if (traffic == 2) __nanosleep(600);      // paced: roughly one S tile per 2048 cycles
            ++n;
        }
        if (threadIdx.x == 0 && rank == 0) cycles[1] = n;
        if (acc == 0x12345678u) cycles[2] = acc;
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 4) tmem_dealloc<2>(tmem_base, 512);
}

int main() {
    long long* d_cyc;
    cudaMalloc(&d_cyc, 4 * sizeof(long long));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110592);
    const char* names[] = {"S-like  SS M256 N128 (K-major x K-major)", "PV-like SS M256 N256 (A smem, B = V MN-major)",
                           "PV-like TS M256 N256 (A TMEM, B = V MN-major)", "GEMM    SS M256 N256 (K-major x K-major)"};
    const int reps = 4000;
    for (int traffic = 0; traffic < 3; ++traffic)
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(d_cyc, 0, 4 * sizeof(long long));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2);
            cfg.blockDim = dim3(160);
            cfg.dynamicSmemBytes = 110592;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d_cyc, mode, traffic, reps);
            cudaError_t e2 = cudaDeviceSynchronize();
            if (e != cudaSuccess || e2 != cudaSuccess) {
                printf("mode %d traffic %d: launch %s / sync %s\n", mode, traffic, cudaGetErrorString(e), cudaGetErrorString(e2));
                return 1;
            }
            long long h[4];
            cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("TIMING %-48s tmem_traffic=%s cycles_per_mma=%.1f  (softmax-like iterations per warp during the run: %lld)\n",
                   names[mode], traffic == 0 ? "none  " : traffic == 1 ? "flood " : "paced ", (double)h[0] / reps, h[1]);
        }
    return 0;
}
