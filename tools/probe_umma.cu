// Probe (run on the B200 through gpurun): TMEM accumulator layout and issue-to-completion time of tcgen05.mma
// cta_group::2 for M = 256 vs M = 128 (64 rows per CTA) and narrow N -- the facts the GEMM / attention tail units rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_umma.bin tools/probe_umma.cu
// Operands are written to shared memory by hand in the 128B-swizzled K-major layout TMA would produce:
//   A[m][0] = m + 1, A[m][1] = 1, B[n][0] = 1, B[n][1] = 512 n  =>  D[m][n] = m + 1 + 512 n  (decodes to (m, n)).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../f-lite_b200/csrc/common.cuh"

using namespace flite;

__device__ void fill_sw128(uint8_t* tile, int rows, int row0_global, bool is_a) {
    // tile: [rows][64 bf16] K-major, 128 B per row, 16-byte chunk c of row r stored at chunk (c ^ (r & 7))
    for (int i = threadIdx.x; i < rows * 64; i += blockDim.x) {
        const int r = i >> 6, k = i & 63;
        float v = 0.f;
        const int g = row0_global + r;
        if (is_a) v = (k == 0) ? (float)(g + 1) : (k == 1 ? 1.f : 0.f);
        else v = (k == 0) ? 1.f : (k == 1 ? 512.f * g : 0.f);
        const int chunk = k >> 3, within = k & 7;
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(tile + r * 128 + ((chunk ^ (r & 7)) << 4)) + within;
        *p = __float2bfloat16(v);
    }
}

// mode: M of the 2-CTA instruction (256 or 128); N: instruction N.  out[cta][128 lanes][256 cols] fp32.
__global__ void __launch_bounds__(128, 1) probe_kernel(float* out, long long* cycles, int M, int N, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                 // up to 128 rows
    uint8_t* sb = smem + 16384;         // up to 128 rows (N/2 per CTA)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    const int rows_a = M / 2, rows_b = N / 2;
    fill_sw128(sa, rows_a, (int)rank * rows_a, true);
    fill_sw128(sb, rows_b, (int)rank * rows_b, false);
    if (warp == 0) {
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
    // zero the accumulator region first so untouched lanes/columns read back as a sentinel
    {
        uint32_t z[32];
        for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-7.0f);
        for (int c = 0; c < 8; ++c) tmem_st_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, z);
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
        const uint64_t da = make_smem_desc_sw128(smem_u32(sa), 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sb), 16, 1024);
        t0 = clock64();
        for (int i = 0; i < reps; ++i) umma_ss<2>(tmem_base, da, db, idesc, 0u);
        umma_commit_cg2(bar, 0x3);
    }
    mbar_wait<true>(bar, 0, 1);
    if (rank == 0 && threadIdx.x == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc_fence_after();
    for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i)
            out[((size_t)rank * 128 + threadIdx.x) * 256 + c * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc<2>(tmem_base, 512);
}

static void run(int M, int N, int reps, bool dump) {
    float* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, 2 * 128 * 256 * sizeof(float));
    cudaMalloc(&d_cyc, sizeof(long long));
    cudaMemset(d_out, 0, 2 * 128 * 256 * sizeof(float));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 40960;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, d_out, d_cyc, M, N, reps);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        printf("M=%d N=%d: launch %s / sync %s\n", M, N, cudaGetErrorString(e), cudaGetErrorString(e2));
        exit(1);
    }
    std::vector<float> h(2 * 128 * 256);
    long long cyc = 0;
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("TIMING M=%d N=%d reps=%d cycles=%lld per_mma=%.1f\n", M, N, reps, cyc, (double)cyc / reps);
    if (!dump) return;
    // decode: value v = m + 1 + 512 n
    for (int cta = 0; cta < 2; ++cta) {
        printf("LAYOUT M=%d N=%d cta=%d: lane -> (row m, col-range n) per 32-column group\n", M, N, cta);
        for (int lane = 0; lane < 128; lane += 1) {
            if (!(lane % 16 == 0 || lane % 16 == 15)) continue;
            printf("  lane %3d:", lane);
            for (int c = 0; c < 256; c += 32) {
                const float v0 = h[((size_t)cta * 128 + lane) * 256 + c], v1 = h[((size_t)cta * 128 + lane) * 256 + c + 31];
                auto dec = [](float v, int& m, int& n) {
                    if (v < 0) { m = -1; n = -1; return; }
                    const long long iv = (long long)(v + 0.5f);
                    n = (int)(iv / 512); m = (int)(iv % 512) - 1;
                };
                int m0, n0, m1, n1;
                dec(v0, m0, n0); dec(v1, m1, n1);
                printf(" [c%3d: m%d n%d..m%d n%d]", c, m0, n0, m1, n1);
            }
            printf("\n");
        }
    }
    cudaFree(d_out);
    cudaFree(d_cyc);
}

int main() {
    run(256, 256, 1, true);
    run(128, 256, 1, true);
    run(128, 128, 1, true);
    const int reps = 4000;
    const int shapes[][2] = {{256, 256}, {128, 256}, {256, 128}, {128, 128}, {256, 64}, {256, 32}, {256, 16}, {128, 32}};
    for (auto& s : shapes) run(s[0], s[1], reps, false);
    return 0;
}
