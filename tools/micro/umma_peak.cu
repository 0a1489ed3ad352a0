// umma_peak.cu — full-chip INT8 tensor-core peak of this B200: every SM runs one persistent CTA (or CTA pair) that issues
// tcgen05.mma.kind::i8 (M = 128 per CTA, N = 256, K = 32) back to back on shared-memory operands that never change, accumulators
// in TMEM.  No global traffic, no epilogue: this is the ceiling the convolution kernels' tensor-bound layers are judged
// against ("the builder must measure the INT8 tcgen05 peak on the box").
//
//   mode 1: cta_group::1, one CTA per SM, M128 N256 K32 per instruction
//   mode 2: cta_group::2, clusters of two CTAs (M256 N256 K32 per instruction, each CTA holds half of B)
//
// Output: one JSON object per (mode, duration) on stdout.  TOPS = 2 * M * N * K * instructions / time (CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../yolo-compression-and-deployment-in-fpga_b200/csrc -o umma_peak umma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int BATCH = 64;       // MMAs per commit

// A: 128 rows x 128 B (K = 128: four K = 32 steps), B: NB rows x 128 B, both canonical K-major SWIZZLE_128B tiles.
template <int CTAS>
__global__ void __launch_bounds__(128, 1) peak_kernel(long long batches, unsigned long long *cycles)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *bp = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    constexpr int NB = CTAS == 2 ? 128 : 256;                 // rows of B held by this CTA
    for (int i = threadIdx.x; i < (128 * 128 + NB * 128) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(bp)[i] = (uint32_t)(i + 977 * blockIdx.x) * 2654435761u ^ 0x5bd1e995u * (uint32_t)(i >> 3);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_barrier_init(); }
    if (warp == 0) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    }
    fence_proxy_async();
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const bool leader = CTAS == 1 || cluster_ctarank() == 0;
    if (warp == 0 && leader) {
        // idesc: D s32 (2 << 4), A s8 (1 << 7), B s8 (1 << 10), N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t M = CTAS == 2 ? 256u : 128u;
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((M >> 4) << 24);
        const uint32_t a0 = base, b0 = base + 128 * 128;
        uint32_t ph[2] = {0, 0};
        const long long t0 = clock64();
        for (long long b = 0; b < batches; ++b) {
            const int s = (int)(b & 1);
            if (b >= 2) { mbar_wait(smem_u32(&bar[s]), ph[s]); ph[s] ^= 1u; }      // at most two batches in flight
            if (elect_one()) {
#pragma unroll 8
                for (int m = 0; m < BATCH; ++m) {
                    const uint32_t k = (uint32_t)(m & 3) * 32u;
                    const uint64_t ad = make_desc(a0 + k, 16, 1024, 2), bd = make_desc(b0 + k, 16, 1024, 2);
                    const uint32_t d = tm + (uint32_t)((m >> 2) & 1) * 256u;
                    if (CTAS == 2) umma_i8_2cta(d, ad, bd, idesc, 1u); else umma_i8(d, ad, bd, idesc, 1u);
                }
                if (CTAS == 2) umma_commit_2cta(smem_u32(&bar[s]), 1); else umma_commit(smem_u32(&bar[s]));
            }
            __syncwarp();
        }
        for (long long b = batches > 2 ? batches - 2 : 0; b < batches; ++b) { const int s = (int)(b & 1); mbar_wait(smem_u32(&bar[s]), ph[s]); ph[s] ^= 1u; }
        if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
        else tmem_dealloc(tm, 512);
    }
}

template <int CTAS>
static double run(int grid, long long batches, unsigned long long *d_cyc, double *cyc_per_mma)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int smem = 128 * 128 + 256 * 128 + 2048;
    cudaFuncSetAttribute(peak_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaMemset(d_cyc, 0, 148 * 8);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, peak_kernel<CTAS>, batches, d_cyc);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "mode %d: %s\n", CTAS, cudaGetErrorString(e)); exit(1); }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    *cyc_per_mma = (double)mx / (double)(batches * BATCH);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

int main(int argc, char **argv)
{
    const double target_s = argc > 1 ? atof(argv[1]) : 1.0;       // length of the sustained run
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int grid = prop.multiProcessorCount & ~1;
    unsigned long long *d; cudaMalloc(&d, 148 * 8);
    for (int mode = 1; mode <= 2; ++mode) {
        double cpm;
        // warm-up + calibration of the batch count
        double ms = mode == 1 ? run<1>(grid, 2000, d, &cpm) : run<2>(grid, 2000, d, &cpm);
        const double per_batch_ms = ms / 2000.0;
        struct { const char *name; double secs; } legs[] = { {"burst", 0.004}, {"sustained", target_s} };
        for (auto &lg : legs) {
            long long batches = (long long)(lg.secs * 1000.0 / per_batch_ms);
            if (batches < 16) batches = 16;
            double best = 1e30, cp = 0;
            const int reps = lg.secs < 0.1 ? 10 : 1;
            for (int r = 0; r < reps; ++r) { double c1; double t = mode == 1 ? run<1>(grid, batches, d, &c1) : run<2>(grid, batches, d, &c1); if (t < best) { best = t; cp = c1; } }
            // instructions: every CTA (mode 1) or every leader CTA (mode 2) issues batches * BATCH MMAs of M x 256 x 32
            const double mmas = (double)batches * BATCH * (mode == 1 ? grid : grid / 2);
            const double ops = 2.0 * (mode == 1 ? 128.0 : 256.0) * 256.0 * 32.0 * mmas;
            printf("{\"mode\": \"cta_group::%d\", \"leg\": \"%s\", \"ctas\": %d, \"ms\": %.4f, \"tops\": %.1f, \"cycles_per_mma\": %.2f, \"implied_sm_mhz\": %.0f}\n",
                   mode, lg.name, grid, best, ops / (best * 1e-3) / 1e12, cp, cp * (double)batches * BATCH / (best * 1e-3) / 1e6);
            fflush(stdout);
        }
    }
    return 0;
}
