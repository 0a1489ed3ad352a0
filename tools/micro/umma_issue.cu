// umma_issue.cu — how fast can one warp ISSUE tcgen05.mma (M=128,K=32,N small)?  Variants of the issuing code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_i8_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

template <int V>
__global__ void __launch_bounds__(128, 1) bench(int N, int nmma, long long *out, int reps, uint32_t step_a, uint32_t step_b)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad0 = make_desc(base, 2960, 160, 0), bd0 = make_desc(base + 96 * 1024, 128, 2304, 0);
    if (warp == 0) {
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64();
            if (V == 0) {            // one lane, rolled loop
                if (lane == 0) {
                    uint64_t ad = ad0, bd = bd0;
                    for (int m = 0; m < nmma; ++m) { umma_i8_acc(tm, ad, bd, idesc); ad += step_a; bd += step_b; }
                    umma_commit(smem_u32(&bar));
                }
            } else if (V == 1) {     // one lane, unrolled by 8
                if (lane == 0) {
                    uint64_t ad = ad0, bd = bd0;
#pragma unroll 8
                    for (int m = 0; m < nmma; ++m) { umma_i8_acc(tm, ad, bd, idesc); ad += step_a; bd += step_b; }
                    umma_commit(smem_u32(&bar));
                }
            } else if (V == 2) {     // whole warp runs the loop, one elected lane issues
                uint64_t ad = ad0, bd = bd0;
#pragma unroll 8
                for (int m = 0; m < nmma; ++m) { if (elect_one()) umma_i8_acc(tm, ad, bd, idesc); ad += step_a; bd += step_b; }
                if (elect_one()) umma_commit(smem_u32(&bar));
            } else {                 // whole warp, elect once outside, unrolled
                if (elect_one()) {
                    uint64_t ad = ad0, bd = bd0;
#pragma unroll 8
                    for (int m = 0; m < nmma; ++m) { umma_i8_acc(tm, ad, bd, idesc); ad += step_a; bd += step_b; }
                    umma_commit(smem_u32(&bar));
                }
            }
            __syncwarp();
            mbar_wait(smem_u32(&bar), phase); phase ^= 1;
            long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (lane == 0) out[blockIdx.x] = best;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int V> void run(const char *name, long long *d)
{
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {16, 32, 64, 128, 256})
        for (int nmma : {8, 64}) {
            bench<V><<<1, 128, 200 * 1024>>>(N, nmma, d, 20, 1, 16);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d nmma=%2d: total %6lld cycles, %6.1f per MMA (floor %d)\n", name, N, nmma, h, (double)h / nmma, 128 * N / 256);
        }
}

int main()
{
    long long *d; cudaMalloc(&d, 8);
    run<0>("lane0 rolled   ", d);
    run<1>("lane0 unroll8  ", d);
    run<2>("warp elect/iter", d);
    run<3>("warp elect once", d);
    return 0;
}
