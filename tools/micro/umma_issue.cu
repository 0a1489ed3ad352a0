// umma_issue.cu — how fast can one warp ISSUE tcgen05.mma (M=128,K=32,N small)?  Variants of the issuing code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

__device__ __forceinline__ void umma_i8_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

template <int V>
__global__ void __launch_bounds__(128, 1) bench(int N, int nmma, long long *out, int reps, uint32_t step_a, uint32_t step_b, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_layout, uint32_t a_off)
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
    const uint64_t ad0 = make_desc(base + a_off, a_lbo, a_sbo, a_layout), bd0 = make_desc(base + 96 * 1024, 128, 2304, 0);
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

template <int V> void run(const char *name, long long *d, uint32_t a_lbo = 2960, uint32_t a_sbo = 160, uint32_t a_layout = 0, uint32_t a_off = 0, uint32_t step_a = 1)
{
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {32, 64, 128, 256})
        for (int nmma : {64}) {
            bench<V><<<1, 128, 200 * 1024>>>(N, nmma, d, 20, step_a, 16, a_lbo, a_sbo, a_layout, a_off);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d nmma=%2d: total %6lld cycles, %6.1f per MMA (floor %d)\n", name, N, nmma, h, (double)h / nmma, 128 * N / 256);
        }
}

int main()
{
    long long *d; cudaMalloc(&d, 8);
    run<2>("noswz sbo128 lbo16         ", d, 16, 128, 0, 0, 0);
    run<2>("noswz sbo128 lbo1664       ", d, 1664, 128, 0, 32, 0);
    run<2>("noswz sbo128 lbo16 +1696+16", d, 16, 128, 0, 1696 + 16, 0);
    run<2>("noswz sbo160 lbo16         ", d, 16, 160, 0, 0, 0);
    run<2>("noswz sbo160 lbo2960 step1 ", d, 2960, 160, 0, 0, 1);
    run<2>("noswz sbo160 lbo2960 step0 ", d, 2960, 160, 0, 0, 0);
    run<2>("noswz sbo128 lbo2048 step0 ", d, 2048, 128, 0, 0, 0);
    run<2>("noswz sbo128 lbo2048 +16B  ", d, 2048, 128, 0, 16, 0);
    run<2>("noswz sbo288 lbo9920 step1 ", d, 9920, 288, 0, 0, 1);
    run<2>("noswz sbo256 lbo9920 step1 ", d, 9920, 256, 0, 0, 1);
    run<2>("noswz sbo256 lbo9936 +16   ", d, 9936, 256, 0, 16, 0);
    run<2>("noswz sbo144 lbo2320       ", d, 2320, 144, 0, 0, 1);
    run<2>("swz128 canonical step0     ", d, 16, 1024, 2, 0, 0);
    run<2>("swz128 canonical step2     ", d, 16, 1024, 2, 0, 2);
    return 0;
}
