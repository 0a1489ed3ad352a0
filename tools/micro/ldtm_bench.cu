// ldtm_bench.cu — tcgen05.ld throughput: how many bytes per cycle can the warps of one CTA read from TMEM?
// Each warp reads 32 columns x 32 lanes (4 KB) per instruction from its own lane quarter; W warps, R rounds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

template <int X>
__global__ void __launch_bounds__(1024, 1) bench(long long *out, int reps, int wait_every)
{
    __shared__ uint32_t slot;
    __shared__ long long t_begin[32], t_end[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    int acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (X == 32) {
            int v[32];
            tmem_ld32(tm + (uint32_t)((r * 32) & 511), v);
            if ((r + 1) % wait_every == 0) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
        } else {
            int v[16];
            tmem_ld16(tm + (uint32_t)((r * 16) & 511), v);
            if ((r + 1) % wait_every == 0) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= v[j];
        }
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    if (lane == 0) { t_begin[warp] = t0; t_end[warp] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = t_begin[0], e = t_end[0];
        for (int w = 1; w < (int)blockDim.x / 32; ++w) { b = min(b, t_begin[w]); e = max(e, t_end[w]); }
        out[0] = e - b;
    }
    if (acc == 0x12345678) out[1] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main()
{
    long long *d; cudaMalloc(&d, 16);
    for (int x : {16, 32})
        for (int warps : {1, 4, 8, 16, 32})
            for (int we : {1, 4}) {
                const int reps = 256;
                if (x == 32) bench<32><<<1, warps * 32>>>(d, reps, we); else bench<16><<<1, warps * 32>>>(d, reps, we);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                const double bytes = (double)warps * reps * x * 32 * 4;
                printf("tcgen05.ld.32x32b.x%d  warps=%2d wait_every=%d: %7lld cycles, %6.1f B/clk per SM, %5.1f cycles per instruction per warp\n",
                       x, warps, we, h, bytes / h, (double)h / reps);
            }
    return 0;
}
