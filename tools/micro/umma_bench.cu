// umma_bench.cu — microbenchmark: cycles per tcgen05.mma.kind::i8 (M=128, K=32) as a function of N and of the shared-memory
// layout of the A operand (swizzle-128B canonical, no-swizzle aligned, no-swizzle with 16-byte-shifted start / odd SBO).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../yolo-compression-and-deployment-in-fpga_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

struct Cfg { int N; uint32_t a_off, a_lbo, a_sbo, a_layout; uint32_t b_lbo, b_sbo, b_layout; int nmma; int vary; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long *out, int reps)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (base - smem_u32(smem_raw)))[i] = i * 2654435761u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = base, b0 = base + 96 * 1024;
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64();
            for (int m = 0; m < c.nmma; ++m) {
                // vary: rotate the A start among 9 tap-like offsets so consecutive MMAs do not hit identical addresses
                uint32_t off = c.a_off + (c.vary ? (uint32_t)((m % 3) * 16 + ((m / 3) % 3) * c.a_sbo) : 0u);
                uint64_t ad = make_desc(a0 + off, c.a_lbo, c.a_sbo, c.a_layout);
                uint64_t bd = make_desc(b0 + (uint32_t)(m % 8) * 256u, c.b_lbo, c.b_sbo, c.b_layout);
                umma_i8(tm, ad, bd, idesc, m > 0);
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), phase); phase ^= 1;
            long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        out[blockIdx.x] = best;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main()
{
    long long *d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Named { const char *name; Cfg c; };
    const int NM = 64;
    Named cfgs[] = {
        {"swz128 canonical        ", {0, 0, 16, 1024, 2, 16, 1024, 2, NM, 0}},
        {"noswz aligned sbo128    ", {0, 0, 2048, 128, 0, 128, 2304, 0, NM, 0}},
        {"noswz +16B   sbo128     ", {0, 16, 2048, 128, 0, 128, 2304, 0, NM, 0}},
        {"noswz +0     sbo160     ", {0, 0, 2960, 160, 0, 128, 2304, 0, NM, 0}},
        {"noswz taps   sbo160     ", {0, 0, 2960, 160, 0, 128, 2304, 0, NM, 1}},
        {"noswz taps   sbo288     ", {0, 0, 4944, 288, 0, 128, 2304, 0, NM, 1}},
        {"noswz taps   sbo256     ", {0, 0, 4624, 256, 0, 128, 2304, 0, NM, 1}},
        {"noswz lbo16  sbo160     ", {0, 0, 16, 160, 0, 128, 2304, 0, NM, 1}},
        {"swzA / noswzB           ", {0, 0, 16, 1024, 2, 128, 2304, 0, NM, 0}},
        {"noswzA taps160 / swzB   ", {0, 0, 2960, 160, 0, 16, 1024, 2, NM, 1}},
    };
    const int Ns[] = {32, 64, 128, 256};
    for (auto &nc : cfgs)
        for (int N : Ns) {
            Cfg c = nc.c; c.N = N;
            for (int grid : {1, 148}) {
                bench<<<grid, 128, 200 * 1024>>>(c, d, 20);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s N=%d: %s\n", nc.name, N, cudaGetErrorString(e)); return 1; }
                long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("%s N=%3d grid=%3d: %6.1f cycles/MMA (floor %d)\n", nc.name, N, grid, (double)mx / c.nmma, 128 * N / 256);
            }
        }
    return 0;
}
