// ldtm_bench.cu — tcgen05.ld throughput: how many bytes per cycle can the warps of one CTA read from TMEM, by shape?
// Each warp reads from its own lane quarter; W warps, R rounds; the values are folded into one register so that the loads
// cannot be dropped.  Result on B200 (profiles/ldtm_bench_r2.txt).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

// SHAPE 0: 32x32b.x16 (16 regs), 1: 32x32b.x32, 2: 32x32b.x8, 3: 16x256b.x4 (16 regs), 4: 16x128b.x8 (16 regs), 5: 32x32b.x64, 6: 16x64b.x16
template <int SHAPE> struct Sh;
template <> struct Sh<0> { static constexpr int R = 16; static constexpr int BYTES = 32 * 16 * 4; static constexpr int COLS = 16; };
template <> struct Sh<1> { static constexpr int R = 32; static constexpr int BYTES = 32 * 32 * 4; static constexpr int COLS = 32; };
template <> struct Sh<2> { static constexpr int R = 8;  static constexpr int BYTES = 32 * 8 * 4;  static constexpr int COLS = 8; };
template <> struct Sh<3> { static constexpr int R = 16; static constexpr int BYTES = 32 * 16 * 4; static constexpr int COLS = 32; };   // 16 lanes x 32 columns
template <> struct Sh<4> { static constexpr int R = 16; static constexpr int BYTES = 32 * 16 * 4; static constexpr int COLS = 32; };   // 16 lanes x 32 columns
template <> struct Sh<5> { static constexpr int R = 64; static constexpr int BYTES = 32 * 64 * 4; static constexpr int COLS = 64; };

template <int SHAPE>
__device__ __forceinline__ void ld(uint32_t a, int *v)
{
    if (SHAPE == 0) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
    } else if (SHAPE == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
    } else if (SHAPE == 3) {
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
    } else if (SHAPE == 4) {
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a) : "memory");
    } else if (SHAPE == 1) {
        tmem_ld32(a, v);
    } else {
        tmem_ld32(a, v); tmem_ld32(a + 32, v + 32);      // (two .x32 back to back: is one wait per 64 columns any better?)
    }
}

template <int SHAPE>
__global__ void __launch_bounds__(1024, 1) bench(long long *out, int reps)
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
        int v[Sh<SHAPE>::R];
        ld<SHAPE>(tm + (uint32_t)((r * Sh<SHAPE>::COLS) & 511 & ~(Sh<SHAPE>::COLS - 1)), v);
        tmem_ld_wait();
        int x = 0;
#pragma unroll
        for (int j = 0; j < Sh<SHAPE>::R; j += 4) x ^= v[j] ^ v[j + 1] ^ v[j + 2] ^ v[j + 3];
        acc += x;
    }
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

template <int SHAPE> void run(const char *name, long long *d)
{
    for (int warps : {4, 8, 16, 32}) {
        const int reps = 512;
        bench<SHAPE><<<1, warps * 32>>>(d, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps=%2d: %7lld cycles, %6.1f B/clk per SM, %6.1f cycles per instruction per warp\n",
               name, warps, h, (double)warps * reps * Sh<SHAPE>::BYTES / h, (double)h / reps);
    }
}

int main()
{
    long long *d; cudaMalloc(&d, 16);
    run<2>("tcgen05.ld.32x32b.x8", d);
    run<0>("tcgen05.ld.32x32b.x16", d);
    run<1>("tcgen05.ld.32x32b.x32", d);
    run<5>("tcgen05.ld.32x32b.x32 twice", d);
    run<3>("tcgen05.ld.16x256b.x4", d);
    run<4>("tcgen05.ld.16x128b.x8", d);
    return 0;
}
