// umma_swz_shift.cu — correctness probe for the next conv_ws step: can the 9 taps of a 3x3 convolution be
// row-shifted descriptor starts into ONE swizzled, pixel-major halo tile [pixel][Cin bytes] (the layout a TMA box with
// SWIZZLE_32B/64B/128B writes)?  The tile is written here the way TMA would (swizzle = XOR of address bits [4,7) with bits
// [7,10), masked by the swizzle span, applied to the byte offset from a 1024-byte-aligned base); A descriptor start =
// base + shift * Cin + kstep * 32, SBO = 10 * Cin (next tile row of a 10-pixel-wide halo); B = 32 x 32 identity, so that
// D[m][n] must equal tile[shift + 10 * (m / 8) + m % 8][32 * kstep + n].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../yolo-compression-and-deployment-in-fpga_b200/csrc -o umma_swz_shift umma_swz_shift.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace yb;

__host__ __device__ inline int8_t pat(int pixel, int k) { return (int8_t)(((pixel * 7 + k * 3) % 251) - 125); }

__global__ void __launch_bounds__(128, 1) probe(int cin, uint32_t layout, int shift, int kstep, int use_base_offset, int *out)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *bp = smem_raw + (base - smem_u32(smem_raw));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t span_mask = cin == 128 ? 7u : cin == 64 ? 3u : 1u;      // 16-byte chunks XORed: 8 / 4 / 2 per row
    const int npix = 220;
    for (int i = threadIdx.x; i < npix * cin; i += blockDim.x) {
        const int p = i / cin, k = i % cin;
        const uint32_t off = (uint32_t)i;
        const uint32_t sw = off ^ ((((off >> 7) & span_mask)) << 4);
        bp[sw] = (uint8_t)pat(p, k);
    }
    // B: 32 x 32 identity, no swizzle, [n/8][kchunk 2][8][16 B]
    uint8_t *bb = bp + 64 * 1024;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) bb[i] = 0;
    __syncthreads();
    if (threadIdx.x < 32) { const int n = threadIdx.x; bb[((n / 8) * 2 + (n / 16)) * 128 + (n % 8) * 16 + (n % 16)] = 1; }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 32); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t start = base + (uint32_t)(shift * cin + kstep * 32);
        uint64_t ad = make_desc(start, 16, (uint32_t)(10 * cin), layout);
        // descriptor bits [49,52): "matrix base offset" = (start >> 7) & 7 when the start is not aligned to the swizzle repeat
        if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7u) << 49;
        const uint64_t bd = make_desc(base + 64 * 1024, 128, 256, 0);
        umma_i8(tm, ad, bd, idesc, 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    int v[16];
    for (int h = 0; h < 2; ++h) {
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 16 * h, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 32 + 16 * h + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}

int main()
{
    int *d; cudaMalloc(&d, 128 * 32 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    static int h[128 * 32];
    int bad_total = 0;
    for (int ubo = 0; ubo < 2; ++ubo)
    for (int cin : {32, 64, 128}) {
        const uint32_t layout = cin == 128 ? 2u : cin == 64 ? 4u : 6u;
        for (int shift : {0, 1, 2, 3, 5, 10, 11, 12, 20, 21, 22})
            for (int kstep = 0; kstep < cin / 32; ++kstep) {
                probe<<<1, 128, 100 * 1024>>>(cin, layout, shift, kstep, ubo, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("cin %d shift %d kstep %d: %s\n", cin, shift, kstep, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
                int bad = 0, first = -1;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 32; ++n) {
                        const int want = pat(shift + 10 * (m / 8) + m % 8, 32 * kstep + n);
                        if (h[m * 32 + n] != want) { if (first < 0) first = m * 32 + n; ++bad; }
                    }
                printf("base_offset %d cin %3d layout %u shift %2d kstep %d: %s", ubo, cin, layout, shift, kstep, bad ? "MISMATCH" : "ok");
                if (bad) printf(" (%d of 4096, first at m=%d n=%d: got %d want %d)", bad, first / 32, first % 32, h[first],
                                pat(shift + 10 * ((first / 32) / 8) + (first / 32) % 8, 32 * kstep + first % 32));
                printf("\n");
                bad_total += bad;
            }
    }
    printf("%s\n", bad_total ? "SOME CONFIGURATIONS MISMATCH" : "ALL OK");
    return 0;
}
