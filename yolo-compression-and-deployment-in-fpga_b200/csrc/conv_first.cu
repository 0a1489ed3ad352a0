// conv_first.cu — the network's first layer: 3x3/stride-1/pad-1 convolution of the NHWC4 int8 input (R,G,B,0 per pixel)
// to <= 16 channels, with requantisation, leaky-ReLU and the optional 2x2 max-pool fused.
//
// Replaces first_conv (c_embedding/yolo_forward.c:269-418) and conv1 + tracker + pool (models/slim_yolo_v2.py:220-231).
// K = 27 is far too thin for the tcgen05 path (an M=128 MMA costs >= ~43 cycles whatever N and K are), and the layer is
// bound by instruction issue and shared-memory traffic, not by math.  It runs on warp-level integer MMAs
// (mma.sync.m16n8k32.s8) because their A fragment needs NO im2col staging at all: with K ordered (tap, channel) a thread's
// 32-bit A register is exactly one NHWC4 pixel word of the haloed input tile, i.e. one tap.  K = 36 = 9 taps x 4 bytes is
// covered by one k32 step (taps 0-7) and one k16 step (tap 8 + three zero-weight taps).
//
// Pooled variant: a warp owns 2 pooled rows x 8 pooled columns and runs four M=16 tiles, one per position (dy,dx) of the
// 2x2 window, so the four members of a pooled pixel are the same accumulator slot of the four tiles and the max is taken
// in registers (requantisation is monotone: max-then-requantise == requantise-then-max, slim_yolo_v2.py:229-231).
// Output channels are assigned to the two N=8 tiles so that thread t of a quad ends up with channels 4t..4t+3 of its
// pixel: one 32-bit store per pixel, 16 contiguous bytes per quad, 128 per warp row.
#include "kernels.h"
#include "epilogue.cuh"

namespace yb {

struct FirstParams {
    const int8_t *in;          // [n][H][W][4]  (RGB444 = false)
    const uint16_t *in16;      // [n][H][W] 0x0BGR camera pixels (RGB444 = true): quantised on the fly through the LUT
    const int *lut;            // 4096 packed (R,G,B,0) words = pixel_norm_quantize for every code (yolo_forward.c:57-85)
    const uint8_t *in8;        // [n][H][W][3] BGR bytes (SRC = 2): BaseTransform-without-resize + tracker quantiser through lut8
    const uint8_t *lut8;       // [3][256]: R (from byte 2), G, B
    int n_img, H, W;
    int OH, OW;
    int cs_out;                // 16
    const int8_t *wgt;         // [cout_pad][9][4]
    const int *bias_sh;
    LayerQ q;
    EpiConst k;
    int8_t *out;
    unsigned *ovf;
};

constexpr int F_TW = 32, F_TH = 16;                 // CTA tile, pre-pool pixels
constexpr int F_PITCH = 49;                         // words per halo row (>= F_TW + 2; 49 = 17 mod 32 keeps the tap rows on disjoint banks)
constexpr int F_HROWS = F_TH + 2;
constexpr int F_THREADS = 256;

__device__ __forceinline__ void mma_s8_k32(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8_k16(int (&c)[4], unsigned a0, unsigned a1, unsigned b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// POOL: tile = 32 x 16 pre-pool pixels -> 16 x 8 pooled; warp w = pooled rows 2*(w>>1), +1 and pooled columns 8*(w&1)..+7.
// !POOL: tile = 32 x 16 pixels; warp w = rows 2w, 2w+1 and four column groups of 8 (four M=16 tiles, rows g / g+8 = the two rows).
// SRC: 0 = int8 NHWC4; 1 = RGB444 camera frames: the RGB444 -> int8 quantiser (camera_to_inpBuf + pixel_norm_quantize,
// yolo_forward.c:57-123) is applied while the halo tile is staged, one lookup in the 4096-entry table per pixel;
// 2 = uint8 BGR image (three 256-entry tables, see quantize.cu).
// Persistent: a CTA loads the table(s), its B fragments and biases once and then walks tiles blockIdx.x, +gridDim.x, ...
// (per-tile work is then only the halo staging, the MMAs and the epilogue: ~300 instead of ~460 instructions per warp).
template <bool POOL, int EPI, bool ACT, int SRC>
__global__ void __launch_bounds__(F_THREADS) conv3x3_first_kernel(const FirstParams p)
{
    constexpr bool RGB444 = SRC == 1, U8 = SRC == 2;
    __shared__ unsigned s_in[F_HROWS * F_PITCH];
    __shared__ unsigned s_lut[RGB444 ? 4096 : U8 ? 192 : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    if (RGB444) for (int i = threadIdx.x; i < 4096; i += F_THREADS) s_lut[i] = (unsigned)__ldg(p.lut + i);
    if (U8) for (int i = threadIdx.x; i < 192; i += F_THREADS) s_lut[i] = __ldg(reinterpret_cast<const unsigned *>(p.lut8) + i);
    const unsigned char *s_lut8 = reinterpret_cast<const unsigned char *>(s_lut);

    // B fragments.  N-tile n, column c <-> output channel 4*(c>>1) + 2*n + (c&1), so that the C fragment of thread t
    // (columns 2t, 2t+1 of both tiles) is channels 4t .. 4t+3.
    const unsigned *gw = reinterpret_cast<const unsigned *>(p.wgt);       // word (o*9 + tap) = the 4 channel bytes of one tap
    unsigned b0[2], b1[2], b2[2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const int o = 4 * (g >> 1) + 2 * n + (g & 1);
        b0[n] = __ldg(gw + o * 9 + t);             // k32 step, k = 4t..4t+3    : tap t
        b1[n] = __ldg(gw + o * 9 + 4 + t);         //           k = 16+4t..     : tap 4+t
        b2[n] = t == 0 ? __ldg(gw + o * 9 + 8) : 0u;   // k16 step: tap 8, then three zero taps
    }
    const int4 bias = *reinterpret_cast<const int4 *>(p.bias_sh + 4 * t);
    int4 bw = bias;
    if (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI)
        bw = make_int4(__float_as_int((float)bias.x), __float_as_int((float)bias.y), __float_as_int((float)bias.z), __float_as_int((float)bias.w));

    // tap offsets (words) of this thread's A registers: tap t, tap 4+t, tap 8
    const int o_a = (t / 3) * F_PITCH + (t % 3);
    const int o_b = ((4 + t) / 3) * F_PITCH + ((4 + t) % 3);
    const int o_c = 2 * F_PITCH + 2;
    unsigned ovf = 0;

    const int tiles_x = (p.W + F_TW - 1) / F_TW, tiles_y = (p.H + F_TH - 1) / F_TH;
    const int tiles_img = tiles_x * tiles_y;
    const long long total = (long long)tiles_img * p.n_img;
    // haloed input tile, zero outside the image (= the convolution's zero padding): warp w stages halo rows w, w+8, w+16;
    // lane l takes column l, lanes 0 and 1 also columns 32 and 33.  (Prefetching tile k+1 through registers during the MMAs
    // of tile k was tried: +14 registers cost a resident CTA per SM and made the kernel 15 % slower.)
    constexpr int ROWS_PER_WARP = (F_HROWS + F_THREADS / 32 - 1) / (F_THREADS / 32);
    unsigned raw[ROWS_PER_WARP][2];
    auto fetch = [&](long long tile) {
        const int img = (int)(tile / tiles_img), tin = (int)(tile - (long long)img * tiles_img);
        const int ty = tin / tiles_x, tx = tin - ty * tiles_x;
        const int x0 = tx * F_TW, y0 = ty * F_TH;
#pragma unroll
        for (int k = 0; k < ROWS_PER_WARP; ++k) {
            const int hy = warp + k * (F_THREADS / 32);
            const int y = y0 - 1 + hy;
            const bool rowok = hy < F_HROWS && (unsigned)y < (unsigned)p.H;
            const size_t rowbase = ((size_t)img * p.H + (rowok ? y : 0)) * p.W;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int hx = lane + 32 * r;
                const int x = x0 - 1 + hx;
                unsigned v = SRC ? 0xffffffffu : 0u;                   // table sources: all ones marks "outside the image"
                if (rowok && hx < F_TW + 2 && (unsigned)x < (unsigned)p.W) {
                    if (RGB444) v = (unsigned)__ldg(p.in16 + rowbase + x);
                    else if (U8) { const uint8_t *px = p.in8 + 3 * (rowbase + x); v = (unsigned)__ldg(px) | ((unsigned)__ldg(px + 1) << 8) | ((unsigned)__ldg(px + 2) << 16); }
                    else v = __ldg(reinterpret_cast<const unsigned *>(p.in) + rowbase + x);
                }
                raw[k][r] = v;
            }
        }
    };
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int img = (int)(tile / tiles_img), tin = (int)(tile - (long long)img * tiles_img);
        const int ty = tin / tiles_x, tx = tin - ty * tiles_x;
        const int x0 = tx * F_TW, y0 = ty * F_TH;
        fetch(tile);                                                   // global loads first, then the barrier
        __syncthreads();                                               // previous tile fully consumed (first pass: tables loaded)
#pragma unroll
        for (int k = 0; k < ROWS_PER_WARP; ++k) {
            const int hy = warp + k * (F_THREADS / 32);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int hx = lane + 32 * r;
                if (hy < F_HROWS && hx < F_TW + 2) {
                    unsigned v = raw[k][r];
                    if (RGB444) {
                        v = v == 0xffffffffu ? 0u : s_lut[v & 0xfffu];
                    } else if (U8) {
                        const unsigned c = v;      // B | G << 8 | R << 16
                        v = c == 0xffffffffu ? 0u : ((unsigned)s_lut8[(c >> 16) & 255u] | ((unsigned)s_lut8[256 + ((c >> 8) & 255u)] << 8) | ((unsigned)s_lut8[512 + (c & 255u)] << 16));
                    }
                    s_in[hy * F_PITCH + hx] = v;
                }
            }
        }
        __syncthreads();

        if (POOL) {
            const int pr0 = 2 * (warp >> 1), pc0 = 8 * (warp & 1);           // pooled row / column origin inside the tile
            int acc[4][2][4];
#pragma unroll
            for (int ph = 0; ph < 4; ++ph)
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[ph][n][j] = 0;
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                const int dy = ph >> 1, dx = ph & 1;
                // rows g / g+8 of the M=16 tile: pooled pixel (pr0, pc0+g) / (pr0+1, pc0+g), member (dy,dx)
                const unsigned *r0 = s_in + (2 * pr0 + dy) * F_PITCH + 2 * (pc0 + g) + dx;
                const unsigned *r1 = r0 + 2 * F_PITCH;
                const unsigned a0 = r0[o_a], a1 = r1[o_a], a2 = r0[o_b], a3 = r1[o_b], a4 = r0[o_c], a5 = r1[o_c];
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    mma_s8_k32(acc[ph][n], a0, a1, a2, a3, b0[n], b1[n]);
                    mma_s8_k16(acc[ph][n], a4, a5, b2[n]);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int m[4];
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        m[2 * n + c] = max(max(acc[0][n][2 * h + c], acc[1][n][2 * h + c]), max(acc[2][n][2 * h + c], acc[3][n][2 * h + c]));
                const int oy = (y0 >> 1) + pr0 + h, ox = (x0 >> 1) + pc0 + g;
                const bool valid = oy < p.OH && ox < p.OW;
                const unsigned w = requant4v<EPI, ACT>(m, bw, p, ovf, valid);
                if (valid) *reinterpret_cast<unsigned *>(p.out + (((size_t)img * p.OH + oy) * p.OW + ox) * p.cs_out + 4 * t) = w;
            }
        } else {
            const int r = 2 * warp;                                            // rows r, r+1 of the tile
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
                int acc[2][4];
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[n][j] = 0;
                const unsigned *r0 = s_in + r * F_PITCH + 8 * cg + g;
                const unsigned *r1 = r0 + F_PITCH;
                const unsigned a0 = r0[o_a], a1 = r1[o_a], a2 = r0[o_b], a3 = r1[o_b], a4 = r0[o_c], a5 = r1[o_c];
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    mma_s8_k32(acc[n], a0, a1, a2, a3, b0[n], b1[n]);
                    mma_s8_k16(acc[n], a4, a5, b2[n]);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int m[4] = { acc[0][2 * h], acc[0][2 * h + 1], acc[1][2 * h], acc[1][2 * h + 1] };
                    const int y = y0 + r + h, x = x0 + 8 * cg + g;
                    const bool valid = y < p.H && x < p.W;
                    const unsigned w = requant4v<EPI, ACT>(m, bw, p, ovf, valid);
                    if (valid) *reinterpret_cast<unsigned *>(p.out + (((size_t)img * p.H + y) * p.W + x) * p.cs_out + 4 * t) = w;
                }
            }
        }
    }
    if (p.q.contract == CONTRACT_P) {
        ovf = __reduce_add_sync(0xffffffffu, ovf);
        if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
    }
}

bool conv3x3_first_supported(const ConvArgs &a)
{
    if (a.cs_in != 4 || a.cs_out != 16 || a.w_rows < 16) return false;
    if (a.q.pool && (a.H < 2 || a.W < 2)) return false;
    return true;
}

template <bool POOL, int EPI, int SRC>
static cudaError_t launch_first3(const FirstParams &p, cudaStream_t st)
{
    const long long total = (long long)((p.W + F_TW - 1) / F_TW) * ((p.H + F_TH - 1) / F_TH) * p.n_img;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent grid = exactly the CTAs that are resident at once
    int per_sm = 0;
    cudaError_t e = p.q.activ ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv3x3_first_kernel<POOL, EPI, true, SRC>, F_THREADS, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, conv3x3_first_kernel<POOL, EPI, false, SRC>, F_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sms * per_sm;
    const int grid = (int)(total < cap ? total : cap);
    if (p.q.activ) conv3x3_first_kernel<POOL, EPI, true, SRC><<<grid, F_THREADS, 0, st>>>(p);
    else conv3x3_first_kernel<POOL, EPI, false, SRC><<<grid, F_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

template <bool POOL, int EPI>
static cudaError_t launch_first2(const FirstParams &p, cudaStream_t st)
{
    if (p.in16) return launch_first3<POOL, EPI, 1>(p, st);
    if (p.in8) return launch_first3<POOL, EPI, 2>(p, st);
    return launch_first3<POOL, EPI, 0>(p, st);
}

template <bool POOL>
static cudaError_t launch_first(const ConvArgs &a, FirstParams &p, cudaStream_t st)
{
    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE:      return launch_first2<POOL, EPI_F_RNE>(p, st);
    case EPI_F_RNE_NOHI: return launch_first2<POOL, EPI_F_RNE_NOHI>(p, st);
    case EPI_P:          return launch_first2<POOL, EPI_P>(p, st);
    default:             return launch_first2<POOL, EPI_GENERIC>(p, st);
    }
}

// src_kind 1 / 2: fused RGB444 / uint8-BGR front end (a.in is ignored); lut = the context's table for that source
cudaError_t conv3x3_first(const ConvArgs &a, cudaStream_t st, int src_kind, const void *src, const void *lut)
{
    if (a.n == 0) return cudaSuccess;
    if (!conv3x3_first_supported(a)) return cudaErrorInvalidValue;
    FirstParams p;
    memset(&p, 0, sizeof p);
    p.in = a.in; p.n_img = a.n; p.H = a.H; p.W = a.W;
    if (src_kind == 1) { p.in16 = (const uint16_t *)src; p.lut = (const int *)lut; }
    else if (src_kind == 2) { p.in8 = (const uint8_t *)src; p.lut8 = (const uint8_t *)lut; }
    p.OH = a.q.pool ? a.H / 2 : a.H; p.OW = a.q.pool ? a.W / 2 : a.W;
    p.cs_out = a.cs_out; p.wgt = a.wgt; p.bias_sh = a.bias_sh; p.q = a.q; p.out = a.out; p.ovf = a.ovf;
    return a.q.pool ? launch_first<true>(a, p, st) : launch_first<false>(a, p, st);
}

}  // namespace yb
