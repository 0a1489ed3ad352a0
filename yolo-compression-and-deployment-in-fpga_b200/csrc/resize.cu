// Bilinear resize of uint8 HxWx3 images in front of the first layer: the GPU replacement of
// `cv2.resize(image, (size[1], size[0]))` in base_transform (reference data/__init__.py:36), bit-exact with OpenCV's 8-bit
// INTER_LINEAR (11-bit fixed-point taps, horizontal pass at scale 2^11, vertical pass `((b*(H>>4))>>16)`, `(+2)>>2`).
// The per-column / per-row taps and weights are computed once on the host in the float/double arithmetic OpenCV uses
// (resize_axis_table) and kept in two small device tables; exact 2x decimation, where OpenCV switches to the 2x2 area
// mean, falls out of the same formula (both weights are 1024 and the shifts are exact), so there is one kernel.
//
// HBM-bound byte work: per output pixel 12 source bytes (mostly L1/L2 hits: neighbouring outputs share taps) and 3 output
// bytes.  What limits such a kernel is the number of 128-byte lines each load instruction touches, so the source is read
// through aligned 32-bit words (see resize_row_taps) and a warp's 96 output bytes leave as 24 word stores.
// History (256 frames 480x640 -> 416x416): four pixels per thread with byte loads 0.344 ms (every byte load touched 5-6
// lines); one pixel per thread numbered flat over the batch with word loads 0.40 ms -- ncu: 226 instructions per pixel,
// issue-bound on the 64-bit index divisions; one row per CTA iteration (grid-stride) with DP2A / high-multiply
// arithmetic 0.376 ms -- still 200 instructions per 32-pixel chunk, 72 of them row bookkeeping repeated by every warp and
// 24 of them 64-bit addresses of the clamped word loads; this version (one warp per run of contiguous rows, clamping only
// in the rows that can reach the end of the buffer) 0.207 ms at 47 registers / 40 resident warps per SM (98 instructions per
// chunk, 57 % issue utilisation, long-scoreboard stalls: latency-bound), 0.193 ms (ncu: profiles/resize_r1_ncu.txt) =
// 1.91 TB/s of algorithmic bytes capped at 40 registers / 48 warps (what ships).  Measured slower: the chunk loop unrolled
// by two (54 registers, 32 warps: 0.233 ms), a 32-register cap (64 warps, spills: 0.226 ms).
#include "kernels.h"

#include <cmath>
#include <vector>

namespace yb {

void resize_axis_table(int src, int dst, bool clamp_weights, int index_scale, int4 *out)
{
    // cv::resize: inv_scale = (double)dst/src; cv::hal::resize: scale = 1./inv_scale
    volatile double inv_scale = (double)dst / (double)src;
    volatile double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        volatile double t = (d + 0.5) * scale;                // volatile: no fused multiply-add on the host
        float f = (float)(t - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (clamp_weights) {                                   // horizontal pass: the tap is re-anchored at the border
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src - 1) { s = src - 1; f = 0.f; }
        }
        const int c0 = (int)lrintf((1.f - f) * 2048.f), c1 = (int)lrintf(f * 2048.f);    // saturate_cast<short>: RNE
        int i0 = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
        int i1 = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
        out[d] = make_int4(i0 * index_scale, i1 * index_scale, c0, c1);
    }
}

// One source row's contribution: the six bytes (tap 0 = bytes 0..2, tap 1 = bytes 3..5) that start `off` bytes into a
// word-aligned view of the row, fetched as three aligned 32-bit words and realigned with funnel shifts (a warp's lanes
// are ~4.6 bytes apart for 640 -> 416, so a word load touches two 128-byte lines).  CLAMP (only the rows whose windows
// could reach past the end of the buffer, i.e. the last source rows of the last image): byte loads with indices clamped
// to `lim`, the last byte of the buffer relative to this row; a clamped byte only ever supplies a zero-weight value
// (tap 1 of a right-border pixel).  compute-sanitizer memcheck runs clean on the GPU tests with exact-size allocations.
template <bool CLAMP>
__device__ __forceinline__ void resize_row_taps(const unsigned char *__restrict__ rowb, unsigned off, unsigned lim, unsigned &lo, unsigned &hi)
{
    if (CLAMP) {
        // byte loads, indices clamped to the last byte of the buffer (`lim`, relative to rowb)
        lo = 0; hi = 0;
#pragma unroll
        for (unsigned k = 0; k < 4; ++k) lo |= (unsigned)__ldg(rowb + min(off + k, lim)) << (8 * k);
#pragma unroll
        for (unsigned k = 0; k < 2; ++k) hi |= (unsigned)__ldg(rowb + min(off + 4 + k, lim)) << (8 * k);
    } else {
        const unsigned sh = (off & 3u) * 8u;
        const unsigned *a = reinterpret_cast<const unsigned *>(rowb + (off & ~3u));     // one address, three immediates
        const unsigned w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2);
        lo = __funnelshift_r(w0, w1, sh);          // bytes off .. off+3
        hi = __funnelshift_r(w1, w2, sh);          // bytes off+4 .. off+7
    }
}

// a01 = weight0 | weight1 << 16 (units of 1/2048), B0/B1 = row weights << 16.  Per channel: the horizontal pass of one
// row is one PRMT (tap 0 and tap 1 bytes side by side) + one unsigned DP2A; `(b * (h >> 4)) >> 16` is a high multiply.
__device__ __forceinline__ unsigned resize_combine(unsigned lo0, unsigned hi0, unsigned lo1, unsigned hi1, unsigned a01, unsigned B0, unsigned B1)
{
    unsigned o[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const unsigned sel = (unsigned)ch | ((unsigned)(3 + ch) << 4);          // bytes ch and 3 + ch of the 8-byte window
        const unsigned h0 = __dp2a_lo(a01, __byte_perm(lo0, hi0, sel), 0u);    // horizontal pass, scale 2^11
        const unsigned h1 = __dp2a_lo(a01, __byte_perm(lo1, hi1, sel), 0u);
        o[ch] = (__umulhi(B0, h0 >> 4) + __umulhi(B1, h1 >> 4) + 2u) >> 2;      // <= 255 by construction
    }
    return o[0] | (o[1] << 8) | (o[2] << 16);
}

// One output row of the batch: one thread = one output pixel, a warp = 32 consecutive pixels = 96 contiguous output bytes.
// WORD_STORE: the 32 three-byte pixels of a full warp are regrouped by two shuffles into 24 aligned 32-bit stores
// (output word j = bytes 4j .. 4j+3 of the chunk = pixels p = 4j/3 and p + 1 from byte rsh/8 of p); partial chunks and other
// geometries store bytes.
template <bool WORD_STORE, bool CLAMP>
__device__ __forceinline__ void resize_one_row(const unsigned char *__restrict__ r0, const unsigned char *__restrict__ r1,
                                               unsigned d0, unsigned d1, unsigned lim0, unsigned lim1, unsigned B0, unsigned B1,
                                               const int2 *__restrict__ xtab, uint8_t *__restrict__ drow, int dw,
                                               int lane, int warp, int nwarps, int p, unsigned rsh)
{
    for (int dx0 = warp * 32; dx0 < dw; dx0 += nwarps * 32) {
        const int dx = dx0 + lane;
        const bool live = dx < dw;
        unsigned v = 0;
        if (live) {
            const int2 xt = __ldg(xtab + dx);
            unsigned lo0, hi0, lo1, hi1;
            resize_row_taps<CLAMP>(r0, (unsigned)xt.x + d0, lim0, lo0, hi0);
            resize_row_taps<CLAMP>(r1, (unsigned)xt.x + d1, lim1, lo1, hi1);
            v = resize_combine(lo0, hi0, lo1, hi1, (unsigned)xt.y, B0, B1);
        }
        if (WORD_STORE && dx0 + 32 <= dw) {
            const unsigned va = __shfl_sync(0xffffffffu, v, p & 31), vb = __shfl_sync(0xffffffffu, v, (p + 1) & 31);
            const unsigned word = __funnelshift_r(va | (vb << 24), vb >> 8, rsh);
            if (lane < 24) reinterpret_cast<unsigned *>(drow)[(dx0 >> 5) * 24 + lane] = word;
        } else if (live) {
            uint8_t *o = drow + (size_t)dx * 3;
            o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16);
        }
    }
}

// One WARP = a contiguous run of output rows of the batch (so the row bookkeeping is incremental: no divisions), one
// whole row at a time, chunk after chunk: the ~70 instructions of row bookkeeping are paid once per row, not once per
// 32-pixel chunk (with one CTA per row every warp repeated them for its single chunk: 150 instead of 85 instructions per
// chunk).  Everything that depends on the row only (taps, weights, row pointers, their misalignment) is warp-uniform.
// WORD_STORE needs dst 4-byte aligned and dw % 4 == 0, so that every row and every 32-pixel chunk starts on a word.
template <bool WORD_STORE>
__global__ void __launch_bounds__(256, 6) resize_u8c3_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                          const int2 *__restrict__ xtab, const int4 *__restrict__ ytab,
                                                          unsigned rows, unsigned rows_per_warp, int sh, int sw, int dh, int dw, size_t src_bytes)
{
    const int lane = threadIdx.x & 31;
    unsigned row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rows_per_warp;
    if (row >= rows) return;
    const unsigned row_end = min(rows, row + rows_per_warp);
    const size_t src_frame = (size_t)sh * sw * 3, src_row = (size_t)sw * 3;
    const size_t delta = reinterpret_cast<uintptr_t>(src) & 3;                             // word-aligned view of the source
    const unsigned char *base = src - delta;
    const size_t end = delta + src_bytes;                                                  // one past the last source byte
    const int p = (4 * lane) / 3;
    const unsigned rsh = 8u * (unsigned)(4 * lane - 3 * p);
    unsigned dy = row % (unsigned)dh;
    size_t frame_off = delta + (size_t)(row / (unsigned)dh) * src_frame;
    uint8_t *drow = dst + (size_t)row * dw * 3;
    for (; row < row_end; ++row) {
        const int4 yt = __ldg(ytab + dy);
        const size_t o0 = frame_off + (size_t)yt.x * src_row, o1 = frame_off + (size_t)yt.y * src_row;
        const unsigned char *r0 = base + (o0 & ~(size_t)3), *r1 = base + (o1 & ~(size_t)3);
        const unsigned d0 = (unsigned)o0 & 3u, d1 = (unsigned)o1 & 3u;
        // a window reaches at most 11 bytes past the start of its pixel; rows for which that stays inside the buffer need no clamp
        const size_t omax = o0 > o1 ? o0 : o1;
        if (omax + src_row + 12 <= end) {
            resize_one_row<WORD_STORE, false>(r0, r1, d0, d1, 0u, 0u, (unsigned)yt.z, (unsigned)yt.w, xtab, drow, dw, lane, 0, 1, p, rsh);
        } else {
            const size_t room0 = end - 1 - (o0 & ~(size_t)3), room1 = end - 1 - (o1 & ~(size_t)3);
            resize_one_row<WORD_STORE, true>(r0, r1, d0, d1, (unsigned)room0, (unsigned)room1, (unsigned)yt.z, (unsigned)yt.w, xtab, drow, dw, lane, 0, 1, p, rsh);
        }
        drow += (size_t)dw * 3;
        if (++dy == (unsigned)dh) { dy = 0; frame_off += src_frame; }
    }
}

cudaError_t resize_u8bgr(const uint8_t *src, int n, int sh, int sw, uint8_t *dst, int dh, int dw,
                         const int2 *xtab_dev, const int4 *ytab_dev, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const size_t rows = (size_t)n * dh;
    if (rows > 0x7fffffffull) return cudaErrorInvalidValue;
    // one wave of resident CTAs (8 warps each), equal runs of rows per warp
    const int threads = 256;
    const size_t max_warps = (size_t)(sm_count > 0 ? sm_count : 148) * 6 * (threads / 32);     // 6 resident CTAs per SM at 40 registers
    const unsigned rows_per_warp = (unsigned)((rows + max_warps - 1) / max_warps);
    const size_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
    const size_t blocks = (warps + threads / 32 - 1) / (threads / 32);
    const size_t src_bytes = (size_t)n * sh * sw * 3;
    if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0 && dw % 4 == 0)
        resize_u8c3_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(src, dst, xtab_dev, ytab_dev, (unsigned)rows, rows_per_warp, sh, sw, dh, dw, src_bytes);
    else
        resize_u8c3_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(src, dst, xtab_dev, ytab_dev, (unsigned)rows, rows_per_warp, sh, sw, dh, dw, src_bytes);
    return cudaGetLastError();
}

}  // namespace yb
