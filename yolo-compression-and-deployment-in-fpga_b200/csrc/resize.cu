// Bilinear resize of uint8 HxWx3 images in front of the first layer: the GPU replacement of
// `cv2.resize(image, (size[1], size[0]))` in base_transform (reference data/__init__.py:36), bit-exact with OpenCV's 8-bit
// INTER_LINEAR (11-bit fixed-point taps, horizontal pass at scale 2^11, vertical pass `((b*(H>>4))>>16)`, `(+2)>>2`).
// The per-column / per-row taps and weights are computed once on the host in the float/double arithmetic OpenCV uses
// (resize_axis_table) and kept in two small device tables; exact 2x decimation, where OpenCV switches to the 2x2 area
// mean, falls out of the same formula (both weights are 1024 and the shifts are exact), so there is one kernel.
//
// HBM-bound byte work: per output pixel 12 source bytes (mostly L1/L2 hits: neighbouring outputs share taps) and 3 output
// bytes.  A thread produces four consecutive pixels of one output row = 12 bytes = three aligned 32-bit stores, so a warp
// writes 384 contiguous bytes; widths that are not a multiple of 4 (or unaligned buffers) take the one-pixel variant.
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

__device__ __forceinline__ int resize_px(const uint8_t *__restrict__ r0, const uint8_t *__restrict__ r1, int4 xt, int b0, int b1, int ch)
{
    const int h0 = (int)__ldg(r0 + xt.x + ch) * xt.z + (int)__ldg(r0 + xt.y + ch) * xt.w;
    const int h1 = (int)__ldg(r1 + xt.x + ch) * xt.z + (int)__ldg(r1 + xt.y + ch) * xt.w;
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;        // <= 255 by construction
}

// PX = 4: one thread = 4 consecutive pixels of a row (dw % 4 == 0, dst 4-byte aligned); PX = 1: one pixel.
template <int PX>
__global__ void __launch_bounds__(256) resize_u8c3_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                          const int4 *__restrict__ xtab, const int4 *__restrict__ ytab,
                                                          int n, int sh, int sw, int dh, int dw)
{
    const int per_row = dw / PX;
    const size_t items = (size_t)n * dh * per_row;
    const size_t src_frame = (size_t)sh * sw * 3, src_row = (size_t)sw * 3;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % per_row);
        const size_t row = it / per_row;                      // frame * dh + dy
        const int dy = (int)(row % dh);
        const size_t frame = row / dh;
        const int4 yt = __ldg(ytab + dy);
        const uint8_t *r0 = src + frame * src_frame + (size_t)yt.x * src_row;
        const uint8_t *r1 = src + frame * src_frame + (size_t)yt.y * src_row;
        if (PX == 4) {
            unsigned v[12];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int4 xt = __ldg(xtab + q * 4 + p);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) v[p * 3 + ch] = (unsigned)resize_px(r0, r1, xt, yt.z, yt.w, ch);
            }
            unsigned *o = reinterpret_cast<unsigned *>(dst + (row * dw + (size_t)q * 4) * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o[k] = v[4 * k] | (v[4 * k + 1] << 8) | (v[4 * k + 2] << 16) | (v[4 * k + 3] << 24);
        } else {
            const int4 xt = __ldg(xtab + q);
            uint8_t *o = dst + (row * dw + q) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) o[ch] = (uint8_t)resize_px(r0, r1, xt, yt.z, yt.w, ch);
        }
    }
}

cudaError_t resize_u8bgr(const uint8_t *src, int n, int sh, int sw, uint8_t *dst, int dh, int dw,
                         const int4 *xtab_dev, const int4 *ytab_dev, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const bool vec = (dw % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0);
    const size_t items = (size_t)n * dh * (vec ? dw / 4 : dw);
    size_t blocks = (items + 255) / 256;
    const size_t cap = (size_t)(sm_count > 0 ? sm_count : 148) * 8;               // whole waves of 8 resident CTAs per SM
    if (blocks > cap) blocks = cap;
    if (vec) resize_u8c3_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(src, dst, xtab_dev, ytab_dev, n, sh, sw, dh, dw);
    else     resize_u8c3_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(src, dst, xtab_dev, ytab_dev, n, sh, sw, dh, dw);
    return cudaGetLastError();
}

}  // namespace yb
