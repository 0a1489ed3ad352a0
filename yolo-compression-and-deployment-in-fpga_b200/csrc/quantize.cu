// quantize.cu — the two input quantisers in front of layer 0 (bandwidth-bound, vectorised).
#include "kernels.h"

namespace yb {

// pixel_norm_quantize (c_embedding/yolo_forward.c:57-85) is a pure function of the 12-bit 0x0BGR code: the host
// evaluates the reference arithmetic once per code into a 4096-entry table of packed (R,G,B,0) words
// (yolo_b200.cu: build_rgb444_lut); the kernel is a gather through shared memory.
__global__ void __launch_bounds__(256) quantize_rgb444_kernel(const uint16_t *__restrict__ frames, size_t npix,
                                                              const int *__restrict__ lut, int *__restrict__ out)
{
    __shared__ int s_lut[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    // 8 pixels (16 B in, 32 B out) per thread per step; a frame pointer that is not 16-byte aligned takes the scalar loop
    size_t nvec = (reinterpret_cast<uintptr_t>(frames) & 15) ? 0 : npix / 8;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        uint4 p = reinterpret_cast<const uint4 *>(frames)[v];
        unsigned w[4] = { p.x, p.y, p.z, p.w };
        int4 o0, o1;
        o0.x = s_lut[w[0] & 0xfff]; o0.y = s_lut[(w[0] >> 16) & 0xfff];
        o0.z = s_lut[w[1] & 0xfff]; o0.w = s_lut[(w[1] >> 16) & 0xfff];
        o1.x = s_lut[w[2] & 0xfff]; o1.y = s_lut[(w[2] >> 16) & 0xfff];
        o1.z = s_lut[w[3] & 0xfff]; o1.w = s_lut[(w[3] >> 16) & 0xfff];
        reinterpret_cast<int4 *>(out)[2 * v] = o0;
        reinterpret_cast<int4 *>(out)[2 * v + 1] = o1;
    }
    // tail
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) out[i] = s_lut[frames[i] & 0xfff];
}

cudaError_t quantize_rgb444(const uint16_t *frames, size_t npix, const int *lut_dev, int8_t *nhwc4, cudaStream_t st)
{
    size_t nvec = ((uintptr_t)frames & 15) ? (npix + 7) / 8 : npix / 8;      // sizes the grid only
    int blocks = (int)((nvec + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    quantize_rgb444_kernel<<<blocks, 256, 0, st>>>(frames, npix, lut_dev, reinterpret_cast<int *>(nhwc4));
    return cudaGetLastError();
}

// BaseTransform without the resize (data/__init__.py:30-56: /255, -mean, /std on the BGR bytes cv2 delivers), test.py:79's
// BGR -> RGB swap and a_tracker_in's quantisation (slim_yolo_v2.py:218,35) are a pure function of one byte per channel:
// the host evaluates the reference's float32 arithmetic once per byte value into three 256-entry tables
// (yolo_b200.cu: build_u8_lut); lut8[0..255] = R (from BGR byte 2), [256..511] = G, [512..767] = B.
// lut8[768..1535] flags the byte values whose quantised value had to be saturated (the reference never clamps): counted.
__global__ void __launch_bounds__(256) quantize_u8bgr_kernel(const uint8_t *__restrict__ bgr, size_t npix,
                                                             const uint8_t *__restrict__ lut8, int *__restrict__ out,
                                                             unsigned *__restrict__ ovf_counter)
{
    __shared__ uint8_t s_lut[1536];
    for (int i = threadIdx.x; i < 1536; i += blockDim.x) s_lut[i] = lut8[i];
    __syncthreads();
    unsigned ovf = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *px = bgr + 3 * i;
        const unsigned r = px[2], g = 256u + px[1], b = 512u + px[0];
        out[i] = (int)((unsigned)s_lut[r] | ((unsigned)s_lut[g] << 8) | ((unsigned)s_lut[b] << 16));
        ovf += s_lut[768 + r] + s_lut[768 + g] + s_lut[768 + b];
    }
    ovf = __reduce_add_sync(0xffffffffu, ovf);
    if ((threadIdx.x & 31) == 0 && ovf) atomicAdd(ovf_counter, ovf);
}

cudaError_t quantize_u8bgr(const uint8_t *bgr, size_t npix, const uint8_t *lut8_dev, int8_t *nhwc4, unsigned *ovf, cudaStream_t st)
{
    int blocks = (int)((npix + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    quantize_u8bgr_kernel<<<blocks, 256, 0, st>>>(bgr, npix, lut8_dev, reinterpret_cast<int *>(nhwc4), ovf);
    return cudaGetLastError();
}

// max |x| over a float tensor (AveragedRangeTracker's min/max, slim_yolo_v2.py:22-23): non-negative floats order like their
// bit patterns, so the reduction is an integer atomicMax.  *out_bits must be zeroed by the caller.
__global__ void __launch_bounds__(256) absmax_f32_kernel(const float *__restrict__ x, size_t count, unsigned *__restrict__ out_bits)
{
    unsigned m = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(x[i])));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, m);
}

cudaError_t absmax_f32(const float *x, size_t count, unsigned *out_bits, cudaStream_t st)
{
    int blocks = (int)((count + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    absmax_f32_kernel<<<blocks, 256, 0, st>>>(x, count, out_bits);
    return cudaGetLastError();
}

// a_tracker_in.quantize_activation with a frozen power-of-two scale (models/slim_yolo_v2.py:33-35):
// q = round-half-even(x * 2^sa).  The reference does not clamp; int8 storage saturates and counts.
// float NCHW (3 planes) -> int8 NHWC4, 4 pixels per thread (3 x 16 B loads, one 16 B store).
__global__ void __launch_bounds__(256) quantize_f32_kernel(const float *__restrict__ nchw, int n, size_t plane, float scale,
                                                           int8_t *__restrict__ out, unsigned *__restrict__ ovf_counter)
{
    size_t nvec = plane / 4;
    unsigned ovf = 0;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec * n; v += (size_t)gridDim.x * blockDim.x) {
        size_t img = v / nvec, i4 = v % nvec;
        const float *base = nchw + img * 3 * plane + 4 * i4;
        float4 c[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) c[ch] = *reinterpret_cast<const float4 *>(base + ch * plane);
        unsigned w[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            unsigned word = 0;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float x = (p == 0 ? c[ch].x : p == 1 ? c[ch].y : p == 2 ? c[ch].z : c[ch].w) * scale;
                float r = rintf(x);
                float cl = fminf(fmaxf(r, -128.f), 127.f);
                ovf += (cl != r);
                word |= ((unsigned)((int)cl) & 0xffu) << (8 * ch);
            }
            w[p] = word;
        }
        reinterpret_cast<uint4 *>(out)[img * nvec + i4] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    ovf = __reduce_add_sync(0xffffffffu, ovf);
    if ((threadIdx.x & 31) == 0 && ovf) atomicAdd(ovf_counter, ovf);
}

cudaError_t quantize_f32(const float *nchw, int n, int h, int w, int sa, int8_t *nhwc4, unsigned *ovf, cudaStream_t st)
{
    size_t plane = (size_t)h * w;
    if (plane % 4) return cudaErrorInvalidValue;      // callers guarantee w % 4 == 0 (network stride is 16)
    size_t work = plane / 4 * n;
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    quantize_f32_kernel<<<blocks, 256, 0, st>>>(nchw, n, plane, ldexpf(1.0f, sa), nhwc4, ovf);
    return cudaGetLastError();
}

}  // namespace yb
