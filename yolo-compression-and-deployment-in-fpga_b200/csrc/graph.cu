// graph.cu — the two data-movement ops yolo_v2 adds around the convolutions (models/yolo_v2.py:165-177): a stand-alone
// 2x2/2 max-pool (darknet19's maxpool_4 / maxpool_5 follow maps that a route also reads, backbone/darknet.py:78,88,99-104;
// slim_yolo_v2 fuses its pools into the convolution epilogues) and reorg + concat (reorg_layer, utils/modules.py:43-57;
// torch.cat([fp_1, fp_2], dim=1), yolo_v2.py:171-174) with the two sources brought to one activation exponent.
// Both are pure HBM streams: 16-byte (pool) / 4-byte (concat) vectors, coalesced along the channel axis.
#include "kernels.h"

namespace yb {

__global__ void __launch_bounds__(256) maxpool2x2_kernel(const int8_t *__restrict__ in, int n, int H, int W, int cs, int8_t *__restrict__ out)
{
    const int OH = H / 2, OW = W / 2, vec = cs / 16;
    const size_t total = (size_t)n * OH * OW * vec;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        size_t r = i / vec;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int img = (int)(r / OH);
        const uint4 *src = reinterpret_cast<const uint4 *>(in + (((size_t)img * H + 2 * oy) * W + 2 * ox) * cs) + v;
        const uint4 a = src[0], b = src[vec], c = src[(size_t)W * vec], d = src[(size_t)W * vec + vec];
        uint4 m;
        m.x = __vmaxs4(__vmaxs4(a.x, b.x), __vmaxs4(c.x, d.x));
        m.y = __vmaxs4(__vmaxs4(a.y, b.y), __vmaxs4(c.y, d.y));
        m.z = __vmaxs4(__vmaxs4(a.z, b.z), __vmaxs4(c.z, d.z));
        m.w = __vmaxs4(__vmaxs4(a.w, b.w), __vmaxs4(c.w, d.w));
        reinterpret_cast<uint4 *>(out + (((size_t)img * OH + oy) * OW + ox) * cs)[v] = m;
    }
}

cudaError_t maxpool2x2(const int8_t *in, int n, int H, int W, int cs, int8_t *out, cudaStream_t st)
{
    if (n == 0 || H < 2 || W < 2) return cudaSuccess;
    if (cs % 16) return cudaErrorInvalidValue;
    const size_t total = (size_t)n * (H / 2) * (W / 2) * (cs / 16);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    maxpool2x2_kernel<<<blocks, 256, 0, st>>>(in, n, H, W, cs, out);
    return cudaGetLastError();
}

// four packed int8 values -> sat8(rne(v >> s)) each
__device__ __forceinline__ unsigned shift_word(unsigned w, int s)
{
    if (s == 0) return w;
    unsigned o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = (int)(int8_t)(w >> (8 * k));
        o |= (unsigned)(clampi(shr_round_rt(v, s, ROUND_RNE), -128, 127) & 0xff) << (8 * k);
    }
    return o;
}

// out[n][h][w][cs_out] = cat(reorg?(A) >> sh_a, B >> sh_b); channels beyond ca_eff + cb are zero.
__global__ void __launch_bounds__(256) concat_kernel(const int8_t *__restrict__ A, int cs_a, int ca, int reorg, int sh_a,
                                                      const int8_t *__restrict__ B, int cs_b, int cb, int sh_b,
                                                      int n, int h, int w, int cs_out, int8_t *__restrict__ out)
{
    const int words = cs_out / 4;
    const int ca_eff = reorg ? 4 * ca : ca;
    const size_t total = (size_t)n * h * w * words;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = 4 * (int)(i % words);
        size_t r = i / words;
        const int x = (int)(r % w); r /= w;
        const int y = (int)(r % h);
        const int img = (int)(r / h);
        unsigned v = 0u;
        if (c < ca_eff) {
            if (reorg) {
                const int s = c / ca, cc = c - s * ca;               // (ca % 4 == 0: a word never straddles two sub-pixels)
                const int sy = 2 * y + (s >> 1), sx = 2 * x + (s & 1);
                v = *reinterpret_cast<const unsigned *>(A + (((size_t)img * (2 * h) + sy) * (2 * w) + sx) * cs_a + cc);
            } else v = *reinterpret_cast<const unsigned *>(A + (((size_t)img * h + y) * w + x) * cs_a + c);
            v = shift_word(v, sh_a);
        } else if (c < ca_eff + cb) {
            v = shift_word(*reinterpret_cast<const unsigned *>(B + (((size_t)img * h + y) * w + x) * cs_b + (c - ca_eff)), sh_b);
        }
        reinterpret_cast<unsigned *>(out)[i] = v;
    }
}

cudaError_t concat_reorg(const int8_t *A, int cs_a, int ca, int reorg, int sh_a, const int8_t *B, int cs_b, int cb, int sh_b,
                         int n, int h, int w, int cs_out, int8_t *out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (ca % 4 || cb % 4 || cs_out % 4 || sh_a < 0 || sh_b < 0) return cudaErrorInvalidValue;
    const size_t total = (size_t)n * h * w * (cs_out / 4);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    concat_kernel<<<blocks, 256, 0, st>>>(A, cs_a, ca, reorg, sh_a, B, cs_b, cb, sh_b, n, h, w, cs_out, out);
    return cudaGetLastError();
}

}  // namespace yb
