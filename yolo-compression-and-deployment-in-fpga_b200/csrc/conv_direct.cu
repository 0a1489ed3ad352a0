// conv_direct.cu — 3x3/stride-1/pad-1 int8 NHWC convolution on the integer dot-product pipe (dp4a),
// with the layer's requantisation, leaky-ReLU and optional 2x2 max-pool fused into the epilogue.
//
// Used for the thin first layer (3->16 channels, K = 27: HBM/issue bound, no tensor-core shape) and as the
// exact reference path every tensor-core kernel is checked against on the device.
//
// Replaces first_conv/second_conv/conv_normal/conv_last (c_embedding/yolo_forward.c:269,420,575,772) and
// Conv2d_fuse + tracker + pool (models/slim_yolo_v2.py:220-328).
//
// Tiling: one CTA = 8 x 16 output pixels (pre-pool) x COT output channels; one thread = one 2x2 pixel quad x 8
// output channels, so a pool window never leaves a thread.  Input channels are consumed in chunks of CK bytes
// staged in shared memory with their 1-pixel halo (zero filled = the convolution's zero padding).
#include "common.cuh"
#include "kernels.h"
#include <climits>

namespace yb {

constexpr int TILE_H = 8, TILE_W = 16;
constexpr int QUADS = (TILE_H / 2) * (TILE_W / 2);   // 32 quads per tile -> one warp per 8-channel group

// CK: bytes of input channels per chunk (4 for the NHWC4 network input, 16 otherwise)
// COT: output channels per CTA (multiple of 8)
template <int CK, int COT>
__global__ void __launch_bounds__(QUADS *(COT / 8))
conv3x3_direct_kernel(const int8_t *__restrict__ in, int n_img, int H, int W, int cs_in,   // cs_in: channel stride (bytes)
                      const int8_t *__restrict__ wgt,                                      // [cout_pad][9][cs_in]
                      const int *__restrict__ bias_sh,                                     // [cout_pad]
                      int cout, int cs_out, LayerQ q, int8_t *__restrict__ out, unsigned *__restrict__ ovf_counter,
                      int *__restrict__ stats)
{
    constexpr int PH = TILE_H + 2, PW = TILE_W + 2;
    constexpr int WPC = CK / 4;                       // 32-bit words per pixel per chunk
    __shared__ int s_in[PH * PW * WPC];
    __shared__ int s_w[COT * 9 * WPC];

    const int tiles_x = (W + TILE_W - 1) / TILE_W;
    const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
    const int co_base = blockIdx.y * COT;
    const int img = blockIdx.z;
    const int quad = threadIdx.x % QUADS, cog = threadIdx.x / QUADS;
    const int qy = quad / (TILE_W / 2), qx = quad % (TILE_W / 2);
    const int y0 = tile_y * TILE_H, x0 = tile_x * TILE_W;

    int acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[p][c] = 0;

    const int8_t *img_in = in + (size_t)img * H * W * cs_in;
    for (int ck = 0; ck < cs_in; ck += CK) {
        __syncthreads();
        // stage the (TILE_H+2) x (TILE_W+2) x CK patch, zero outside the image
        for (int i = threadIdx.x; i < PH * PW * WPC; i += blockDim.x) {
            int wd = i % WPC, px = (i / WPC) % PW, py = i / (WPC * PW);
            int gy = y0 + py - 1, gx = x0 + px - 1;
            int v = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = *reinterpret_cast<const int *>(img_in + ((size_t)gy * W + gx) * cs_in + ck + 4 * wd);
            s_in[i] = v;
        }
        // stage the weights of this chunk: [COT][9][CK]
        for (int i = threadIdx.x; i < COT * 9 * WPC; i += blockDim.x) {
            int wd = i % WPC, tap = (i / WPC) % 9, co = i / (WPC * 9);
            s_w[i] = *reinterpret_cast<const int *>(wgt + ((size_t)(co_base + co) * 9 + tap) * cs_in + ck + 4 * wd);
        }
        __syncthreads();
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int wd = 0; wd < WPC; ++wd) {
                    int a[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        int py = 2 * qy + (p >> 1) + kh, px = 2 * qx + (p & 1) + kw;
                        a[p] = s_in[(py * PW + px) * WPC + wd];
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        int wv = s_w[((cog * 8 + c) * 9 + kh * 3 + kw) * WPC + wd];
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[p][c] = dp4a_s8(a[p], wv, acc[p][c]);
                    }
                }
    }

    // epilogue: requantise each of the 4 pixels, then (optionally) max over the quad
    unsigned ovf = 0;
    const int co0 = co_base + cog * 8;
    int bsh[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bsh[c] = bias_sh[co0 + c];
    if (stats) {
        // calibration pass (yolo_b200_calibrate_f32): no output; stats[0] = max, stats[1] = min over every pre-pool
        // element of the layer numerator num = (acc << la) + (b << lb), the real-valued activation at scale 2^-E
        int mx = INT_MIN, mn = INT_MAX;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int oy = y0 + 2 * qy + (p >> 1), ox = x0 + 2 * qx + (p & 1);
            if (oy < H && ox < W)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (co0 + c < cout) { const int num = (acc[p][c] << q.la) + bsh[c]; mx = max(mx, num); mn = min(mn, num); }
        }
        mx = __reduce_max_sync(0xffffffffu, mx); mn = __reduce_min_sync(0xffffffffu, mn);
        if ((threadIdx.x & 31) == 0) { atomicMax(&stats[0], mx); atomicMin(&stats[1], mn); }
        return;
    }
    int o[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) o[p][c] = (co0 + c < cout) ? requant(acc[p][c], bsh[c], q) : 0;

    if (co0 < cs_out) {
        if (q.pool) {
            int OH = H / 2, OW = W / 2;
            int oy = y0 / 2 + qy, ox = x0 / 2 + qx;
            if (oy < OH && ox < OW) {
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    int m = store8(max(max(o[0][c], o[1][c]), max(o[2][c], o[3][c])), ovf);
                    if (c < 4) lo |= (unsigned)(m & 0xff) << (8 * c); else hi |= (unsigned)(m & 0xff) << (8 * (c - 4));
                }
                int8_t *dst = out + (((size_t)img * OH + oy) * OW + ox) * cs_out + co0;
                *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                int oy = y0 + 2 * qy + (p >> 1), ox = x0 + 2 * qx + (p & 1);
                if (oy < H && ox < W) {
                    unsigned lo = 0, hi = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        int v = store8(o[p][c], ovf);
                        if (c < 4) lo |= (unsigned)(v & 0xff) << (8 * c); else hi |= (unsigned)(v & 0xff) << (8 * (c - 4));
                    }
                    int8_t *dst = out + (((size_t)img * H + oy) * W + ox) * cs_out + co0;
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
                }
            }
        }
    }
    if (q.contract == CONTRACT_P) {
        ovf = __reduce_add_sync(0xffffffffu, ovf);
        if ((threadIdx.x & 31) == 0 && ovf) atomicAdd(ovf_counter, ovf);
    }
}

template <int CK, int COT>
static cudaError_t launch(const ConvArgs &a, cudaStream_t st)
{
    int tiles = ((a.H + TILE_H - 1) / TILE_H) * ((a.W + TILE_W - 1) / TILE_W);
    dim3 grid(tiles, (a.cs_out + COT - 1) / COT, a.n);
    conv3x3_direct_kernel<CK, COT><<<grid, QUADS *(COT / 8), 0, st>>>(a.in, a.n, a.H, a.W, a.cs_in, a.wgt, a.bias_sh,
                                                                       a.cout, a.cs_out, a.q, a.out, a.ovf, a.stats);
    return cudaGetLastError();
}

// Weight buffer must be padded to a multiple of 32 output channels (zeros) so any COT tile can read it.
cudaError_t conv3x3_direct(const ConvArgs &a, cudaStream_t st)
{
    if (a.cs_in == 4) return launch<4, 16>(a, st);
    if (a.cs_out % 32 == 0) return launch<16, 32>(a, st);
    return launch<16, 16>(a, st);
}

}  // namespace yb
