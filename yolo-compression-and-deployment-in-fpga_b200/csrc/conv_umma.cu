// conv_umma.cu — 3x3/stride-1/pad-1 int8 NHWC convolution as an implicit GEMM on the 5th-generation tensor
// cores (tcgen05.mma.kind::i8, int32 accumulators in TMEM), operands staged by TMA, with the layer's
// requantisation / leaky-ReLU / 2x2 max-pool fused into the epilogue so the int32 accumulators never reach HBM.
//
// Replaces second_conv/conv_normal/conv_last (c_embedding/yolo_forward.c:420,575,772): where the C driver walks
// 16x20 tiles x 32 output x 16 input channels through the FPGA's MAC array, this kernel walks 128-pixel tiles
// through UMMA with the whole K = 9*Cin reduction kept in one uninterrupted accumulation (set_offset is issued
// once per layer, yolo_forward.c:318,471,635,831, so the shift applies to the full-K sum).
//
// GEMM view:  D[128 pixels][N = cstride(cout)] += A[128][32 B] * B[N][32 B]^T   per tcgen05.mma (K = 32 int8)
//   A: for tap (kh,kw) and a chunk of CB input-channel bytes, the 128 pixels of the output tile shifted by the tap.
//      One 4-D TMA box {CB, TW, TH, TN} at (c0, x0+kw-1, y0+kh-1, n0): out-of-image coordinates are zero-filled by
//      the TMA unit, which IS the convolution's zero padding.  The box lands in shared memory as 128 rows of CB
//      bytes in the K-major UMMA canonical layout (swizzle mode = CB: 128B/64B/32B, none for CB=16).
//   B: the layer's weights [N][9*cs_in] (K-major, tap-major K), one 2-D TMA box {CB, N} per (tap, chunk).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 = tile rows).  Three pipelines: smem full/empty
// (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two accumulator buffers), and a static persistent tile loop.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <cuda.h>
#include <climits>
#include <cstring>

namespace yb {

// ---------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------
struct UmmaParams {
    int n_img, H, W;
    int TN, TH, TW;              // tile = TN images x TH rows x TW cols (<= 128 pixels; TH, TW even when pooling)
    int tiles_x, tiles_y, tiles_n, num_tiles;
    int N;                       // GEMM N = output channels of one slice (= cs_out when cs_out <= 256, else 256)
    int nslices;                 // cs_out / N: wide layers (yolo_v2 / darknet19: 512, 1024 channels) run as N-slices of 256 channels;
                                 // a work item is (tile, slice), so small maps still fill the SMs
    int taps;                    // 9 = 3x3, 1 = 1x1 (the single tap sits at the window centre)
    int cout, cs_out;
    int kblocks;                 // TMA box pairs per tile: 9 * cs_in / CB  (CB = 16: the 9 taps + 1 zero tap)
    int G;                       // kblocks per pipeline stage
    int stages;
    uint32_t a_box_bytes;        // TN*TH*TW*CB
    uint32_t stage_bytes;        // G * (128*CB + N*CB), multiple of 1024
    uint32_t tmem_cols, tmem_buf_stride;
    uint32_t off_stage, off_bias, off_bar;       // dynamic smem offsets (from the 1024-aligned base)
    LayerQ q;
    EpiConst k;
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    const uint8_t *w_swz;        // CB == 128: weights pre-swizzled into the shared-memory image of every (tap, chunk) block, [kb][cs_out][128]
};

constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each takes half of the columns
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int UMMA_THREADS = 64 + EPI_THREADS;

template <int EPI, bool ACT>
__device__ __forceinline__ void epilogue_tile(const UmmaParams &p, uint32_t taddr, int cbeg, int cend, int row, int et,
                                              int x0, int y0, int n0, const int *s_bias, int *s_stage, uint32_t bar_tempty,
                                              unsigned &ovf, int8_t *out)
{
    const int tile_px = p.TN * p.TH * p.TW;
    if (!p.q.pool) {
        const int wl = row % p.TW, hl = (row / p.TW) % p.TH, nl = row / (p.TW * p.TH);
        const int x = x0 + wl, y = y0 + hl, n = n0 + nl;
        const bool valid = row < tile_px && x < p.W && y < p.H && n < p.n_img;
        int8_t *dst = out + (((size_t)n * p.H + y) * p.W + x) * p.cs_out;
        // software pipeline: the TMEM load of chunk i+1 is in flight while chunk i is requantised
        int va[16], vb[16];
        if (cbeg < cend) tmem_ld16(taddr + cbeg, va);
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            tmem_ld_wait();
            if (c0 + 16 < cend) tmem_ld16(taddr + c0 + 16, vb);
            uint4 w = requant16<EPI, ACT>(va, s_bias, c0, p, ovf, valid);
            if (valid) *reinterpret_cast<uint4 *>(dst + c0) = w;
            if (c0 + 16 < cend) {
                tmem_ld_wait();
                if (c0 + 32 < cend) tmem_ld16(taddr + c0 + 32, va);
                w = requant16<EPI, ACT>(vb, s_bias, c0 + 16, p, ovf, valid);
                if (valid) *reinterpret_cast<uint4 *>(dst + c0 + 16) = w;
            }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty);
    } else {
        // raw accumulators -> smem (16-byte chunks XOR-swizzled by row to spread banks), 2x2 max, then requantise the
        // maxima only: requantisation is monotone, so max-then-requantise == requantise-then-max.
        const int chunks = p.N / 4;                        // 16-byte chunks per row (multiple of 8)
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            int v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int ch = (c0 / 4 + j) ^ (row & 7);
                *reinterpret_cast<int4 *>(s_stage + (size_t)row * p.N + 4 * ch) = make_int4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty);                           // TMEM buffer is free; the rest works from smem
        named_bar_sync(1, EPI_THREADS);
        const int PW = p.TW / 2, PH = p.TH / 2;
        const int pooled_px = p.TN * PH * PW;
        const int OH = p.H / 2, OW = p.W / 2;
        for (int item = et; item < pooled_px * chunks; item += EPI_THREADS) {
            const int cg = item % chunks, pp = item / chunks;
            const int pw = pp % PW, ph = (pp / PW) % PH, nl = pp / (PW * PH);
            const int r00 = (nl * p.TH + 2 * ph) * p.TW + 2 * pw;
            int4 m = make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = r00 + (k >> 1) * p.TW + (k & 1);
                const int4 t = *reinterpret_cast<const int4 *>(s_stage + (size_t)r * p.N + 4 * (cg ^ (r & 7)));
                m.x = max(m.x, t.x); m.y = max(m.y, t.y); m.z = max(m.z, t.z); m.w = max(m.w, t.w);
            }
            const int ox = x0 / 2 + pw, oy = y0 / 2 + ph, n = n0 + nl;
            if (ox < OW && oy < OH && n < p.n_img) {
                const int mv[4] = { m.x, m.y, m.z, m.w };
                *reinterpret_cast<unsigned *>(out + (((size_t)n * OH + oy) * OW + ox) * p.cs_out + 4 * cg) =
                    requant4<EPI, ACT>(mv, s_bias, 4 * cg, p, ovf, true);
            }
        }
        named_bar_sync(1, EPI_THREADS);                    // staging buffer is reused by the next tile
    }
}

template <int CB, int EPI>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const UmmaParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    constexpr uint32_t A_BOX = 128u * CB;                          // smem footprint of one A box (128 rows)
    const uint32_t B_BOX = (uint32_t)p.N * CB;
    constexpr uint32_t LAYOUT = CB == 128 ? 2u : CB == 64 ? 4u : CB == 32 ? 6u : 0u;
    constexpr uint32_t SBO = CB == 16 ? 128u : 8u * CB;            // 8-row group stride

    const uint32_t stage0 = base;
    int *s_stage = reinterpret_cast<int *>(base_ptr + p.off_stage);      // pooled epilogue staging [128][N] int32
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    // barrier layout: full[stages], empty[stages], tmem_full[2], tmem_empty[2], then the TMEM address slot
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (p.stages + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * p.stages + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * p.stages + 2 + b); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * p.stages + 4);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (2 * p.stages + 4));

    if (warp == 0 && lane == 0) {
        tmap_prefetch(&map_a);
        tmap_prefetch(&map_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), EPI_THREADS); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    for (int i = threadIdx.x; i < p.N * p.nslices; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;                 // |b| < 2^21: exact
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int stages_per_tile = (p.kblocks + p.G - 1) / p.G;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;   // global stage counter
            for (int work = blockIdx.x; work < p.num_tiles * p.nslices; work += gridDim.x) {
                const int tile = work % p.num_tiles, slice = work / p.num_tiles;
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tn = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
                for (int st = 0; st < stages_per_tile; ++st, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(bar_empty(s), ph ^ 1u);
                    const int kb0 = st * p.G, kb1 = min(kb0 + p.G, p.kblocks);
                    mbar_expect_tx(bar_full(s), (uint32_t)(kb1 - kb0) * (p.a_box_bytes + B_BOX));
                    const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
                    const uint32_t sb = sa + (uint32_t)p.G * A_BOX;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        // kblock -> (tap, channel chunk); K index of B = kb * CB bytes since K is tap-major over cs_in
                        int tap, c0;
                        if (CB == 128) { int per = p.kblocks / p.taps; tap = kb / per; c0 = (kb % per) * CB; }
                        else { tap = kb; c0 = 0; }
                        const int kh = p.taps == 1 ? 1 : tap / 3, kw = p.taps == 1 ? 1 : tap % 3;
                        tma_load_4d(sa + (uint32_t)(kb - kb0) * A_BOX, &map_a, bar_full(s), c0, x0 + kw - 1, y0 + kh - 1, n0);
                        // B: the TMA unit spends ~4 cycles per box ROW whatever its length, and the N rows of a weight block
                        // are 2/3 of all rows of a stage; when the host has laid the block out exactly as the 128B-swizzled
                        // shared-memory image, one 1-D bulk copy replaces the N-row box
                        if (CB == 128 && p.w_swz)
                            bulk_load_1d(sb + (uint32_t)(kb - kb0) * B_BOX, p.w_swz + ((size_t)kb * p.cs_out + (size_t)slice * p.N) * 128u, B_BOX, bar_full(s));
                        else tma_load_2d(sb + (uint32_t)(kb - kb0) * B_BOX, &map_b, bar_full(s), kb * CB, slice * p.N);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp converged, one elected lane issues (keeps the descriptor arithmetic in uniform registers)
        {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
            int it = 0, tcount = 0;
            for (int work = blockIdx.x; work < p.num_tiles * p.nslices; work += gridDim.x, ++tcount) {
                const int buf = tcount & 1;
                const uint32_t bph = (uint32_t)(tcount >> 1) & 1u;
                mbar_wait(bar_tempty(buf), bph ^ 1u);             // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * p.tmem_buf_stride;
                uint32_t accum = 0;
                for (int st = 0; st < stages_per_tile; ++st, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(bar_full(s), ph);
                    tc_fence_after();
                    const int kb0 = st * p.G, kb1 = min(kb0 + p.G, p.kblocks);
                    const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
                    const uint32_t sb = sa + (uint32_t)p.G * A_BOX;
                    if (elect_one()) {
                        if (CB == 16) {
                            // 16-byte channel vectors: one MMA (K = 32) covers two taps; the two 16-byte K halves are
                            // separate core-matrix columns LBO apart (the next tap's box).  Tap 9 does not exist: its B
                            // rows are zero (the host packs a 10th all-zero tap); its A box is whatever TMA fetched there.
                            for (int kb = kb0; kb < kb1; kb += 2) {
                                uint64_t ad = make_desc(sa + (uint32_t)(kb - kb0) * A_BOX, A_BOX, SBO, LAYOUT);
                                uint64_t bd = make_desc(sb + (uint32_t)(kb - kb0) * B_BOX, B_BOX, SBO, LAYOUT);
                                umma_i8(d, ad, bd, idesc, accum);
                                accum = 1;
                            }
                        } else {
                            uint64_t ad = make_desc(sa, 16, SBO, LAYOUT);
                            uint64_t bd = make_desc(sb, 16, SBO, LAYOUT);
                            for (int kb = kb0; kb < kb1; ++kb) {
#pragma unroll
                                for (int ks = 0; ks < CB / 32; ++ks) {
                                    umma_i8(d, ad + 2u * ks, bd + 2u * ks, idesc, accum);     // +32 bytes of K
                                    accum = 1;
                                }
                                ad += A_BOX >> 4; bd += B_BOX >> 4;
                            }
                        }
                        umma_commit(bar_empty(s));                 // smem slot free once these MMAs have read it
                        if (st == stages_per_tile - 1) umma_commit(bar_tfull(buf));   // accumulator complete
                    }
                    __syncwarp();
                    accum = 1;
                }
            }
        }
    } else {
        // ===================== epilogue warps (2..9) =====================
        const int ew = warp - 2;
        const int q4 = warp & 3;                                   // TMEM lane quarter this warp may access
        const int row = q4 * 32 + lane;                            // tile row = TMEM lane
        const int et = threadIdx.x - 64;                           // 0..255 within the epilogue group
        const int cmid = ((p.N / 16 + 1) / 2) * 16;                // column split between the two warps of a quarter
        const int cbeg = ew < 4 ? 0 : cmid, cend = ew < 4 ? cmid : p.N;
        unsigned ovf = 0;
        int tcount = 0;
        for (int work = blockIdx.x; work < p.num_tiles * p.nslices; work += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            const uint32_t bph = (uint32_t)(tcount >> 1) & 1u;
            const int tile = work % p.num_tiles, slice = work / p.num_tiles;
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, tn = tile / (p.tiles_x * p.tiles_y);
            const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
            mbar_wait(bar_tfull(buf), bph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * p.tmem_buf_stride + ((uint32_t)(q4 * 32) << 16);
            const int *sb_ = s_bias + slice * p.N;                 // this slice's channels: biases and output columns
            int8_t *out_ = p.out + slice * p.N;
            if (p.q.activ) epilogue_tile<EPI, true>(p, taddr, cbeg, cend, row, et, x0, y0, n0, sb_, s_stage, bar_tempty(buf), ovf, out_);
            else epilogue_tile<EPI, false>(p, taddr, cbeg, cend, row, et, x0, y0, n0, sb_, s_stage, bar_tempty(buf), ovf, out_);
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static cudaError_t get_encode()
{
    if (g_encode) return cudaSuccess;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    g_encode = (EncodeTiledFn)fn;
    return cudaSuccess;
}

static CUtensorMapSwizzle swizzle_for(int cb)
{
    return cb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : cb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : cb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// Pick the tile (TN, TH, TW), TN*TH*TW <= 128, that wastes the fewest MMA rows on this map.
static void pick_tile(int n, int H, int W, bool pool, int *TN, int *TH, int *TW)
{
    double best = -1; int bn = 1, bh = 1, bw = 1;
    const int step = pool ? 2 : 1;
    for (int tw = step; tw <= 128 && tw <= ((W + step - 1) / step) * step; tw += step)
        for (int th = step; th * tw <= 128 && th <= ((H + step - 1) / step) * step; th += step) {
            int maxn = 128 / (tw * th);
            for (int tn = 1; tn <= maxn && tn <= (n > 0 ? n : 1); ++tn) {
                long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((n + tn - 1) / tn);
                double eff = (double)n * H * W / (double)(tiles * 128);
                eff += 1e-6 * tw;                                  // prefer long contiguous rows
                if (eff > best) { best = eff; bn = tn; bh = th; bw = tw; }
            }
        }
    *TN = bn; *TH = bh; *TW = bw;
}

bool conv3x3_umma_supported(const ConvArgs &a)
{
    if (a.cs_in < 16 || a.cs_in % 16) return false;
    if (a.cs_in > 128 && a.cs_in % 128) return false;
    if (a.cs_in != 16 && a.cs_in != 32 && a.cs_in != 64 && a.cs_in % 128) return false;
    if (a.cs_out < 16 || a.cs_out % 16) return false;
    if (a.cs_out > 256 && (a.cs_out % 256 || a.q.pool)) return false;        // wide layers: slices of 256 channels, pool not fused
    if (a.taps != 9 && (a.taps != 1 || a.cs_in % 128 || !a.wgt1)) return false;   // 1x1: channel chunks of 128 bytes
    if (a.q.pool && (a.H < 2 || a.W < 2 || a.cs_out % 32)) return false;
    return true;
}

template <int CB, int EPI>
static cudaError_t launch_umma(const ConvArgs &a, cudaStream_t st, int sm_count, const EpiConst &kc)
{
    cudaError_t e = get_encode();
    if (e != cudaSuccess) return e;
    UmmaParams p;
    memset(&p, 0, sizeof p);
    p.n_img = a.n; p.H = a.H; p.W = a.W;
    pick_tile(a.n, a.H, a.W, a.q.pool != 0, &p.TN, &p.TH, &p.TW);
    p.tiles_x = (a.W + p.TW - 1) / p.TW; p.tiles_y = (a.H + p.TH - 1) / p.TH; p.tiles_n = (a.n + p.TN - 1) / p.TN;
    p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.N = a.cs_out > 256 ? 256 : a.cs_out; p.nslices = a.cs_out / p.N; p.cout = a.cout; p.cs_out = a.cs_out;
    p.taps = a.taps;
    p.kblocks = CB == 16 ? 10 : a.taps * (a.cs_in / CB);
    p.a_box_bytes = (uint32_t)(p.TN * p.TH * p.TW) * CB;
    const uint32_t kb_bytes = 128u * CB + (uint32_t)p.N * CB;
    if (CB == 16) p.G = 10;
    else if (CB == 128) p.G = 1;
    else p.G = 3;
    p.stage_bytes = ((uint32_t)p.G * kb_bytes + 1023u) & ~1023u;
    const uint32_t staging = a.q.pool ? 128u * p.N * 4u : 0u;
    const uint32_t fixed = staging + (uint32_t)a.cs_out * 4u + 256u + 1024u /*alignment slack*/;
    const uint32_t budget = 227u * 1024u;
    int stages = (int)((budget - fixed) / p.stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorInvalidConfiguration;
    p.stages = stages;
    p.off_stage = (uint32_t)stages * p.stage_bytes;
    p.off_bias = p.off_stage + staging;
    p.off_bar = (p.off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    const uint32_t smem_bytes = p.off_bar + 256u + 1024u;
    uint32_t nb = 32; while (nb < (uint32_t)p.N) nb <<= 1;
    p.tmem_buf_stride = nb; p.tmem_cols = 2 * nb;
    p.q = a.q; p.k = kc; p.bias_sh = a.bias_sh; p.out = a.out; p.ovf = a.ovf;
    p.w_swz = (CB == 128 && a.cs_out == a.wgt_swz_rows) ? a.wgt_swz : nullptr;      // (taps == 1: the one-tap image)

    CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n };
        cuuint64_t strides[3] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.cs_in * a.W, (cuuint64_t)a.cs_in * a.W * a.H };
        cuuint32_t box[4] = { (cuuint32_t)CB, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN };
        cuuint32_t es[4] = { 1, 1, 1, 1 };
        CUresult r = g_encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_for(CB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    {
        // weights [cout_pad][K], K = kblocks*CB bytes (CB = 16: a 10th all-zero tap is part of the packed buffer)
        const cuuint64_t K = (cuuint64_t)(CB == 16 ? 10 * 16 : a.taps * a.cs_in);
        cuuint64_t dims[2] = { K, (cuuint64_t)a.w_rows };
        cuuint64_t strides[1] = { K };
        cuuint32_t box[2] = { (cuuint32_t)CB, (cuuint32_t)p.N };
        cuuint32_t es[2] = { 1, 1 };
        CUresult r = g_encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)(CB == 16 ? a.wgt_k160 : a.taps == 1 ? a.wgt1 : a.wgt), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(CB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        e = cudaFuncSetAttribute(conv3x3_umma_kernel<CB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int work = p.num_tiles * p.nslices;
    int grid = work < sm_count ? work : sm_count;
    conv3x3_umma_kernel<CB, EPI><<<grid, UMMA_THREADS, smem_bytes, st>>>(map_a, map_b, p);
    return cudaGetLastError();
}

// ---- debug / property-test hook: the epilogue arithmetic alone on caller-supplied accumulators -------------------
// Runs the SAME requant4v the convolution epilogues run, four consecutive elements per thread.
template <int EPI>
__global__ void requant_probe_kernel(const int *__restrict__ acc, size_t count, int cout, const int *__restrict__ bias_sh,
                                     UmmaParams p, int8_t *__restrict__ out)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= count) return;
    int a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const size_t i = i0 + j < count ? i0 + j : count - 1;
        a[j] = acc[i];
        const int bb = bias_sh[(int)(i % (size_t)cout)];
        b[j] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)bb) : bb;
    }
    unsigned ovf = 0;
    const unsigned w = p.q.activ ? requant4v<EPI, true>(a, make_int4(b[0], b[1], b[2], b[3]), p, ovf, true)
                                 : requant4v<EPI, false>(a, make_int4(b[0], b[1], b[2], b[3]), p, ovf, true);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (i0 + j < count) out[i0 + j] = (int8_t)((w >> (8 * j)) & 0xff);
}

cudaError_t requant_probe(const ConvArgs &a, const int *acc, size_t count, int8_t *out, int *epi_used, cudaStream_t st)
{
    UmmaParams p;
    memset(&p, 0, sizeof p);
    p.q = a.q;
    const int epi = epi_mode_for(a, &p.k);
    if (epi_used) *epi_used = epi;
    if (count == 0) return cudaSuccess;
    const int blocks = (int)((count + 1023) / 1024);
    if (epi == EPI_F_RNE) requant_probe_kernel<EPI_F_RNE><<<blocks, 256, 0, st>>>(acc, count, a.cout, a.bias_sh, p, out);
    else if (epi == EPI_F_RNE_NOHI) requant_probe_kernel<EPI_F_RNE_NOHI><<<blocks, 256, 0, st>>>(acc, count, a.cout, a.bias_sh, p, out);
    else if (epi == EPI_P) requant_probe_kernel<EPI_P><<<blocks, 256, 0, st>>>(acc, count, a.cout, a.bias_sh, p, out);
    else requant_probe_kernel<EPI_GENERIC><<<blocks, 256, 0, st>>>(acc, count, a.cout, a.bias_sh, p, out);
    return cudaGetLastError();
}

template <int CB>
static cudaError_t launch_epi(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    EpiConst kc;
    switch (epi_mode_for(a, &kc)) {
    case EPI_F_RNE: return launch_umma<CB, EPI_F_RNE>(a, st, sm_count, kc);
    case EPI_F_RNE_NOHI: return launch_umma<CB, EPI_F_RNE_NOHI>(a, st, sm_count, kc);
    case EPI_P:     return launch_umma<CB, EPI_P>(a, st, sm_count, kc);
    default:        return launch_umma<CB, EPI_GENERIC>(a, st, sm_count, kc);
    }
}

cudaError_t conv3x3_umma(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    if (a.cs_in == 16) return launch_epi<16>(a, st, sm_count);
    if (a.cs_in == 32) return launch_epi<32>(a, st, sm_count);
    if (a.cs_in == 64) return launch_epi<64>(a, st, sm_count);
    return launch_epi<128>(a, st, sm_count);
}

}  // namespace yb
