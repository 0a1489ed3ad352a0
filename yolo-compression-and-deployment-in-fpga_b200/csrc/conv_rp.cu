// conv_rp.cu — "row-pair" tcgen05 kernels for the thin pooled layers at the top of the network (slim_yolo_v2 conv2:
// 16 -> 32 channels + 2x2 max-pool, second_conv, c_embedding/yolo_forward.c:420-573 / models/slim_yolo_v2.py:232-243).
//
// These layers have so little K (144) that an M=128 tensor-core instruction is bound by its fixed cost (~43-51 cycles
// whatever N and K are, tools/micro/umma_issue.cu), and the weight-stationary kernel (conv_ws.cu) additionally pays 612
// 16-byte cp.async pieces per tile to split the halo tile by x parity for the pooled accumulators.  Here:
//
//   * the GEMM N dimension carries TWO output rows: column dy*Cout + co is output channel co of row 2Y + dy.  The K
//     dimension then spans the FOUR input rows 2Y-1 .. 2Y+2 (12 chunks of one tap x 16 channels = 6 K-steps of 32); the
//     weight image has zero blocks where kh = khh - dy falls outside 0..2.  One MMA sequence of 6 instructions at N = 64
//     (51 cycles each) covers 2 x 128 output pixels where the phase-decomposed tile needed 10 at N = 32 (43 cycles each),
//     and the VERTICAL half of the max-pool becomes a maximum over two column ranges of the same TMEM lane;
//   * the M rows are 128 x-consecutive pixels of one row pair, so the halo tile is a DENSE pixel-major block
//     [4 input rows][segw + 2 pixels][16 B]: in the UMMA no-swizzle K-major layout 8 consecutive pixels are one core
//     matrix, every tap is a start offset, a tap pair is (start, LBO).  One TMA box per tile writes it: the tensor map
//     declares a row of the NHWC16 map as W*2 64-bit elements, so a box row is one contiguous run of (segw + 2) * 16
//     bytes instead of segw + 2 runs of 16 (the TMA engine costs ~4 cycles per innermost run, profiles/README.md); out of
//     bounds rows / columns arrive as zeros = the convolution's padding;
//   * the HORIZONTAL half of the pool pairs TMEM lanes 2i and 2i+1: after the vertical maximum each lane keeps one half
//     of the channels and sends the other half to its partner (16 shuffles), so the even lane requantises and stores
//     channels 0-15 and the odd lane channels 16-31 of the pooled pixel: a warp writes 512 contiguous bytes.
//     Requantisation is monotone, so max-then-requantise == requantise-then-max (slim_yolo_v2.py:229-231).
//
// Warp roles (640 threads): warps 0-1 = MMA issuers (alternate tiles; warp 0 also allocates TMEM), warps 2-3 = TMA producers (one lane
// each, alternate tiles), warps 4-19 = four epilogue groups of four warps (one per TMEM lane quarter); tile i of a CTA uses accumulator buffer
// i % 8 (8 x 64 columns = all of TMEM) and epilogue group i % 4, so four epilogues are in flight behind the MMA warp.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <cstdio>
#include <cstdlib>

namespace yb {

#ifdef YB_WS_TIMELINE
#define RP_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 64 && lane == 0) p.dbg[it * 8 + (slot)] = clock64(); } while (0)
#else
#define RP_STAMP(slot) do { } while (0)
#endif

constexpr int RP_EPI_GROUPS = 4;
constexpr int RP_THREADS = 128 + RP_EPI_GROUPS * 128;      // 640
constexpr int RP_STAGES = 8;
constexpr int RP_TBUF = 8;
constexpr int RP_N = 64;                                   // 2 output rows x 32 channels
constexpr int RP_KC = 12;                                  // 16-byte K chunks per GEMM column: 4 input rows x 3 taps

struct RpParams {
    int n_img, H, W, OH, OW;
    int nseg, segw;              // x segments per row pair; positions per segment (even, <= 126)
    unsigned nseg_magic, oh_magic;   // ceil(2^32 / nseg), ceil(2^32 / OH)
    int num_tiles;               // n_img * OH * nseg
    uint32_t plane_bytes;        // (segw + 2) * 16: one input row of the halo tile
    uint32_t stage_bytes;
    uint32_t w_bytes;
    uint32_t off_stage, off_bias, off_bar;
    int cs_out;                  // 32
    LayerQ q;
    EpiConst k;
    const uint8_t *wimg;         // [N/8][12][8][16 B]
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    long long *dbg;              // YB_WS_TIMELINE builds: clock64 stamps of CTA 0's first 64 tiles
};

template <int EPI, bool ACT>
__device__ __forceinline__ void rp16_epilogue_tile(const RpParams &p, uint32_t taddr, int lane, int q4, int img, int Y, int x0,
                                                   const int *s_bias, uint32_t bar_tempty, unsigned &ovf)
{
    const int m = q4 * 32 + lane;                              // TMEM lane = position x0 + m of the row pair
    // .x16 loads: 16 warps sustain ~200 B/clk/SM of TMEM reads with them, ~120 B/clk with .x32 (tools/micro/ldtm_bench.cu)
    int va[16], vb[16], vc[16], vd[16], v[64];
#if defined(YB_RP_EXP) && YB_RP_EXP == 1      // experiment: no TMEM loads (wrong results): does the MMA rate recover?
#pragma unroll
    for (int j = 0; j < 16; ++j) { va[j] = lane + j; vb[j] = lane - j; vc[j] = j; vd[j] = -j; }
#else
    tmem_ld16(taddr, va);                                      // row 2Y:     channels 0..31
    tmem_ld16(taddr + 16, vb);
    tmem_ld16(taddr + 32, vc);                                 // row 2Y + 1
    tmem_ld16(taddr + 48, vd);
    tmem_ld_wait();
#endif
#pragma unroll
    for (int j = 0; j < 16; ++j) { v[j] = va[j]; v[16 + j] = vb[j]; v[32 + j] = vc[j]; v[48 + j] = vd[j]; }
    tc_fence_before();
    mbar_arrive(bar_tempty);                                   // the accumulators are in registers: the buffer is free
#if defined(YB_RP_EXP) && YB_RP_EXP == 2      // experiment: loads but (almost) no arithmetic
    {
        int x = 0;
#pragma unroll
        for (int j = 0; j < 64; ++j) x ^= v[j];
        if (x == 0x12345) *reinterpret_cast<int *>(p.out) = x;
        return;
    }
#endif
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = max(v[j], v[32 + j]);
    const bool odd = lane & 1;
    int w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int send = odd ? v[j] : v[16 + j];
        const int keep = odd ? v[16 + j] : v[j];
        w[j] = max(keep, __shfl_xor_sync(0xffffffffu, send, 1));
    }
    const int x = x0 + m;
    const bool valid = m < p.segw && (x | 1) < p.W;            // both lanes of a pair agree
    const int ox = x >> 1, c0 = odd ? 16 : 0;
    const uint4 o = requant16<EPI, ACT>(w, s_bias, c0, p, ovf, valid);
    if (valid) *reinterpret_cast<uint4 *>(p.out + (((size_t)img * p.OH + Y) * p.OW + ox) * p.cs_out + c0) = o;
}

template <int EPI>
__global__ void __launch_bounds__(RP_THREADS, 1) conv3x3_rp16_kernel(const RpParams p, const __grid_constant__ CUtensorMap map)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t wsm = base;
    const uint32_t stage0 = base + p.off_stage;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (RP_STAGES + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * RP_STAGES + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * RP_STAGES + RP_TBUF + b); };
    const uint32_t bar_w = bar0 + 8u * (2 * RP_STAGES + 2 * RP_TBUF);
    const uint32_t tmem_slot = bar_w + 8u;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (2 * RP_STAGES + 2 * RP_TBUF + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < RP_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < RP_TBUF; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 128); }
        mbar_init(bar_w, 1);
        fence_barrier_init();
        mbar_expect_tx(bar_w, p.w_bytes);
        bulk_load_1d(wsm, p.wimg, p.w_bytes, bar_w);
        tmap_prefetch(&map);
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = threadIdx.x; i < p.cs_out; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp < 2) {
        // ===================== MMA issuers =====================
        // Two warps take alternate tiles.  A single issuing warp spends ~300 cycles per tile on its own serial overhead
        // (barrier polls, descriptor arithmetic: a lone warp issues one dependent instruction every ~5 cycles) on top of
        // the ~370 cycles its six MMAs occupy the tensor pipe (clock64 timeline, tools/ws_timeline.py); with two warps
        // one's overhead hides behind the other's MMAs.  Tiles are independent accumulation chains in different TMEM
        // buffers, so their order in the pipe is free; a commit tracks the issuing thread's own MMAs.
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RP_N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t bhi = ((uint32_t)RP_KC * 8u) | (1u << 14);       // SBO = 12 chunks x 128 B between 8-column groups
        const uint32_t ahi = 8u | (1u << 14);                           // SBO = 128 B: 8-pixel groups are contiguous
        const uint32_t b16 = wsm >> 4, ps16 = p.plane_bytes >> 4;
        mbar_wait(bar_w, 0);
        for (int it = warp, tile = blockIdx.x + warp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int buf = it & (RP_TBUF - 1), s = it & (RP_STAGES - 1);
            const uint32_t bph = (uint32_t)(it >> 3) & 1u;              // RP_TBUF == RP_STAGES == 8: one phase bit serves both rings
            mbar_wait(bar_tempty(buf), bph ^ 1u);
            RP_STAMP(0);
            mbar_wait(bar_full(s), bph);
            RP_STAMP(1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)buf * RP_N;
                const uint32_t sa16 = (stage0 + (uint32_t)s * p.stage_bytes) >> 4;
                // K-step j = chunks 2j, 2j+1 of (input row khh, tap kw): A start = plane khh + kw pixels, LBO = distance to the second chunk
                umma_i8_lohi<false>(d, (sa16) | (1u << 16), ahi, (b16) | (8u << 16), bhi, idesc);                                   // (0,0) (0,1)
                umma_i8_lohi<true>(d, (sa16 + 2u) | ((ps16 - 2u) << 16), ahi, (b16 + 16u) | (8u << 16), bhi, idesc);               // (0,2) (1,0)
                umma_i8_lohi<true>(d, (sa16 + ps16 + 1u) | (1u << 16), ahi, (b16 + 32u) | (8u << 16), bhi, idesc);                 // (1,1) (1,2)
                umma_i8_lohi<true>(d, (sa16 + 2u * ps16) | (1u << 16), ahi, (b16 + 48u) | (8u << 16), bhi, idesc);                 // (2,0) (2,1)
                umma_i8_lohi<true>(d, (sa16 + 2u * ps16 + 2u) | ((ps16 - 2u) << 16), ahi, (b16 + 64u) | (8u << 16), bhi, idesc);   // (2,2) (3,0)
                umma_i8_lohi<true>(d, (sa16 + 3u * ps16 + 1u) | (1u << 16), ahi, (b16 + 80u) | (8u << 16), bhi, idesc);            // (3,1) (3,2)
                umma_commit(bar_empty(s));
                umma_commit(bar_tfull(buf));
            }
            __syncwarp();
            RP_STAMP(2);
        }
    } else if (warp < 4) {
        // ===================== TMA producers: two warps (one lane each) take alternate tiles =====================
        if (lane == 0) {
            for (int it = warp - 2, tile = blockIdx.x + (warp - 2) * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
                const int s = it & (RP_STAGES - 1);
                const uint32_t ph = (uint32_t)(it >> 3) & 1u;
                const int r = p.nseg == 1 ? tile : (int)__umulhi((unsigned)tile, p.nseg_magic), seg = tile - r * p.nseg;
                const int img = p.OH == 1 ? r : (int)__umulhi((unsigned)r, p.oh_magic), Y = r - img * p.OH;
                mbar_wait(bar_empty(s), ph ^ 1u);
                RP_STAMP(3);
                mbar_expect_tx(bar_full(s), 4u * p.plane_bytes);
                tma_load_3d(stage0 + (uint32_t)s * p.stage_bytes, &map, bar_full(s), 2 * (seg * p.segw - 1), 2 * Y - 1, img);
                RP_STAMP(4);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int ew = warp - 4, grp = ew >> 2, q4 = warp & 3;    // TMEM lane quarter = warp id % 4
        unsigned ovf = 0;
        for (int it = grp, tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += RP_EPI_GROUPS * gridDim.x, it += RP_EPI_GROUPS) {
            const int buf = it & (RP_TBUF - 1);
            const uint32_t bph = (uint32_t)(it >> 3) & 1u;
            const int r = p.nseg == 1 ? tile : (int)__umulhi((unsigned)tile, p.nseg_magic), seg = tile - r * p.nseg;
            const int img = p.OH == 1 ? r : (int)__umulhi((unsigned)r, p.oh_magic), Y = r - img * p.OH;
            mbar_wait(bar_tfull(buf), bph);
            if (q4 == 0) RP_STAMP(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * RP_N + ((uint32_t)(q4 * 32) << 16);
            if (p.q.activ) rp16_epilogue_tile<EPI, true>(p, taddr, lane, q4, img, Y, seg * p.segw, s_bias, bar_tempty(buf), ovf);
            else rp16_epilogue_tile<EPI, false>(p, taddr, lane, q4, img, Y, seg * p.segw, s_bias, bar_tempty(buf), ovf);
            if (q4 == 0) RP_STAMP(7);
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// x-split variant: the producing layer (conv_first.cu, out_xsplit) stores every row of the 16-channel input map as
// [even pixels][odd pixels].  Pixels at stride 2 are then contiguous, so the accumulators can be split by the x parity of
// the OUTPUT pixel like the phase-decomposed tiles of conv_ws.cu, but from two dense TMA boxes per tile instead of 612
// cp.async pieces: TMEM lane m of accumulator E / O holds pre-pool pixel 2m / 2m+1 of both rows, all four members of pooled
// pixel m sit in one lane, and the pool is two integer maxima per channel: no shuffles, no selects.  (The shuffle variant
// above spends 8.5 ALU-pipe operations per output value and is bound by that pipe, ncu: alu 63 %; this one 5.5.)
//   smem stage: O block [4 input rows][PB pixels][16 B] (odd input pixels 2(px0-1)+1 ...), then E block (even pixels 2 px0 ...)
//   E accumulator (output x = 2m):   taps kw 0 / 1 / 2 = O(r, m) / E(r, m) / O(r, m+1)
//   O accumulator (output x = 2m+1): taps kw 0 / 1 / 2 = E(r, m) / O(r, m+1) / E(r, m+1)
//   K chunk order (one weight image serves both): chunks 2r, 2r+1 = (row r, kw 0), (row r, kw 2): same block, LBO = 16 B;
//   chunks 8 + r = (row r, kw 1): rows r, r+1 pair up with LBO = one block row.
constexpr int RS_TBUF = 4;                                 // 4 x 128 TMEM columns
constexpr int RS_STAGES = 8;

struct RsParams {
    int n_img, H, W, OH, OW;     // input map H x W (W even), output OH x OW = H/2 x W/2
    int nseg, segp;              // pooled-pixel segments per row pair; pooled pixels per segment (<= 127)
    unsigned nseg_magic, oh_magic;
    int num_tiles;
    uint32_t pb16;               // PB = segp + 1 pixels per box row, in 16-byte units (= pixels)
    uint32_t blk16;              // one block (4 box rows) rounded up to 128 bytes (TMA destinations are 128-byte aligned), in 16-byte units
    uint32_t stage_bytes;
    uint32_t w_bytes;
    uint32_t off_stage, off_bias, off_bar;
    int cs_out;
    LayerQ q;
    EpiConst k;
    const uint8_t *wimg;
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    long long *dbg;
};

template <int EPI, bool ACT>
__device__ __forceinline__ void rs16_epilogue_tile(const RsParams &p, uint32_t taddr, int lane, int q4, int img, int Y, int px0,
                                                   const int *s_bias, uint32_t bar_tempty, unsigned &ovf)
{
    const int m = q4 * 32 + lane;                              // TMEM lane = pooled pixel px0 + m
    const int ox = px0 + m;
    const bool valid = m < p.segp && ox < p.OW;
    int8_t *dst = p.out + (((size_t)img * p.OH + Y) * p.OW + ox) * p.cs_out;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                              // channels 16h .. 16h + 15
        int e0[16], e1[16], o0[16], o1[16];
        tmem_ld16(taddr + 16 * h, e0);                         // E, row 2Y
        tmem_ld16(taddr + 32 + 16 * h, e1);                    // E, row 2Y + 1
        tmem_ld16(taddr + 64 + 16 * h, o0);                    // O, row 2Y
        tmem_ld16(taddr + 96 + 16 * h, o1);                    // O, row 2Y + 1
        tmem_ld_wait();
        if (h == 1) { tc_fence_before(); mbar_arrive(bar_tempty); }      // all of the tile's accumulators are in registers
#pragma unroll
        for (int j = 0; j < 16; ++j) e0[j] = max(max(e0[j], e1[j]), max(o0[j], o1[j]));
        const uint4 w = requant16<EPI, ACT>(e0, s_bias, 16 * h, p, ovf, valid);
        if (valid) *reinterpret_cast<uint4 *>(dst + 16 * h) = w;
    }
}

template <int EPI>
__global__ void __launch_bounds__(RP_THREADS, 1) conv3x3_rs16_kernel(const RsParams p, const __grid_constant__ CUtensorMap map)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t wsm = base;
    const uint32_t stage0 = base + p.off_stage;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (RS_STAGES + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * RS_STAGES + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * RS_STAGES + RS_TBUF + b); };
    const uint32_t bar_w = bar0 + 8u * (2 * RS_STAGES + 2 * RS_TBUF);
    const uint32_t tmem_slot = bar_w + 8u;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (2 * RS_STAGES + 2 * RS_TBUF + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < RS_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < RS_TBUF; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), 128); }
        mbar_init(bar_w, 1);
        fence_barrier_init();
        mbar_expect_tx(bar_w, p.w_bytes);
        bulk_load_1d(wsm, p.wimg, p.w_bytes, bar_w);
        tmap_prefetch(&map);
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    for (int i = threadIdx.x; i < p.cs_out; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp < 2) {
        // ===================== MMA issuers: alternate tiles (see conv3x3_rp16_kernel) =====================
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RP_N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t bhi = ((uint32_t)RP_KC * 8u) | (1u << 14);
        const uint32_t ahi = 8u | (1u << 14);
        const uint32_t b16 = wsm >> 4, pb = p.pb16;
        const uint32_t lb1 = 1u << 16, lbr = pb << 16, blb = 8u << 16;
        mbar_wait(bar_w, 0);
        for (int it = warp, tile = blockIdx.x + warp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int buf = it & (RS_TBUF - 1), s = it & (RS_STAGES - 1);
            mbar_wait(bar_tempty(buf), ((uint32_t)(it >> 2) & 1u) ^ 1u);
            RP_STAMP(0);
            mbar_wait(bar_full(s), (uint32_t)(it >> 3) & 1u);
            RP_STAMP(1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t dE = tmem_base + (uint32_t)buf * 128u, dO = dE + 64u;
                const uint32_t o16 = (stage0 + (uint32_t)s * p.stage_bytes) >> 4, e16 = o16 + p.blk16;   // O block, E block
                // E accumulator: (kw 0, kw 2) of row r = O(r, 0), O(r, 1); kw 1 of rows (0,1), (2,3) = E(r, 0)
                umma_i8_lohi<false>(dE, (o16) | lb1, ahi, (b16) | blb, bhi, idesc);
                umma_i8_lohi<true>(dE, (o16 + pb) | lb1, ahi, (b16 + 16u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dE, (o16 + 2u * pb) | lb1, ahi, (b16 + 32u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dE, (o16 + 3u * pb) | lb1, ahi, (b16 + 48u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dE, (e16) | lbr, ahi, (b16 + 64u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dE, (e16 + 2u * pb) | lbr, ahi, (b16 + 80u) | blb, bhi, idesc);
                // O accumulator: (kw 0, kw 2) of row r = E(r, 0), E(r, 1); kw 1 = O(r, 1)
                umma_i8_lohi<false>(dO, (e16) | lb1, ahi, (b16) | blb, bhi, idesc);
                umma_i8_lohi<true>(dO, (e16 + pb) | lb1, ahi, (b16 + 16u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dO, (e16 + 2u * pb) | lb1, ahi, (b16 + 32u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dO, (e16 + 3u * pb) | lb1, ahi, (b16 + 48u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dO, (o16 + 1u) | lbr, ahi, (b16 + 64u) | blb, bhi, idesc);
                umma_i8_lohi<true>(dO, (o16 + 2u * pb + 1u) | lbr, ahi, (b16 + 80u) | blb, bhi, idesc);
                umma_commit(bar_empty(s));
                umma_commit(bar_tfull(buf));
            }
            __syncwarp();
            RP_STAMP(2);
        }
    } else if (warp < 4) {
        // ===================== TMA producers: two warps (one lane each) take alternate tiles =====================
        if (lane == 0) {
            const uint32_t box_bytes = 4u * p.pb16 * 16u;
            for (int it = warp - 2, tile = blockIdx.x + (warp - 2) * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
                const int s = it & (RS_STAGES - 1);
                const uint32_t ph = (uint32_t)(it >> 3) & 1u;
                const int r = p.nseg == 1 ? tile : (int)__umulhi((unsigned)tile, p.nseg_magic), seg = tile - r * p.nseg;
                const int img = p.OH == 1 ? r : (int)__umulhi((unsigned)r, p.oh_magic), Y = r - img * p.OH;
                const int px0 = seg * p.segp;
                mbar_wait(bar_empty(s), ph ^ 1u);
                RP_STAMP(3);
                mbar_expect_tx(bar_full(s), 2u * box_bytes);
                const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
                tma_load_4d(sa, &map, bar_full(s), 2 * (px0 - 1), 1, 2 * Y - 1, img);           // odd input pixels 2(px0-1)+1, ...
                tma_load_4d(sa + p.blk16 * 16u, &map, bar_full(s), 2 * px0, 0, 2 * Y - 1, img); // even input pixels 2 px0, ...
                RP_STAMP(4);
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 4, grp = ew >> 2, q4 = warp & 3;
        unsigned ovf = 0;
        for (int it = grp, tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += RP_EPI_GROUPS * gridDim.x, it += RP_EPI_GROUPS) {
            const int buf = it & (RS_TBUF - 1);
            const uint32_t bph = (uint32_t)(it >> 2) & 1u;
            const int r = p.nseg == 1 ? tile : (int)__umulhi((unsigned)tile, p.nseg_magic), seg = tile - r * p.nseg;
            const int img = p.OH == 1 ? r : (int)__umulhi((unsigned)r, p.oh_magic), Y = r - img * p.OH;
            mbar_wait(bar_tfull(buf), bph);
            if (q4 == 0) RP_STAMP(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * 128u + ((uint32_t)(q4 * 32) << 16);
            if (p.q.activ) rs16_epilogue_tile<EPI, true>(p, taddr, lane, q4, img, Y, seg * p.segp, s_bias, bar_tempty(buf), ovf);
            else rs16_epilogue_tile<EPI, false>(p, taddr, lane, q4, img, Y, seg * p.segp, s_bias, bar_tempty(buf), ovf);
            if (q4 == 0) RP_STAMP(7);
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static bool rp_enabled()
{
    static const bool on = [] { const char *e = getenv("YOLO_B200_RP"); return e ? atoi(e) != 0 : true; }();
    return on;
}

static unsigned rp_magic(int d) { return (unsigned)(((1ull << 32) + (unsigned)d - 1) / (unsigned)d); }

bool conv3x3_rp_supported(const ConvArgs &a)
{
    if (!rp_enabled() || !a.wimg_rp) return false;
    if (a.cs_in != 16 || a.cs_out != 32 || !a.q.pool || a.H < 2 || a.W < 2) return false;
    if ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) != 0) return false;
    const long long oh = a.H / 2, nseg = (a.W + 125) / 126;
    if ((long long)a.n * oh * nseg * (oh > nseg ? oh : nseg) >= (1ll << 32)) return false;      // multiply-high divisions stay exact
    if ((long long)a.n * oh * nseg >= (1ll << 31)) return false;
    return true;
}

typedef CUresult (*RpEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static RpEncodeTiledFn rp_encoder()
{
    static RpEncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
        if (qres != cudaDriverEntryPointSuccess || !fn) return nullptr;
        enc = (RpEncodeTiledFn)fn;
    }
    return enc;
}

template <int EPI>
static cudaError_t launch_rp16(RpParams &p, const CUtensorMap &map, cudaStream_t st, int sm_count)
{
#ifdef YB_WS_TIMELINE
    {
        static long long *dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 64 * 8 * sizeof(long long), st);
        p.dbg = dbg;
    }
#endif
    const uint32_t smem_bytes = p.off_bar + 8u * (2 * RP_STAGES + 2 * RP_TBUF + 2) + 128u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_rp16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    { cudaError_t le = launch_pdl(conv3x3_rp16_kernel<EPI>, dim3(grid), dim3(RP_THREADS), smem_bytes, st, p, map); if (le != cudaSuccess) return le; }
#ifdef YB_WS_TIMELINE
    {
        long long h[64 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.dbg, sizeof h, cudaMemcpyDeviceToHost);
        const long long t0 = h[3] ? h[3] : h[0];
        printf("RP timeline segw=%d nseg=%d tiles=%d (cycles since first stamp): tile | mma: tempty_ok full_ok issued | prod: empty_ok issued | epi: tfull_ok done\n", p.segw, p.nseg, p.num_tiles);
        for (int i = 0; i < 32; ++i)
            printf("  %2d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0, h[i*8+6]-t0, h[i*8+7]-t0);
    }
#endif
    return cudaGetLastError();
}


bool conv3x3_rp_split_supported(const ConvArgs &a)
{
    if (!rp_enabled() || !a.wimg_rps) return false;
    if (a.cs_in != 16 || a.cs_out != 32 || !a.q.pool || a.H < 2 || a.W < 2 || (a.W & 1)) return false;
    if ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) != 0) return false;
    const long long oh = a.H / 2, nseg = (a.W / 2 + 126) / 127;
    if ((long long)a.n * oh * nseg * (oh > nseg ? oh : nseg) >= (1ll << 32)) return false;
    if ((long long)a.n * oh * nseg >= (1ll << 31)) return false;
    return true;
}

template <int EPI>
static cudaError_t launch_rs16(RsParams &p, const CUtensorMap &map, cudaStream_t st, int sm_count)
{
#ifdef YB_WS_TIMELINE
    {
        static long long *dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 64 * 8 * sizeof(long long), st);
        p.dbg = dbg;
    }
#endif
    const uint32_t smem_bytes = p.off_bar + 8u * (2 * RS_STAGES + 2 * RS_TBUF + 2) + 128u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_rs16_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    { cudaError_t le = launch_pdl(conv3x3_rs16_kernel<EPI>, dim3(grid), dim3(RP_THREADS), smem_bytes, st, p, map); if (le != cudaSuccess) return le; }
#ifdef YB_WS_TIMELINE
    {
        long long h[64 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.dbg, sizeof h, cudaMemcpyDeviceToHost);
        const long long t0 = h[3] ? h[3] : h[0];
        printf("RS timeline segp=%d nseg=%d tiles=%d (cycles since first stamp): tile | mma: tempty_ok full_ok issued | prod: empty_ok issued | epi: tfull_ok done\n", p.segp, p.nseg, p.num_tiles);
        for (int i = 0; i < 32; ++i)
            printf("  %2d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0, h[i*8+6]-t0, h[i*8+7]-t0);
    }
#endif
    return cudaGetLastError();
}

static cudaError_t conv3x3_rp_split(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (!conv3x3_rp_split_supported(a)) return cudaErrorInvalidValue;
    RsParams p;
    memset(&p, 0, sizeof p);
    p.n_img = a.n; p.H = a.H; p.W = a.W; p.OH = a.H / 2; p.OW = a.W / 2;
    p.nseg = (p.OW + 126) / 127;
    p.segp = (p.OW + p.nseg - 1) / p.nseg;
    p.nseg_magic = rp_magic(p.nseg); p.oh_magic = rp_magic(p.OH);
    p.num_tiles = a.n * p.OH * p.nseg;
    p.pb16 = (uint32_t)p.segp + 1u;
    p.blk16 = (4u * p.pb16 + 7u) & ~7u;
    // an M = 128 instruction reads 128 + 1 pixels from its start whatever segp is: the tail of the last block row runs into slack
    p.stage_bytes = (2u * p.blk16 * 16u + 132u * 16u + 127u) & ~127u;
    p.w_bytes = (uint32_t)RP_N * RP_KC * 16u;
    p.off_stage = (p.w_bytes + 127u) & ~127u;
    p.off_bias = p.off_stage + (uint32_t)RS_STAGES * p.stage_bytes;
    p.off_bar = (p.off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    p.cs_out = a.cs_out; p.q = a.q; p.wimg = a.wimg_rps; p.bias_sh = a.bias_sh; p.out = a.out; p.ovf = a.ovf;

    RpEncodeTiledFn enc = rp_encoder();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap map;
    // (one x-parity half of a row as W/2 * 2 64-bit elements, parity, row, image)
    const cuuint64_t half = (cuuint64_t)(a.W / 2);
    cuuint64_t dims[4] = { half * 2, 2, (cuuint64_t)a.H, (cuuint64_t)a.n };
    cuuint64_t strides[3] = { half * 16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * 16 * a.H };
    cuuint32_t box[4] = { p.pb16 * 2, 1, 4, 1 };
    cuuint32_t es[4] = { 1, 1, 1, 1 };
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE:      return launch_rs16<EPI_F_RNE>(p, map, st, sm_count);
    case EPI_F_RNE_NOHI: return launch_rs16<EPI_F_RNE_NOHI>(p, map, st, sm_count);
    case EPI_P:          return launch_rs16<EPI_P>(p, map, st, sm_count);
    default:             return launch_rs16<EPI_GENERIC>(p, map, st, sm_count);
    }
}

cudaError_t conv3x3_rp(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    if (a.in_xsplit) return conv3x3_rp_split(a, st, sm_count);
    if (!conv3x3_rp_supported(a)) return cudaErrorInvalidValue;
    RpParams p;
    memset(&p, 0, sizeof p);
    p.n_img = a.n; p.H = a.H; p.W = a.W; p.OH = a.H / 2; p.OW = a.W / 2;
    p.nseg = (a.W + 125) / 126;
    p.segw = ((a.W + p.nseg - 1) / p.nseg + 1) & ~1;
    p.nseg_magic = rp_magic(p.nseg); p.oh_magic = rp_magic(p.OH);
    p.num_tiles = a.n * p.OH * p.nseg;
    p.plane_bytes = (uint32_t)(p.segw + 2) * 16u;
    // an M = 128 instruction reads 128 + 2 pixels from its start whatever segw is: the tail of the last plane runs into slack
    p.stage_bytes = (3u * p.plane_bytes + 132u * 16u + 127u) & ~127u;
    p.w_bytes = (uint32_t)RP_N * RP_KC * 16u;
    p.off_stage = (p.w_bytes + 127u) & ~127u;
    p.off_bias = p.off_stage + (uint32_t)RP_STAGES * p.stage_bytes;
    p.off_bar = (p.off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    p.cs_out = a.cs_out; p.q = a.q; p.wimg = a.wimg_rp; p.bias_sh = a.bias_sh; p.out = a.out; p.ovf = a.ovf;

    RpEncodeTiledFn enc = rp_encoder();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap map;
    // a row of the NHWC16 map as W * 2 64-bit elements: a box row is ONE contiguous run of (segw + 2) * 16 bytes
    cuuint64_t dims[3] = { (cuuint64_t)a.W * 2, (cuuint64_t)a.H, (cuuint64_t)a.n };
    cuuint64_t strides[2] = { (cuuint64_t)a.W * 16, (cuuint64_t)a.W * 16 * a.H };
    cuuint32_t box[3] = { (cuuint32_t)(p.segw + 2) * 2, 4, 1 };
    cuuint32_t es[3] = { 1, 1, 1 };
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;

    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE:      return launch_rp16<EPI_F_RNE>(p, map, st, sm_count);
    case EPI_F_RNE_NOHI: return launch_rp16<EPI_F_RNE_NOHI>(p, map, st, sm_count);
    case EPI_P:          return launch_rp16<EPI_P>(p, map, st, sm_count);
    default:             return launch_rp16<EPI_GENERIC>(p, map, st, sm_count);
    }
}

}  // namespace yb
