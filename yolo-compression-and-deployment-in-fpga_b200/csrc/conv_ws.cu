// conv_ws.cu — weight-stationary 3x3/stride-1/pad-1 int8 NHWC convolution on the 5th-generation tensor cores.
//
// Replaces second_conv/conv_normal/conv_last (c_embedding/yolo_forward.c:420,575,772) for every layer whose packed
// weights fit in shared memory (slim_yolo_v2: conv2 .. conv4_2 and pred).  Where the C driver re-sends the 3x3 weights
// and a haloed 18x22 input tile to the FPGA for every 16x20 output tile (load_weight/psram_to_inpBuf,
// yolo_forward.c:125-173), this kernel keeps the WHOLE layer's weights resident in shared memory for the life of a
// persistent CTA and fetches each input pixel from L2/HBM once per tile:
//
//   * B (weights): the host packs them once at load into the UMMA no-swizzle K-major core-matrix image
//     [N/8][K/16][8 rows][16 B]; one 1-D bulk copy brings it in at kernel start.
//   * A (activations): the haloed input tile is written ONCE per tile as 16-byte channel planes
//     [Cin/16][halo pixel][16 B] by cp.async (zero-filled outside the image = the convolution's zero padding).  In this
//     layout the 8 rows of a core matrix are 8 x-adjacent pixels and moving the window by one pixel moves the start
//     address by 16 bytes, so each of the 9 taps is just a different descriptor start address into the same tile: no
//     im2col copy and no re-fetch per tap (the tap-per-TMA-box kernel in conv_umma.cu is bound by the TMA request rate).
//   * images are stacked vertically into one tall canvas separated by `gut` zero rows, so 16-row tiles waste no rows on
//     small maps; a gutter row serves as the bottom padding of one image and the top padding of the next.
//   * pooled layers (PHASE): the tile is 16 x 32 pre-pool pixels held as four accumulators, one per position (dy,dx) of
//     the 2x2 window (the planes are split by x parity so that each accumulator's rows are contiguous again), so TMEM
//     lane i of the four accumulators holds the four members of pooled pixel i and the max-pool happens in registers.
//     Requantisation is monotone, so max-then-requantise == requantise-then-max (slim_yolo_v2.py:229-231).
//
// Warp roles (640 threads): warp 0 = TMEM allocator + MMA issuer (one lane), warps 1-3 = cp.async producers (one halo
// row per warp at a time), warps 4-19 = two epilogue groups of 8 warps (two warps per TMEM lane quarter, each half of the
// columns); group g drains accumulator buffer g, so the epilogues of two consecutive tiles run concurrently (a single
// group is latency-bound: ~0.2 instructions per cycle per warp, measured with clock64 stamps, tools/ws_timeline.py).
// Un-phased layers with four accumulator buffers run FOUR groups of four warps instead (one warp per quarter, all columns).
// The MMA warp stays converged and elects one lane per tile to issue (elect.sync): with `if (lane == 0)` the issue interval
// is ~80-120 cycles per MMA, with an elected lane ~42 (tools/micro/umma_issue.cu), which is what the thin layers need.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <climits>
#include <cstdio>

namespace yb {

#ifdef YB_WS_TIMELINE
#define WS_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 64 && lane == 0) p.dbg[it * 8 + (slot)] = clock64(); } while (0)
#else
#define WS_STAMP(slot) do { } while (0)
#endif

constexpr int WS_PROD_WARPS = 3;                        // 1 + 3 + 16 = 20 warps: registers are allocated per 4 warps, 20 x 96 x 32 fits
constexpr int WS_PROD_THREADS = WS_PROD_WARPS * 32;
constexpr int WS_EPI_WARPS = 8;                         // per group: two warps per TMEM lane quarter
constexpr int WS_EPI_THREADS = WS_EPI_WARPS * 32;
constexpr int WS_EPI_GROUPS = 2;                        // group g drains accumulator buffer g: two tiles' epilogues overlap
constexpr int WS_THREADS = 32 + WS_PROD_THREADS + WS_EPI_GROUPS * WS_EPI_THREADS;
constexpr int WS_MAX_STAGES = 4;
constexpr int WS_MAX_TBUF = 4;                          // TMEM accumulator buffers (2 when 4 of them exceed 512 columns)

struct WsParams {
    const int8_t *in;
    int n_img, H, W, cs_in;
    int period;                  // H + gut: canvas rows per image
    unsigned period_magic;       // ceil(2^32 / period)
    int canvas_rows;             // n_img * period
    int tiles_x, num_tiles;
    int step_x, step_y;          // gridDim.x % tiles_x, gridDim.x / tiles_x: tile coordinates advance without divisions
    int N;                       // GEMM N = cs_out
    int cs_out;
    int nplanes;                 // cs_in / 16 (power of two)
    int nplanes_log2;
    int kc;                      // 16-byte K chunks per output channel in the weight image
    uint32_t plane_stride;       // bytes between 16-byte channel planes of a stage (x2 parities when PHASE)
    uint32_t stage_bytes;
    int stages;
    uint32_t w_bytes;
    uint32_t off_stage, off_bias, off_bar;
    uint32_t tmem_cols, tmem_buf_stride;
    int tbufs, tbufs_log2;       // accumulator buffers in TMEM: 4 when they fit 512 columns, else 2
    int OH, OW;                  // output map (pooled when q.pool)
    LayerQ q;
    EpiConst k;
    const uint8_t *wimg;
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    long long *dbg;              // optional timeline (YB_WS_TIMELINE builds): [cta][tile][8] clock64 stamps
    // weight-streaming variant (BSTREAM): the layer's weights do not fit shared memory; one tap's weights for all output
    // channels (b_chunk_bytes = N * cs_in) travel through a ring of b_slots slots at the start of shared memory
    const uint8_t *wtap;         // [tap][128-channel plane][N/8][8][8][16 B]: one chunk = b_chunk_bytes = N * 128
    uint32_t b_chunk_bytes;
    int b_slots;
    int b_resident;              // all 9 * planes chunks fit in shared memory (pred: 18 x 6 KB): loaded once, no ring traffic per tile
    // flattened-raster tiles (BSTREAM, un-pooled, narrow maps): the canvas rows are W + 1 pixels wide (the extra one is out of
    // bounds for the TMA box = zero = the horizontal padding of both neighbours), an M tile is 128 consecutive pixels of
    // that stream, its halo is raster_rows whole rows (+ one zero pixel in front), every tap is a start offset
    int raster;                  // 0 / 1
    int rP;                      // W + 1
    unsigned rP_magic;           // ceil(2^32 / rP)
    int raster_rows;             // rows of a halo tile
};

// TMA-fed variant (TMAIN, un-phased tiles with 32 / 64 / 128 input channels): the halo tile is pixel-major
// [18 rows][WS_TMA_PITCH(cs_in) pixels][cs_in bytes], written by swizzled TMA boxes (one swizzle row = one pixel), and the 9
// taps are row-shifted descriptor starts into it (tools/micro/umma_swz_shift.cu: the swizzle is a function of the absolute
// shared-memory address, so shifted starts read the right bytes with descriptor base offset 0).  Boxes have a fixed height,
// so there is one tensor map per power of two and a run of rows (a tile can straddle images of the canvas) is cut into them.
struct WsMaps { CUtensorMap m[5]; };                                    // box heights 1, 2, 4, 8, 16
__host__ __device__ constexpr int ws_tma_pitch(int cs_in) { return cs_in == 32 ? 12 : 10; }   // rows of 384 / 640 / 1280 B: 128-byte multiples

// tile geometry
//   !PHASE: 8 x 16 pixels; halo 10 x 18, plane = [18][10] pixels, tap (kh,kw) -> +(kh*10+kw) pixels, row group stride 10
//    PHASE: 16 x 32 pre-pool pixels; halo 18 x 34, planes split by x parity: [par][34][9], accumulator (dy,dx) and tap
//           (kh,kw) -> parity (dx+kw)&1, +((dy+kh)*9 + ((dx+kw)>>1)) pixels, row group stride 2*9
template <bool PHASE> struct WsGeom {
    static constexpr int TW = PHASE ? 16 : 8, TH = PHASE ? 32 : 16;
    static constexpr int HW = TW + 2, HH = TH + 2;
    static constexpr int PITCH = PHASE ? 9 : 10;
    static constexpr int PIX = HW * HH;                       // halo pixels per 16-byte channel chunk
    static constexpr int NACC = PHASE ? 4 : 1;
    static constexpr uint32_t SBO = PHASE ? 2u * 9u * 16u : 10u * 16u;
};

template <bool PHASE, int EPI, bool ACT>
__device__ __forceinline__ void ws_epilogue_tile(const WsParams &p, uint32_t taddr, int cbeg, int cend, int lane, int q4,
                                                 int tx0, int ty0, const int *s_bias, uint32_t bar_tempty, unsigned &ovf)
{
    const int r = q4 * 32 + lane;
    const int g = r >> 3, xl = r & 7;
    if (PHASE) {
        // row r = pooled pixel (ty0/2 + g, tx0/2 + xl); its four members are lane r of the four accumulators
        const int cy = ty0 + 2 * g;                            // canvas row of the window's top row (even)
        const int n = (int)__umulhi((unsigned)cy, p.period_magic);
        const int y = cy - n * p.period;
        const int oy = y >> 1, ox = (tx0 >> 1) + xl;
        const bool valid = cy < p.canvas_rows && oy < p.OH && ox < p.OW;
        int8_t *dst = p.out + (((size_t)n * p.OH + oy) * p.OW + ox) * p.cs_out;
        const uint32_t accs = (uint32_t)p.N;                   // column stride between the four accumulators
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
            int v0[16], v1[16], v2[16], v3[16];
            tmem_ld16(taddr + c0, v0);
            tmem_ld16(taddr + accs + c0, v1);
            tmem_ld16(taddr + 2 * accs + c0, v2);
            tmem_ld16(taddr + 3 * accs + c0, v3);
            tmem_ld_wait();
            if (c0 + 16 >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }      // the last chunk is in registers: release the accumulators
#pragma unroll
            for (int j = 0; j < 16; ++j) v0[j] = max(max(v0[j], v1[j]), max(v2[j], v3[j]));
            const uint4 w = requant16<EPI, ACT>(v0, s_bias, c0, p, ovf, valid);
            if (valid) *reinterpret_cast<uint4 *>(dst + c0) = w;
        }
        if (cbeg >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }
    } else {
        int cy = ty0 + g, x = tx0 + xl;
        if (p.raster) {                                        // row r = stream pixel 128 * tile + r (ty0 = 16 * tile)
            const int q = 8 * ty0 + r;
            cy = (int)__umulhi((unsigned)q, p.rP_magic);
            x = q - cy * p.rP;
        }
        const int n = (int)__umulhi((unsigned)cy, p.period_magic);
        const int y = cy - n * p.period;
        const bool inside = cy < p.canvas_rows && y < p.H && x < p.W;
        if (!p.q.pool) {
            int8_t *dst = p.out + (((size_t)n * p.H + y) * p.W + x) * p.cs_out;
            // both 16-column chunks of a 32-column block are loaded at once: one wait, the accumulator buffer is released as soon
            // as the last block is in registers (the MMA warp does not wait for the requantisation), and the two requantisations
            // are independent instruction streams
            int va[16], vb[16];
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                const bool two = c0 + 16 < cend;
                tmem_ld16(taddr + c0, va);
                if (two) tmem_ld16(taddr + c0 + 16, vb);
                tmem_ld_wait();
                if (c0 + 32 >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }
                const uint4 w0 = requant16<EPI, ACT>(va, s_bias, c0, p, ovf, inside);
                if (inside) *reinterpret_cast<uint4 *>(dst + c0) = w0;
                if (two) {
                    const uint4 w1 = requant16<EPI, ACT>(vb, s_bias, c0 + 16, p, ovf, inside);
                    if (inside) *reinterpret_cast<uint4 *>(dst + c0 + 16) = w1;
                }
            }
            if (cbeg >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }
        } else {
            // pooled layer on the un-phased tile (weights + phased tile do not fit): the 2x2 window of a pixel is lanes
            // {l, l^1, l^8, l^9} of this warp.  Max the raw accumulators with two shuffles, then every lane requantises
            // and stores its own quarter (4 channels) of each 16-channel chunk.
            const int oy = y >> 1, ox = x >> 1;
            const bool valid = inside && oy < p.OH && ox < p.OW;       // all four lanes of a window agree
            const int role = (xl & 1) | ((g & 1) << 1);
            int8_t *dst = p.out + (((size_t)n * p.OH + oy) * p.OW + ox) * p.cs_out + 4 * role;
            for (int c0 = cbeg; c0 < cend; c0 += 16) {
                int v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (c0 + 16 >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = max(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                    v[j] = max(v[j], __shfl_xor_sync(0xffffffffu, v[j], 8));
                }
                int mine[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    mine[j] = role == 0 ? v[j] : role == 1 ? v[4 + j] : role == 2 ? v[8 + j] : v[12 + j];
                const unsigned w = requant4<EPI, ACT>(mine, s_bias, c0 + 4 * role, p, ovf, valid);
                if (valid) *reinterpret_cast<unsigned *>(dst + c0) = w;
            }
            if (cbeg >= cend) { tc_fence_before(); mbar_arrive(bar_tempty); }
        }
    }
}

// All MMAs of one tile, fully unrolled (KHALF = channel-plane pairs per tap = cs_in / 32; 0 = the 16-channel case).
// Descriptor words: lo = start>>4 | LBO>>4 << 16, hi = SBO>>4 | version | layout.  Everything except the stage base
// (sa16), the parity half-plane (half16) and the plane-pair step (cstep16) is a compile-time constant.
template <bool PHASE, int KHALF, bool TMAIN = false>
__device__ __forceinline__ void ws_issue_tile(uint32_t d0, uint32_t N, uint32_t sa16, uint32_t plane16, uint32_t half16, uint32_t cstep16,
                                              uint32_t b16, uint32_t bhi, uint32_t idesc)
{
    using G = WsGeom<PHASE>;
    static_assert(!TMAIN || (!PHASE && (KHALF == 1 || KHALF == 2 || KHALF == 4)), "TMA-fed tiles: un-phased, 32 / 64 / 128 input channels");
    constexpr uint32_t T_CIN = 32u * (KHALF ? KHALF : 1), T_PITCH = (uint32_t)ws_tma_pitch((int)T_CIN);
    constexpr uint32_t T_LAYOUT = T_CIN == 128 ? 2u : T_CIN == 64 ? 4u : 6u;          // swizzle 128B / 64B / 32B
    const uint32_t ahi = TMAIN ? (((T_PITCH * T_CIN) >> 4) | (1u << 14) | (T_LAYOUT << 29)) : ((G::SBO >> 4) | (1u << 14));
    const uint32_t blo0 = b16 | (8u << 16);                         // LBO = 128 B between the two K halves of the weights
#pragma unroll
    for (int acc = 0; acc < G::NACC; ++acc) {
        const int dy = acc >> 1, dx = acc & 1;
        const uint32_t d = d0 + (uint32_t)acc * N;
        if (KHALF == 0) {
            // 16 input channels: one MMA (K = 32) spans two taps; its two 16-byte K halves are the same plane at two tap
            // offsets (LBO = their distance).  Tap 9 has zero weights; its A half reads tap 8 + 16 B.
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int t0 = 2 * m, t1 = 2 * m + 1;
                const int kh0 = t0 / 3, kw0 = t0 % 3, kh1 = t1 / 3, kw1 = t1 % 3;
                uint32_t o0, o1;
                bool swapped = false;
                if (PHASE) {
                    const int p0 = (dx + kw0) & 1, p1 = (dx + kw1) & 1;
                    const uint32_t c0 = (uint32_t)((dy + kh0) * G::PITCH + ((dx + kw0) >> 1));
                    const uint32_t c1 = (uint32_t)((dy + kh1) * G::PITCH + ((dx + kw1) >> 1));
                    o0 = (p0 ? half16 : 0u) + c0;
                    o1 = m < 4 ? (p1 ? half16 : 0u) + c1 : o0 + 1u;
                    // descriptors hold unsigned strides: when the second tap sits at the lower address (it is in the even
                    // half-plane and the first in the odd one), start from it and use the copy of the weights whose K
                    // halves are swapped (chunks 10..19)
                    swapped = m < 4 && p0 == 1 && p1 == 0;
                } else {
                    o0 = (uint32_t)(kh0 * G::PITCH + kw0);
                    o1 = m < 4 ? (uint32_t)(kh1 * G::PITCH + kw1) : o0 + 1u;
                }
                const uint32_t lo = swapped ? o1 : o0, hi = swapped ? o0 : o1;
                const uint32_t alo = (sa16 + lo) | ((hi - lo) << 16);
                const uint32_t blo = blo0 + (uint32_t)((swapped ? 10 : 0) + 2 * m) * 8u;      // 128 B per K chunk
                if (m == 0) umma_i8_lohi<false>(d, alo, ahi, blo, bhi, idesc);
                else umma_i8_lohi<true>(d, alo, ahi, blo, bhi, idesc);
            }
        } else {
            const uint32_t a_lbo = TMAIN ? (1u << 16) : (plane16 << 16);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int kh = tap / 3, kw = tap % 3;
                uint32_t abase;
                if (TMAIN) abase = sa16 + (uint32_t)(kh * (int)T_PITCH + kw) * (T_CIN >> 4);   // one pixel = T_CIN bytes
                else if (PHASE) abase = sa16 + (((dx + kw) & 1) ? half16 : 0u) + (uint32_t)((dy + kh) * G::PITCH + ((dx + kw) >> 1));
                else abase = sa16 + (uint32_t)(kh * G::PITCH + kw);
#pragma unroll
                for (int c2 = 0; c2 < KHALF; ++c2) {
                    const uint32_t alo = (abase + (uint32_t)c2 * (TMAIN ? 2u : cstep16)) | a_lbo;   // TMAIN: 32 bytes on inside the pixel
                    const uint32_t blo = blo0 + (uint32_t)(tap * KHALF + c2) * 16u;            // 256 B of weights per K = 32
                    if (tap == 0 && c2 == 0) umma_i8_lohi<false>(d, alo, ahi, blo, bhi, idesc);
                    else umma_i8_lohi<true>(d, alo, ahi, blo, bhi, idesc);
                }
            }
        }
    }
}

// BSTREAM: the MMAs of ONE tap (KHALF K-steps of 32 channels) against the weight chunk at b16 (16-byte units).  The halo
// tile is TMA-written: pixels of 128 bytes (cs_in = 256: two such planes, plane16 apart), or of cs_in bytes below that.
// row_px = pixels between the tile rows a tap moves over (10, or W + 1 for raster tiles); sbo_px = pixels between 8-row groups
template <int KHALF>
__device__ __forceinline__ void ws_issue_tap(uint32_t d, uint32_t sa16, uint32_t plane16, int tap, int row_px, uint32_t sbo_px,
                                             uint32_t b16, bool first, uint32_t idesc)
{
    // one weight chunk = (tap, 128-channel plane): 4 K-steps; sa16 already points at the plane
    constexpr uint32_t CIN_PL = 128u, LAYOUT = 2u, KC_T = 8u;                      // 16-byte K chunks per output channel and chunk
    static_assert(KHALF == 4 || KHALF == 8, "128 / 256 input channels");
    const uint32_t ahi = ((sbo_px * CIN_PL) >> 4) | (1u << 14) | (LAYOUT << 29);
    const uint32_t bhi = (KC_T * 8u) | (1u << 14);
    const int kh = tap / 3, kw = tap - 3 * kh;
    const uint32_t abase = sa16 + (uint32_t)(kh * row_px + kw) * (CIN_PL >> 4);
    (void)plane16;
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
        const uint32_t alo = (abase + (uint32_t)c2 * 2u) | (1u << 16);
        const uint32_t blo = (b16 + (uint32_t)c2 * 16u) | (8u << 16);
        if (c2 == 0 && first) umma_i8_lohi<false>(d, alo, ahi, blo, bhi, idesc);
        else umma_i8_lohi<true>(d, alo, ahi, blo, bhi, idesc);
    }
}

// KHALF = cs_in / 32 (0 for 16 input channels) is a template parameter so that each kernel holds exactly one fully
// unrolled issue sequence with compile-time operand offsets (a run-time switch over all five kept their descriptor words
// live at once and spilled ~1.5 KB in the issuing lane: 72 cycles per MMA instead of the ~45 the hardware needs).
template <bool PHASE, int EPI, int KHALF, bool TMAIN = false, bool BSTREAM = false>
__global__ void __launch_bounds__(WS_THREADS, 1) conv3x3_ws_kernel(const WsParams p, const __grid_constant__ WsMaps maps)
{
    static_assert(!BSTREAM || (TMAIN && !PHASE && (KHALF == 4 || KHALF == 8)), "weight streaming: TMA-fed un-phased tiles, 128 / 256 input channels");
    using G = WsGeom<PHASE>;
    pdl_launch_dependents();
#ifdef YB_WS_TIMELINE
    if (p.dbg && threadIdx.x == 0 && blockIdx.x < 256) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[512 + blockIdx.x] = t; }
#endif
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = TMAIN ? (smem_u32(smem_raw) + 1023u) & ~1023u : (smem_u32(smem_raw) + 127u) & ~127u;   // swizzle atoms: 1024 B
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t wsm = base;                                           // resident weight image (BSTREAM: the weight ring)
    const uint32_t stage0 = base + p.off_stage;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    // barrier layout: full[MAX], empty[MAX], tmem_full[TBUF], tmem_empty[TBUF], weights, then the TMEM address slot
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (WS_MAX_STAGES + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (2 * WS_MAX_STAGES + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (2 * WS_MAX_STAGES + WS_MAX_TBUF + b); };
    const uint32_t bar_w = bar0 + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_TBUF);
    const uint32_t tmem_slot = bar0 + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_TBUF + 1);
    auto bar_bfull = [&](int i) { return bar0 + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_TBUF + 2 + i); };       // BSTREAM: weight ring
    auto bar_bempty = [&](int i) { return bar0 + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_TBUF + 6 + i); };
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_TBUF + 1));

    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_MAX_STAGES; ++s) { mbar_init(bar_full(s), TMAIN ? 1 : WS_PROD_THREADS); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < WS_MAX_TBUF; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), (!PHASE && p.tbufs == 4) ? WS_EPI_THREADS / 2 : WS_EPI_THREADS); }
        mbar_init(bar_w, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bempty(i), 1); }
        fence_barrier_init();
        if (!BSTREAM) {
            // the whole layer's weights, once
            mbar_expect_tx(bar_w, p.w_bytes);
            for (uint32_t o = 0; o < p.w_bytes; o += 32768u) {
                const uint32_t n = p.w_bytes - o < 32768u ? p.w_bytes - o : 32768u;
                bulk_load_1d(wsm + o, p.wimg + o, n, bar_w);
            }
        }
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    if (BSTREAM && p.raster) {
        // the zero pixel in front of every plane of every stage (never written afterwards)
        const int npl = KHALF == 8 ? 2 : 1;
        for (int i = threadIdx.x; i < p.stages * npl * 32; i += blockDim.x) {
            const int st_ = i / (npl * 32), rem = i - st_ * npl * 32, pl = rem / 32, wd = rem & 31;
            reinterpret_cast<uint32_t *>(base_ptr + p.off_stage + (uint32_t)st_ * p.stage_bytes + (uint32_t)pl * p.plane_stride)[wd] = 0u;
        }
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;                 // |b| < 2^21: exact
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();

    if (warp == 0) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop (converged, so the descriptor arithmetic stays in uniform registers); one elected
        // lane issues.  Descriptor start addresses are in 16-byte units (low 14 bits of the descriptor), so moving to
        // another tap / channel-plane pair / weight chunk is an integer add on the 64-bit descriptor.
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t bhi = ((uint32_t)p.kc * 8u) | (1u << 14);   // SBO = kc * 128 B between 8-channel groups of the weights
        const uint32_t b16 = wsm >> 4;
        const uint32_t plane16 = p.plane_stride >> 4;             // LBO of A: the next 16-byte channel plane
        const uint32_t half16 = p.plane_stride >> 5;              // PHASE: x-parity half-plane, in 16-byte units
        const uint32_t cstep16 = p.plane_stride >> 3;             // two channel planes = one K = 32 step
        if (!BSTREAM || p.b_resident) mbar_wait(bar_w, 0);
        int it = 0, s = 0, bslot = 0;
        uint32_t ph = 0, rph = 0;                                 // stage phase, weight-ring phase
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & (p.tbufs - 1);
            const uint32_t bph = (uint32_t)(it >> p.tbufs_log2) & 1u;
            mbar_wait(bar_tempty(buf), bph ^ 1u);                 // epilogue has drained this accumulator buffer
            WS_STAMP(0);
            mbar_wait(bar_full(s), ph);                           // the haloed tile is in shared memory
            WS_STAMP(1);
            tc_fence_after();
            const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
            if constexpr (BSTREAM) {
                // tap by tap: each needs its weight chunk in the ring; the chunk's slot is released by a commit
                const uint32_t d0 = tmem_base + (uint32_t)buf * p.tmem_buf_stride;
                uint32_t a16 = sa >> 4;
                int row_px = ws_tma_pitch(128);
                uint32_t sbo_px = (uint32_t)row_px;
                if (p.raster) {
                    // stream pixel 128 * tile sits toff pixels into the second halo row; tap (0,0) starts one row and one
                    // pixel earlier, which the zero pixel in front of the plane makes a non-negative offset
                    const int q0 = 128 * tile;
                    const int toff = q0 - (int)__umulhi((unsigned)q0, p.rP_magic) * p.rP;
                    a16 += (uint32_t)toff * 8u;
                    row_px = p.rP; sbo_px = 8u;
                }
                constexpr int NPL = KHALF / 4;                     // 128-channel planes = weight chunks per tap
                if (p.b_resident) {
                    // every chunk is in shared memory: one fully unrolled issue sequence with compile-time tap offsets (the
                    // chunk loop below costs ~100 cycles per MMA at N = 48, twice what the tensor pipe needs)
                    if (elect_one()) {
                        const uint32_t w16 = wsm >> 4, ch16 = p.b_chunk_bytes >> 4;
#pragma unroll
                        for (int ck = 0; ck < 9 * NPL; ++ck)
                            ws_issue_tap<KHALF>(d0, a16 + (uint32_t)(ck % NPL) * plane16, plane16, ck / NPL, row_px, sbo_px,
                                                w16 + (uint32_t)ck * ch16, ck == 0, idesc);
                        umma_commit(bar_empty(s));
                        umma_commit(bar_tfull(buf));
                    }
                    __syncwarp();
                } else
                for (int ck = 0; ck < 9 * NPL; ++ck) {
                    const int tap = ck / NPL, pl = ck - tap * NPL;
                    if (!p.b_resident) { mbar_wait(bar_bfull(bslot), rph); tc_fence_after(); }
                    if (elect_one()) {
                        ws_issue_tap<KHALF>(d0, a16 + (uint32_t)pl * plane16, plane16, tap, row_px, sbo_px,
                                            (wsm + (uint32_t)(p.b_resident ? ck : bslot) * p.b_chunk_bytes) >> 4, ck == 0, idesc);
                        if (!p.b_resident) umma_commit(bar_bempty(bslot));
                        if (ck == 9 * NPL - 1) { umma_commit(bar_empty(s)); umma_commit(bar_tfull(buf)); }
                    }
                    __syncwarp();
                    if (!p.b_resident && ++bslot == p.b_slots) { bslot = 0; rph ^= 1u; }
                }
            } else {
                if (elect_one()) {
                    const uint32_t d0 = tmem_base + (uint32_t)buf * p.tmem_buf_stride;
                    const uint32_t sa16 = sa >> 4;
                    ws_issue_tile<PHASE, KHALF, TMAIN>(d0, (uint32_t)p.N, sa16, plane16, half16, cstep16, b16, bhi, idesc);
                    umma_commit(bar_empty(s));                         // stage free once these MMAs have read it
                    umma_commit(bar_tfull(buf));                       // accumulators complete
                }
            }
            __syncwarp();
            WS_STAMP(2);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (TMAIN && warp <= WS_PROD_WARPS) {
        // ===================== TMA producer: one thread =====================
        if (BSTREAM && warp == 2 && lane == 0 && p.b_resident) {
            // every weight chunk once: they all fit
            const uint32_t total = (uint32_t)(9 * (KHALF / 4)) * p.b_chunk_bytes;
            mbar_expect_tx(bar_w, total);
            for (uint32_t o = 0; o < total; o += 32768u) {
                const uint32_t nb = total - o < 32768u ? total - o : 32768u;
                bulk_load_1d(wsm + o, p.wtap + o, nb, bar_w);
            }
        } else if (BSTREAM && warp == 2 && lane == 0) {
            // weight chunks: tile after tile, tap after tap
            int slot = 0;
            uint32_t bph = 0;
            const int nchunks = 9 * (KHALF / 4);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
                for (int ck = 0; ck < nchunks; ++ck) {
                    mbar_wait(bar_bempty(slot), bph ^ 1u);
                    mbar_expect_tx(bar_bfull(slot), p.b_chunk_bytes);
                    const uint32_t dst = wsm + (uint32_t)slot * p.b_chunk_bytes;
                    const uint8_t *src = p.wtap + (size_t)ck * p.b_chunk_bytes;
                    for (uint32_t o = 0; o < p.b_chunk_bytes; o += 32768u) {
                        const uint32_t nb = p.b_chunk_bytes - o < 32768u ? p.b_chunk_bytes - o : 32768u;
                        bulk_load_1d(dst + o, src + o, nb, bar_bfull(slot));
                    }
                    if (++slot == p.b_slots) { slot = 0; bph ^= 1u; }
                }
        }
        if (warp == 1 && lane == 0) {
            constexpr int NPL = KHALF == 8 ? 2 : 1;                                // 128-byte channel planes of the tile
            constexpr int CIN = KHALF >= 4 ? 128 : 32 * (KHALF ? KHALF : 1), PITCHPX = ws_tma_pitch(CIN);
            constexpr uint32_t ROW_BYTES = (uint32_t)(PITCHPX * CIN);
            int it = 0, s = 0;
            uint32_t ph = 0;
            int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int tx0 = tx * G::TW, ty0 = ty * G::TH;
                mbar_wait(bar_empty(s), ph ^ 1u);
                WS_STAMP(3);
                const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
                if (BSTREAM && p.raster) {
                    // raster tile: raster_rows whole canvas rows from the row above stream pixel 128 * tile
                    const uint32_t row_bytes = (uint32_t)p.rP * 128u;
                    mbar_expect_tx(bar_full(s), (uint32_t)(NPL * p.raster_rows) * row_bytes);
                    int r = 0, cy = (int)__umulhi((unsigned)(128 * tile), p.rP_magic) - 1;
                    while (r < p.raster_rows) {
                        const int n = cy < 0 ? 0 : (int)__umulhi((unsigned)cy, p.period_magic);
                        int y = cy - n * p.period;
                        int run = min(p.raster_rows - r, p.period - y);
                        while (run > 0) {
                            const int lg = run >= 16 ? 4 : run >= 8 ? 3 : run >= 4 ? 2 : run >= 2 ? 1 : 0, h = 1 << lg;
#pragma unroll
                            for (int pl = 0; pl < NPL; ++pl)
                                tma_load_4d(sa + (uint32_t)pl * p.plane_stride + 128u + (uint32_t)r * row_bytes, &maps.m[lg], bar_full(s), 128 * pl, 0, y, n);
                            r += h; y += h; cy += h; run -= h;
                        }
                    }
                    WS_STAMP(4);
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                    tx += p.step_x; ty += p.step_y;
                    if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
                    continue;
                }
                mbar_expect_tx(bar_full(s), (uint32_t)(NPL * G::HH) * ROW_BYTES);
                // the 18 halo rows = canvas rows ty0-1 .. ty0+16: runs of rows of one image (its gutter rows and the rows
                // above / below the canvas are out of bounds for the box = zero filled = the convolution's padding)
                int r = 0, cy = ty0 - 1;
                while (r < G::HH) {
                    const int n = cy < 0 ? 0 : (int)__umulhi((unsigned)cy, p.period_magic);
                    int y = cy - n * p.period;
                    int run = min(G::HH - r, p.period - y);
                    while (run > 0) {
                        const int lg = run >= 16 ? 4 : run >= 8 ? 3 : run >= 4 ? 2 : run >= 2 ? 1 : 0, h = 1 << lg;
#pragma unroll
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_load_4d(sa + (uint32_t)(pl * G::HH + r) * ROW_BYTES, &maps.m[lg], bar_full(s), 128 * pl, tx0 - 1, y, n);
                        r += h; y += h; cy += h; run -= h;
                    }
                }
                WS_STAMP(4);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
                tx += p.step_x; ty += p.step_y;
                if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
            }
        }
    } else if (warp <= WS_PROD_WARPS) {
        // ===================== cp.async producers =====================
        // One halo row per warp at a time.  Along a row the 16-byte pieces are contiguous in global memory in the order
        // (pixel, channel chunk), so piece qr of the row lives at row_src + 16*qr; only the shared-memory side scatters
        // (channel chunk -> plane, and pixel parity -> half-plane when PHASE).
        const int pw = warp - 1;
        const int row_pieces = G::HW << p.nplanes_log2;
        const int lag = p.stages >= 3 ? p.stages - 2 : 1;          // tiles in flight per thread before their arrival is signalled
        // Everything about a lane's pieces except the row is tile-invariant: piece qr = lane + 32 t of a row is halo pixel
        // hx = qr / nplanes, channel chunk c = qr % nplanes.
        constexpr int MAXIT = 5;                                    // ceil(10 * 16 / 32): 256 input channels, un-phased tile
        int hxv[MAXIT];
        uint32_t dof[MAXIT];                                        // shared-memory offset inside the stage, without the row part
#pragma unroll
        for (int t = 0; t < MAXIT; ++t) {
            const int qr = lane + 32 * t;
            const int hx = qr >> p.nplanes_log2, c = qr & (p.nplanes - 1);
            hxv[t] = qr < row_pieces ? hx : 0x40000000;            // out of range: fails the x test below
            dof[t] = (uint32_t)c * p.plane_stride + (PHASE ? (uint32_t)(hx & 1) * (p.plane_stride >> 1) + (uint32_t)(hx >> 1) * 16u : (uint32_t)hx * 16u);
        }
        const int gut = p.period - p.H;
        const int cs_shift = 4 + p.nplanes_log2;                    // log2(cs_in)
        // thin rows: per-lane piece slots for the whole tile (tile-invariant), piece q = pt + 96 t
        constexpr int LIN_SLOTS = (G::HH * 32 + WS_PROD_THREADS - 1) / WS_PROD_THREADS;
        const bool linear = row_pieces <= 32 && G::HH * row_pieces <= LIN_SLOTS * WS_PROD_THREADS;
        int lin_hy[LIN_SLOTS], lin_hx[LIN_SLOTS], lin_c16[LIN_SLOTS];
        uint32_t lin_dst[LIN_SLOTS];
#pragma unroll
        for (int t = 0; t < LIN_SLOTS; ++t) {
            const int q = (threadIdx.x - 32) + WS_PROD_THREADS * t;
            const int hy = q / row_pieces, qr = q - hy * row_pieces;
            const int hx = qr >> p.nplanes_log2, c = qr & (p.nplanes - 1);
            lin_hy[t] = (linear && hy < G::HH) ? hy : -1;
            lin_hx[t] = hx; lin_c16[t] = 16 * c;
            lin_dst[t] = (uint32_t)(hy * G::PITCH) * 16u + (uint32_t)c * p.plane_stride +
                         (PHASE ? (uint32_t)(hx & 1) * (p.plane_stride >> 1) + (uint32_t)(hx >> 1) * 16u : (uint32_t)hx * 16u);
        }
        int it = 0, s = 0, s_arrive = 0;
        uint32_t ph = 0;
        int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int tx0 = tx * G::TW, ty0 = ty * G::TH;
            mbar_wait(bar_empty(s), ph ^ 1u);
            if (pw == 0) WS_STAMP(3);
            const uint32_t sa = stage0 + (uint32_t)s * p.stage_bytes;
            if (linear) {
                // thin rows (<= 32 pieces): pieces are dealt to the producer lanes across the whole tile
#pragma unroll
                for (int t = 0; t < LIN_SLOTS; ++t) {
                    if (lin_hy[t] >= 0) {
                        const int cy = ty0 - 1 + lin_hy[t];
                        const int n = (int)__umulhi((unsigned)max(cy, 0), p.period_magic);
                        const int y = cy - n * p.period;
                        const bool ok = cy >= 0 && cy < p.canvas_rows && y < p.H && (unsigned)(tx0 - 1 + lin_hx[t]) < (unsigned)p.W;
                        const int off = (((cy - n * gut) * p.W + (tx0 - 1 + lin_hx[t])) << cs_shift) + lin_c16[t];
                        cp_async16(sa + lin_dst[t], p.in + (ok ? off : 0), ok ? 16u : 0u);
                    }
                }
            } else {
#pragma unroll 2
            for (int hy = pw; hy < G::HH; hy += WS_PROD_WARPS) {
                const int cy = ty0 - 1 + hy;
                const int n = (int)__umulhi((unsigned)max(cy, 0), p.period_magic);
                const int y = cy - n * p.period;
                const bool rowok = cy >= 0 && cy < p.canvas_rows && y < p.H;
                // byte offset of halo pixel hx = 0 of this row (32-bit: the host checks the map is < 2 GB); image row index
                // = canvas row minus the gutter rows before it
                const int row_off = ((cy - n * gut) * p.W + (tx0 - 1)) << cs_shift;
                const uint32_t row_dst = sa + (uint32_t)(hy * G::PITCH) * 16u;
#pragma unroll
                for (int t = 0; t < MAXIT; ++t) {
                    if (t * 32 < row_pieces) {
                        const bool ok = rowok && (unsigned)(tx0 - 1 + hxv[t]) < (unsigned)p.W;
                        const int8_t *src = p.in + (ok ? row_off + 16 * (lane + 32 * t) : 0);
                        if (hxv[t] < 0x40000000) cp_async16(row_dst + dof[t], src, ok ? 16u : 0u);
                    }
                }
            }
            }
            cp_async_commit();
            if (pw == 0) WS_STAMP(4);
            if (it >= lag) {
                if (lag == 1) cp_async_wait<1>(); else cp_async_wait<2>();
                fence_proxy_async();
                mbar_arrive(bar_full(s_arrive));
                if (++s_arrive == p.stages) s_arrive = 0;
            }
            if (pw == 0) WS_STAMP(5);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            tx += p.step_x; ty += p.step_y;
            if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        }
        // drain: the last `lag` tiles
        cp_async_wait<0>();
        fence_proxy_async();
        for (int j = max(it - lag, 0); j < it; ++j) {
            mbar_arrive(bar_full(s_arrive));
            if (++s_arrive == p.stages) s_arrive = 0;
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew_all = warp - (1 + WS_PROD_WARPS);
        // Epilogue groups.  Two groups of eight warps (two warps per TMEM lane quarter, half of the columns each) when only two
        // accumulator buffers fit; FOUR groups of four warps (one warp per quarter, all columns) over four buffers otherwise: the
        // epilogue of a thin layer is paced by the dependency latency of ~400 instructions per warp and tile (two groups were
        // 85-90 % busy at ~2300 cycles per tile on conv3_1), so four tiles in flight with the per-tile address arithmetic
        // amortised over twice the columns beat two.  Tile it -> group it % ngrp, buffer it % tbufs: a buffer always has the
        // same group (parity waits on `tfull` are only safe for a waiter that drained the previous use itself).
        const int ngrp = (!PHASE && p.tbufs == 4) ? 4 : 2;
        const int wpg = (WS_EPI_GROUPS * WS_EPI_WARPS) / ngrp;
        const int grp = ew_all / wpg;
        const int ew = ew_all % wpg;
        const int q4 = warp & 3;                                   // TMEM lane quarter this warp may access
        const int cmid = ((p.N / 16 + 1) / 2) * 16;                // column split between the two warps of a quarter
        const int cbeg = (wpg == 4 || ew < 4) ? 0 : cmid, cend = (wpg == 4 || ew >= 4) ? p.N : cmid;
        unsigned ovf = 0;
        // tiles blockIdx.x + (grp + ngrp k) * gridDim.x: advance the tile coordinates ngrp grid strides at a time
        int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
        for (int g = 0; g < grp; ++g) { tx += p.step_x; ty += p.step_y; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; } }
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += ngrp * gridDim.x, it += ngrp) {
            const int buf = it & (p.tbufs - 1);
            const uint32_t bph = (uint32_t)(it >> p.tbufs_log2) & 1u;
            const int tx0 = tx * G::TW, ty0 = ty * G::TH;
            for (int r = 0; r < ngrp; ++r) { tx += p.step_x; ty += p.step_y; if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; } }
            mbar_wait(bar_tfull(buf), bph);
            if (ew_all == 0) WS_STAMP(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * p.tmem_buf_stride + ((uint32_t)(q4 * 32) << 16);
            if (p.q.activ) ws_epilogue_tile<PHASE, EPI, true>(p, taddr, cbeg, cend, lane, q4, tx0, ty0, s_bias, bar_tempty(buf), ovf);
            else ws_epilogue_tile<PHASE, EPI, false>(p, taddr, cbeg, cend, lane, q4, tx0, ty0, s_bias, bar_tempty(buf), ovf);
            if (ew_all == 0) WS_STAMP(7);
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
#ifdef YB_WS_TIMELINE
    if (p.dbg && threadIdx.x == 0 && blockIdx.x < 256) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbg[768 + blockIdx.x] = t; }
#endif
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Shared-memory plan; returns false when the layer does not fit this kernel.
static bool ws_plan(const ConvArgs &a, bool phase, WsParams *p, bool tma = false)
{
    if (tma && (phase || (a.cs_in != 32 && a.cs_in != 64 && a.cs_in != 128))) return false;
    const int nplanes = a.cs_in / 16;
    const int pix = phase ? WsGeom<true>::PIX : WsGeom<false>::PIX;
    const int nacc = phase ? 4 : 1;
    // pad the plane stride so that the 8 lanes of a cp.async wavefront hit 8 different 16-byte bank groups
    int want = nplanes >= 8 ? 1 : 8 / nplanes;                           // plane stride in 16-B units, mod 8
    uint32_t plane_px;
    if (phase) {
        // two parity half-planes per channel chunk; each half is [34][9] pixels (+2 pixels of slack for the zero tap);
        // the stride must stay even so that a half-plane starts on a 16-byte boundary
        if (want & 1) want = 2;
        plane_px = 2u * (uint32_t)(WsGeom<true>::HH * WsGeom<true>::PITCH + 2);
        while ((int)(plane_px % 8) != want % 8) plane_px += 2;
    } else {
        plane_px = (uint32_t)pix + 2;
        while ((int)(plane_px % 8) != want % 8) ++plane_px;
    }
    p->plane_stride = plane_px * 16u;
    p->nplanes = nplanes; p->nplanes_log2 = ilog2(nplanes);
    p->kc = nplanes == 1 ? 20 : 9 * nplanes;                             // must match the image packed in yolo_b200_load
    p->N = a.cs_out; p->cs_out = a.cs_out;
    p->w_bytes = (uint32_t)p->N * (uint32_t)p->kc * 16u;
    p->stage_bytes = ((uint32_t)nplanes * p->plane_stride + 127u) & ~127u;
    if (tma) p->stage_bytes = ((uint32_t)(WsGeom<false>::HH * ws_tma_pitch(a.cs_in) * a.cs_in) + 1023u) & ~1023u;
    uint32_t nb = 32; while (nb < (uint32_t)(nacc * p->N)) nb <<= 1;
    if (2 * nb > 512) return false;
    p->tbufs = 4 * nb <= 512 ? 4 : 2; p->tbufs_log2 = p->tbufs == 4 ? 2 : 1;
    p->tmem_buf_stride = nb; p->tmem_cols = (uint32_t)p->tbufs * nb;
    const uint32_t walign = tma ? 1023u : 127u;                           // swizzled stages start on 1024-byte boundaries
    const uint32_t fixed = ((p->w_bytes + walign) & ~walign) + (uint32_t)p->N * 4u + 256u + walign + 1u /*alignment slack*/;
    const uint32_t budget = 227u * 1024u;
    if (fixed + 2 * p->stage_bytes > budget) return false;
    int stages = (int)((budget - fixed) / p->stage_bytes);
    if (stages > WS_MAX_STAGES) stages = WS_MAX_STAGES;
    p->stages = stages;
    p->off_stage = (p->w_bytes + walign) & ~walign;
    p->off_bias = p->off_stage + (uint32_t)stages * p->stage_bytes;
    p->off_bar = (p->off_bias + (uint32_t)p->N * 4u + 15u) & ~15u;
    return true;
}

// Shared-memory plan of the weight-streaming variant (weights too large to stay resident): ring of per-tap weight chunks
// at offset 0, then the TMA-written halo stages (one or two 128-byte channel planes).
static bool ws_plan_stream(const ConvArgs &a, WsParams *p)
{
    if (!a.wimg_tap || (a.cs_in != 128 && a.cs_in != 256)) return false;
    const int npl = a.cs_in / 128;
    p->nplanes = a.cs_in / 16; p->nplanes_log2 = ilog2(p->nplanes);
    p->kc = 9 * p->nplanes;
    p->N = a.cs_out; p->cs_out = a.cs_out;
    p->w_bytes = 0;
    p->plane_stride = (uint32_t)(WsGeom<false>::HH * ws_tma_pitch(128) * 128);       // one 128-byte channel plane of a stage
    // narrow un-pooled maps: flattened-raster tiles (see WsParams) waste W+1 : W instead of up to 8 : 1 columns
    static const bool raster_on = [] { const char *e = getenv("YOLO_B200_WS_RASTER"); return e ? atoi(e) != 0 : true; }();
    if (raster_on && !a.q.pool && a.W % 8 != 0 && a.W + 1 <= 64) {
        p->raster = 1; p->rP = a.W + 1;
        p->rP_magic = (unsigned)(((1ull << 32) + (unsigned)p->rP - 1) / (unsigned)p->rP);
        p->raster_rows = (3 * p->rP + 127) / p->rP + 1;
        p->plane_stride = 128u + (uint32_t)(p->raster_rows * p->rP) * 128u;
    }
    p->stage_bytes = ((uint32_t)npl * p->plane_stride + 1023u) & ~1023u;
    uint32_t nb = 32; while (nb < (uint32_t)p->N) nb <<= 1;
    if (2 * nb > 512) return false;
    p->tbufs = 4 * nb <= 512 ? 4 : 2; p->tbufs_log2 = p->tbufs == 4 ? 2 : 1;
    p->tmem_buf_stride = nb; p->tmem_cols = (uint32_t)p->tbufs * nb;
    p->b_chunk_bytes = (uint32_t)p->N * 128u;                                        // one (tap, 128-channel plane) of the weights
    const uint32_t budget = 227u * 1024u, tail = (uint32_t)p->N * 4u + 256u + 1024u + 16u;
    if (p->raster && 2 * p->b_chunk_bytes + 2 * p->stage_bytes + tail > budget) {     // raster tile too large: regular tile
        p->raster = 0;
        p->plane_stride = (uint32_t)(WsGeom<false>::HH * ws_tma_pitch(128) * 128);
        p->stage_bytes = ((uint32_t)npl * p->plane_stride + 1023u) & ~1023u;
    }
    if (2 * p->b_chunk_bytes + 2 * p->stage_bytes + tail > budget) return false;
    p->stages = 2;
    int slots = (int)((budget - tail - 2 * p->stage_bytes) / p->b_chunk_bytes);
    p->b_slots = slots > 4 ? 4 : slots;
    static const bool resident_on = [] { const char *e = getenv("YOLO_B200_WS_RESIDENT"); return e ? atoi(e) != 0 : true; }();
    if (resident_on && slots >= 9 * npl) { p->b_resident = 1; p->b_slots = 9 * npl; }   // the whole layer fits beside two stages (pred)
    int stages = (int)((budget - tail - (uint32_t)p->b_slots * p->b_chunk_bytes) / p->stage_bytes);
    p->stages = stages > WS_MAX_STAGES ? WS_MAX_STAGES : stages;
    p->off_stage = ((uint32_t)p->b_slots * p->b_chunk_bytes + 1023u) & ~1023u;
    p->off_bias = p->off_stage + (uint32_t)p->stages * p->stage_bytes;
    p->off_bar = (p->off_bias + (uint32_t)p->N * 4u + 15u) & ~15u;
    return true;
}

static bool ws_shape_ok(const ConvArgs &a)
{
    if (!a.wimg) return false;
    const int np = a.cs_in / 16;
    if (a.cs_in < 16 || a.cs_in % 16 || (np & (np - 1)) || np > 16) return false;
    if (a.cs_out < 16 || a.cs_out > 256 || a.cs_out % 16) return false;
    if (a.q.pool && (a.H < 2 || a.W < 2)) return false;
    if ((long long)a.n * (a.H + 2) * (a.H + 2) >= (1ll << 32)) return false;   // exactness range of the multiply-high division by H + gut
    if ((long long)a.n * a.H * a.W * a.cs_in >= (1ll << 31)) return false;     // the producers use 32-bit byte offsets
    return true;
}

static bool ws_stream_enabled()
{
    static const bool on = [] { const char *e = getenv("YOLO_B200_WS_STREAM"); return e ? atoi(e) != 0 : true; }();
    return on;
}

bool conv3x3_ws_supported(const ConvArgs &a)
{
    if (!ws_shape_ok(a)) return false;
    WsParams p;
    return ws_plan(a, false, &p) || (a.q.pool && ws_plan(a, true, &p)) || (ws_stream_enabled() && ws_plan_stream(a, &p));
}

// ---- tensor maps of the TMA-fed variant ----
typedef CUresult (*WsEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cudaError_t ws_make_maps(const ConvArgs &a, WsMaps *maps, int raster_px = 0)
{
    static WsEncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        enc = (WsEncodeTiledFn)fn;
    }
    const CUtensorMapSwizzle sw = a.cs_in >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : a.cs_in == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[4] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n };
    cuuint64_t strides[3] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.cs_in * a.W, (cuuint64_t)a.cs_in * a.W * a.H };
    cuuint32_t es[4] = { 1, 1, 1, 1 };
    for (int i = 0; i < 5; ++i) {
        const int cpl = a.cs_in > 128 ? 128 : a.cs_in;                          // bytes of a pixel in one plane of the tile
        cuuint32_t box[4] = { (cuuint32_t)cpl, (cuuint32_t)(raster_px ? raster_px : ws_tma_pitch(cpl)), (cuuint32_t)(1 << i), 1 };
        CUresult r = enc(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

template <bool PHASE, int EPI, int KHALF, bool TMAIN = false, bool BSTREAM = false>
static cudaError_t launch_ws_k(const ConvArgs &a, WsParams &p, cudaStream_t st, int sm_count)
{
    WsMaps maps;
    memset(&maps, 0, sizeof maps);
    if (TMAIN) { cudaError_t e = ws_make_maps(a, &maps, p.raster ? p.rP : 0); if (e != cudaSuccess) return e; }
    using G = WsGeom<PHASE>;
    const int gut = a.q.pool ? ((a.H & 1) ? 1 : 2) : 1;                  // pooled: image origins stay on even canvas rows
    p.in = a.in; p.n_img = a.n; p.H = a.H; p.W = a.W; p.cs_in = a.cs_in;
    p.period = a.H + gut;
    p.period_magic = (unsigned)(((1ull << 32) + (unsigned)p.period - 1) / (unsigned)p.period);
    p.canvas_rows = a.n * p.period;
    p.tiles_x = (a.W + G::TW - 1) / G::TW;
    p.num_tiles = p.tiles_x * ((p.canvas_rows + G::TH - 1) / G::TH);
    if (p.raster) { p.tiles_x = 1; p.num_tiles = (int)(((long long)p.canvas_rows * p.rP + 127) / 128); }
    p.OH = a.q.pool ? a.H / 2 : a.H; p.OW = a.q.pool ? a.W / 2 : a.W;
    p.q = a.q; p.wimg = a.wimg; p.wtap = a.wimg_tap; p.bias_sh = a.bias_sh; p.out = a.out; p.ovf = a.ovf;
#ifdef YB_WS_TIMELINE
    {   // debug builds only: stamps of CTA 0's first 64 tiles, printed after the launch (synchronises)
        static long long *dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, (64 * 8 + 512) * sizeof(long long));
        cudaMemsetAsync(dbg, 0, (64 * 8 + 512) * sizeof(long long), st);
        p.dbg = dbg;
    }
#endif
    const uint32_t smem_bytes = p.off_bar + 256u + (TMAIN ? 1024u : 128u);
    static bool attr_set[64] = {};     // per instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_ws_kernel<PHASE, EPI, KHALF, TMAIN, BSTREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    p.step_x = grid % p.tiles_x; p.step_y = grid / p.tiles_x;
    { cudaError_t le = launch_pdl(conv3x3_ws_kernel<PHASE, EPI, KHALF, TMAIN, BSTREAM>, dim3(grid), dim3(WS_THREADS), smem_bytes, st, p, maps); if (le != cudaSuccess) return le; }
#ifdef YB_WS_TIMELINE
    {
        long long h[64 * 8 + 512];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.dbg, sizeof h, cudaMemcpyDeviceToHost);
        {
            long long s0 = h[512], s1 = h[512], e0 = h[768], e1 = h[768];
            for (int i = 0; i < grid && i < 256; ++i) { s0 = s0 < h[512 + i] ? s0 : h[512 + i]; s1 = s1 > h[512 + i] ? s1 : h[512 + i]; e0 = e0 < h[768 + i] ? e0 : h[768 + i]; e1 = e1 > h[768 + i] ? e1 : h[768 + i]; }
            printf("WS cta times (globaltimer ns): first start 0, last start %lld, first end %lld, last end %lld; cta0 %lld .. %lld\n", s1 - s0, e0 - s0, e1 - s0, h[512] - s0, h[768] - s0);
        }
        const long long t0 = h[3] ? h[3] : h[0];
        printf("WS timeline PHASE=%d N=%d nplanes=%d stages=%d tiles=%d (cycles since first stamp): tile | mma: tempty_ok full_ok issued | prod: empty_ok loads_issued arrived | epi: tfull_ok done\n",
               (int)PHASE, p.N, p.nplanes, p.stages, p.num_tiles);
        for (int i = 0; i < 24; ++i)
            printf("  %2d | %7lld %7lld %7lld | %7lld %7lld %7lld | %7lld %7lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0, h[i*8+5]-t0, h[i*8+6]-t0, h[i*8+7]-t0);
    }
#endif
    return cudaGetLastError();
}

template <bool PHASE, int EPI>
static cudaError_t launch_ws(const ConvArgs &a, WsParams &p, cudaStream_t st, int sm_count, bool tma)
{
    if (!PHASE && p.b_slots)
        switch (p.nplanes >> 1) {
        case 4: return launch_ws_k<false, EPI, 4, true, true>(a, p, st, sm_count);
        case 8: return launch_ws_k<false, EPI, 8, true, true>(a, p, st, sm_count);
        default: return cudaErrorInvalidConfiguration;
        }
    if (!PHASE && tma)
        switch (p.nplanes >> 1) {
        case 1: return launch_ws_k<false, EPI, 1, true>(a, p, st, sm_count);
        case 2: return launch_ws_k<false, EPI, 2, true>(a, p, st, sm_count);
        case 4: return launch_ws_k<false, EPI, 4, true>(a, p, st, sm_count);
        default: return cudaErrorInvalidConfiguration;
        }
    switch (p.nplanes >> 1) {
    case 0: return launch_ws_k<PHASE, EPI, 0>(a, p, st, sm_count);
    case 1: return launch_ws_k<PHASE, EPI, 1>(a, p, st, sm_count);
    case 2: return launch_ws_k<PHASE, EPI, 2>(a, p, st, sm_count);
    case 4: return launch_ws_k<PHASE, EPI, 4>(a, p, st, sm_count);
    case 8: return launch_ws_k<PHASE, EPI, 8>(a, p, st, sm_count);
    default: return cudaErrorInvalidConfiguration;
    }
}

template <bool PHASE>
static cudaError_t launch_ws_epi(const ConvArgs &a, WsParams &p, cudaStream_t st, int sm_count, bool tma = false)
{
    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE: return launch_ws<PHASE, EPI_F_RNE>(a, p, st, sm_count, tma);
    case EPI_F_RNE_NOHI: return launch_ws<PHASE, EPI_F_RNE_NOHI>(a, p, st, sm_count, tma);
    case EPI_P:     return launch_ws<PHASE, EPI_P>(a, p, st, sm_count, tma);
    default:        return launch_ws<PHASE, EPI_GENERIC>(a, p, st, sm_count, tma);
    }
}

cudaError_t conv3x3_ws(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    if (!ws_shape_ok(a)) return cudaErrorInvalidValue;
    WsParams p;
    memset(&p, 0, sizeof p);
    if (a.q.pool && ws_plan(a, true, &p)) return launch_ws_epi<true>(a, p, st, sm_count);
    // 256 input channels on a narrow un-pooled map (pred): the streamed, raster-tiled variant beats resident weights with
    // cp.async producers even though the weights would fit (the layer is MMA-bound; the raster tile wastes fewer rows)
    memset(&p, 0, sizeof p);
    if (ws_stream_enabled() && a.cs_in == 256 && ((uintptr_t)a.in & 15) == 0 && ws_plan_stream(a, &p) && p.raster)
        return launch_ws_epi<false>(a, p, st, sm_count, true);
    // un-phased tiles: TMA-fed halo where the channel count allows it (YOLO_B200_WS_TMA=0 keeps the cp.async producers)
    static const bool use_tma = [] { const char *e = getenv("YOLO_B200_WS_TMA"); return e ? atoi(e) != 0 : true; }();
    memset(&p, 0, sizeof p);
    if (use_tma && ((uintptr_t)a.in & 15) == 0 && ws_plan(a, false, &p, true)) return launch_ws_epi<false>(a, p, st, sm_count, true);
    memset(&p, 0, sizeof p);
    if (ws_plan(a, false, &p)) return launch_ws_epi<false>(a, p, st, sm_count);
    // weights too large to stay resident: stream them tap by tap (YOLO_B200_WS_STREAM=0: conv_umma.cu takes these layers)
    memset(&p, 0, sizeof p);
    if (ws_stream_enabled() && ((uintptr_t)a.in & 15) == 0 && ws_plan_stream(a, &p)) return launch_ws_epi<false>(a, p, st, sm_count, true);
    return cudaErrorInvalidConfiguration;
}

}  // namespace yb
