// conv_wsp.cu — pair-streamed tcgen05 convolution for the deep, narrow layers whose weights do not fit shared memory
// (slim_yolo_v2 conv5 / conv6 / conv7: 128 / 256 -> 256 channels on 26 x 26 maps; conv_normal, c_embedding/yolo_forward.c:575-770,
// models/slim_yolo_v2.py:283-316).
//
// conv_ws.cu streams these layers' weights (295 / 590 KB) through a shared-memory ring once per 128-pixel tile: 11.3 k cycles
// per tile where the 72 MMAs need 9.4 k (clock64 timeline, profiles/README.md).  What bounds it is the SHARED-MEMORY port: an
// M = 128, K = 32 MMA reads (128 + N) x 32 operand bytes (that is exactly the measured 43 / 51 / 67 cycles at N = 32 / 64 / 128:
// 32 + N/4 at 128 B/clk), and the streamed weights are written through the same port: 72 x 12 KB + 590 KB + 55 KB per tile =
// 11.8 k cycles.  Here a CTA works on a PAIR of 128-pixel raster tiles at once and walks the output channels in two halves of
// 128: every weight chunk (one tap x 128 input channels x 128 output channels = 16 KB) feeds EIGHT MMAs (4 K-steps x 2
// tiles) instead of four, so the weights cross the L2 and the shared-memory port once per 256 pixels (288 x 8 KB + 590 KB +
// 110 KB per pair = 11.7 k cycles per tile of port time; measured 10.9-11.1 k).  Two tiles x 128 columns x two passes in
// flight = the 512 TMEM columns: while pass (pair, half) accumulates, the two epilogue groups (one per tile) drain the previous
// pass.  (Beyond this: cta_group::2, where each SM of a pair reads half of B.)
//
//   * A (activations): flattened-raster halo tiles as in conv_ws.cu (rows of W + 1 pixels written by swizzled TMA boxes,
//     one 128-byte channel plane per buffer, every tap a descriptor start offset).  The plane buffers form ONE in-order
//     ring of four: a pair needs 2 x planes of them and the K loop runs plane-major (all nine taps of plane 0, then plane
//     1), so with 256 input channels the pair's plane-0 buffers are released halfway through its last pass and refilled for
//     the next pair while plane 1 is still being multiplied.
//   * B (weights): chunk-major image [half][plane][tap][128 channels / 8][8 K chunks][8][16 B], one bulk copy per chunk
//     through a ring of up to eight 16 KB slots; a slot is released by a tcgen05.commit.
//   * requantisation epilogue: two groups of eight warps, group g drains tile g of the pair (TMEM -> registers -> exact fp32
//     requantisation -> 16-byte stores), 64 columns per warp.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <cstdio>
#include <cstdlib>

namespace yb {

#ifdef YB_WS_TIMELINE
#define WP_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 32 && lane == 0) p.dbg[it * 8 + (slot)] = clock64(); } while (0)
#else
#define WP_STAMP(slot) do { } while (0)
#endif

constexpr int WP_THREADS = 640;               // warp 0 MMA, warp 1 halo TMA, warp 2 weight copies, warp 3 idle, warps 4-19 epilogue
constexpr int WP_PBUF = 4;                    // plane buffers (in-order ring)
constexpr int WP_MAX_BSLOTS = 8;
constexpr uint32_t WP_CHUNK = 128u * 128u;    // bytes of one weight chunk: 128 output channels x 128 input channels

struct WpParams {
    int n_img, H, W;
    int npl;                     // 128-channel planes of the input (1 or 2)
    int period;                  // H + 1 canvas rows per image (one gutter row = the zero padding between images)
    unsigned period_magic;
    int canvas_rows;
    int rP;                      // W + 1 pixels per raster row
    unsigned rP_magic;
    int raster_rows;             // rows of a halo buffer
    int num_tiles, num_pairs;
    uint32_t plane_bytes;        // one plane buffer (1024-byte multiple): [zero pixel][raster_rows * rP pixels] x 128 B
    int b_slots;
    uint32_t off_plane, off_bias, off_bar;
    int cs_out;                  // 256
    LayerQ q;
    EpiConst k;
    const uint8_t *wtap2;        // [half][plane][tap] chunks of WP_CHUNK bytes
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    long long *dbg;
};

struct WpMaps { CUtensorMap m[5]; };          // box heights 1, 2, 4, 8, 16 rows

template <int EPI>
__global__ void __launch_bounds__(WP_THREADS, 1) conv3x3_wsp_kernel(const WpParams p, const __grid_constant__ WpMaps maps)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t ring0 = base;                                          // weight ring
    const uint32_t plane0 = base + p.off_plane;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    auto bar_pfull = [&](int i) { return bar0 + 8u * i; };
    auto bar_pempty = [&](int i) { return bar0 + 8u * (WP_PBUF + i); };
    auto bar_bfull = [&](int i) { return bar0 + 8u * (2 * WP_PBUF + i); };
    auto bar_bempty = [&](int i) { return bar0 + 8u * (2 * WP_PBUF + WP_MAX_BSLOTS + i); };
    auto bar_tfull = [&](int set, int j) { return bar0 + 8u * (2 * WP_PBUF + 2 * WP_MAX_BSLOTS + 2 * set + j); };
    auto bar_tempty = [&](int set, int j) { return bar0 + 8u * (2 * WP_PBUF + 2 * WP_MAX_BSLOTS + 4 + 2 * set + j); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * WP_PBUF + 2 * WP_MAX_BSLOTS + 8);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (2 * WP_PBUF + 2 * WP_MAX_BSLOTS + 8));

    if (threadIdx.x == 0) {
        for (int i = 0; i < WP_PBUF; ++i) { mbar_init(bar_pfull(i), 1); mbar_init(bar_pempty(i), 1); }
        for (int i = 0; i < WP_MAX_BSLOTS; ++i) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bempty(i), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(bar_tfull(i >> 1, i & 1), 1); mbar_init(bar_tempty(i >> 1, i & 1), 256); }
        fence_barrier_init();
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    // the zero pixel in front of every plane buffer (never written afterwards)
    for (int i = threadIdx.x; i < WP_PBUF * 32; i += blockDim.x)
        reinterpret_cast<uint32_t *>(base_ptr + p.off_plane + (uint32_t)(i >> 5) * p.plane_bytes)[i & 31] = 0u;
    fence_proxy_async();
    for (int i = threadIdx.x; i < p.cs_out; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();
    const int per_pair = 2 * p.npl;                                       // plane buffers a pair occupies

    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // M = 128, N = 128
        const uint32_t ahi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);    // SBO = 8 pixels of 128 B, swizzle 128B
        const uint32_t bhi = (8u * 8u) | (1u << 14);                          // SBO = 8 K chunks x 128 B between 8-channel groups
        uint32_t tapoff[9];                                                   // tap (kh, kw) -> pixels from the tile's stream origin
#pragma unroll
        for (int t = 0; t < 9; ++t) tapoff[t] = (uint32_t)((t / 3) * p.rP + (t % 3)) * 8u;   // x 128 B in 16-byte units
        int bslot = 0;
        uint32_t bph = 0;
        int it = 0;
        for (int pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x, ++it) {
            // stream origins of the two tiles inside their plane buffers: pixel 128 * tile sits toff pixels into the second halo row
            uint32_t org[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int q0 = 128 * (2 * pair + j);
                org[j] = (uint32_t)(q0 - (int)__umulhi((unsigned)q0, p.rP_magic) * p.rP) * 8u;
            }
            for (int nh = 0; nh < 2; ++nh) {
                const int pp = 2 * it + nh, set = pp & 1;
                const uint32_t tph = (uint32_t)(pp >> 1) & 1u;
                mbar_wait(bar_tempty(set, 0), tph ^ 1u);
                mbar_wait(bar_tempty(set, 1), tph ^ 1u);
                if (nh == 0) WP_STAMP(0);
                if (nh == 0) WP_STAMP(1);
                const uint32_t dA = tmem_base + (uint32_t)set * 256u, dB = dA + 128u;
                // ONE elected lane runs the whole pass, barrier polls included: re-electing per chunk (wait, fence, elect, issue,
                // __syncwarp) left the tensor pipe idle ~95 cycles per chunk (633 cycles per 8 MMAs against 8 x 67, clock64 timeline)
                if (elect_one()) {
                    int sl = bslot;
                    uint32_t sph = bph;
                    for (int pl = 0; pl < p.npl; ++pl) {
                        const int cA = it * per_pair + 2 * pl, cB = cA + 1;  // running plane-buffer numbers of (tile A, pl), (tile B, pl)
                        if (nh == 0) {
                            mbar_wait(bar_pfull(cA & 3), (uint32_t)(cA >> 2) & 1u);
                            mbar_wait(bar_pfull(cB & 3), (uint32_t)(cB >> 2) & 1u);
                        }
                        const uint32_t aA = ((plane0 + (uint32_t)(cA & 3) * p.plane_bytes) >> 4) + org[0];
                        const uint32_t aB = ((plane0 + (uint32_t)(cB & 3) * p.plane_bytes) >> 4) + org[1];
                        // the NEXT chunk's barrier is polled between this chunk's sixth and seventh MMA: the pipe's queue is only one
                        // or two instructions deep, so anything slower than that between two MMAs opens a bubble (74 against 67 cycles
                        // per MMA with the poll at the chunk boundary, clock64 timeline)
                        if (pl == 0) { mbar_wait(bar_bfull(sl), sph); tc_fence_after(); }
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t b16 = (ring0 + (uint32_t)sl * WP_CHUNK) >> 4;
                            const bool first = pl == 0 && tap == 0;
                            int nsl = sl + 1;
                            uint32_t nph = sph;
                            if (nsl == p.b_slots) { nsl = 0; nph ^= 1u; }
#pragma unroll
                            for (int c2 = 0; c2 < 4; ++c2) {
                                const uint32_t alo = (aA + tapoff[tap] + (uint32_t)c2 * 2u) | (1u << 16);
                                const uint32_t blo = (b16 + (uint32_t)c2 * 16u) | (8u << 16);
                                if (c2 == 0 && first) umma_i8_lohi<false>(dA, alo, ahi, blo, bhi, idesc);
                                else umma_i8_lohi<true>(dA, alo, ahi, blo, bhi, idesc);
                            }
#pragma unroll
                            for (int c2 = 0; c2 < 4; ++c2) {
                                if (c2 == 2 && !(pl == p.npl - 1 && tap == 8)) { mbar_wait(bar_bfull(nsl), nph); tc_fence_after(); }   // next chunk of this pass
                                const uint32_t alo = (aB + tapoff[tap] + (uint32_t)c2 * 2u) | (1u << 16);
                                const uint32_t blo = (b16 + (uint32_t)c2 * 16u) | (8u << 16);
                                if (c2 == 0 && first) umma_i8_lohi<false>(dB, alo, ahi, blo, bhi, idesc);
                                else umma_i8_lohi<true>(dB, alo, ahi, blo, bhi, idesc);
                            }
                            umma_commit(bar_bempty(sl));
                            if (nh == 1 && tap == 8) { umma_commit(bar_pempty(cA & 3)); umma_commit(bar_pempty(cB & 3)); }   // last use of this plane
                            if (pl == p.npl - 1 && tap == 8) { umma_commit(bar_tfull(set, 0)); umma_commit(bar_tfull(set, 1)); }
                            sl = nsl; sph = nph;
                        }
                    }
                }
                __syncwarp();
                bslot += 9 * p.npl;                                        // every lane keeps the ring position (the elected lane may change)
                while (bslot >= p.b_slots) { bslot -= p.b_slots; bph ^= 1u; }
                if (nh == 1) WP_STAMP(2);
            }
        }
    } else if (warp == 1) {
        // ===================== halo producer (one lane): plane buffers in ring order =====================
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)p.rP * 128u;
            int pc = 0, it = 0;
            for (int pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x, ++it)
                for (int pl = 0; pl < p.npl; ++pl)
                    for (int j = 0; j < 2; ++j, ++pc) {
                        const int tile = 2 * pair + j, idx = pc & 3;
                        mbar_wait(bar_pempty(idx), ((uint32_t)(pc >> 2) & 1u) ^ 1u);
                        if (pl == 0 && j == 0) WP_STAMP(3);
                        if (tile >= p.num_tiles) { mbar_arrive(bar_pfull(idx)); continue; }   // odd tile count: the pair's second tile is a dummy
                        const uint32_t dst = plane0 + (uint32_t)idx * p.plane_bytes + 128u;
                        mbar_expect_tx(bar_pfull(idx), (uint32_t)p.raster_rows * row_bytes);
                        // raster_rows whole canvas rows from the row above stream pixel 128 * tile: runs of rows of one image (its
                        // gutter row and the rows above / below the canvas are out of bounds for the box = zero = the padding)
                        int r = 0, cy = (int)__umulhi((unsigned)(128 * tile), p.rP_magic) - 1;
                        while (r < p.raster_rows) {
                            const int n = cy < 0 ? 0 : (int)__umulhi((unsigned)cy, p.period_magic);
                            int y = cy - n * p.period;
                            int run = min(p.raster_rows - r, p.period - y);
                            while (run > 0) {
                                const int lg = run >= 16 ? 4 : run >= 8 ? 3 : run >= 4 ? 2 : run >= 2 ? 1 : 0, h = 1 << lg;
                                tma_load_4d(dst + (uint32_t)r * row_bytes, &maps.m[lg], bar_pfull(idx), 128 * pl, 0, y, n);
                                r += h; y += h; cy += h; run -= h;
                            }
                        }
                        if (pl == p.npl - 1 && j == 1) WP_STAMP(4);
                    }
        }
    } else if (warp == 2) {
        // ===================== weight producer (one lane): chunks in the order the MMA warp consumes them =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            const int nchunks = 9 * p.npl;
            for (int pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x)
                for (int nh = 0; nh < 2; ++nh)
                    for (int ck = 0; ck < nchunks; ++ck) {
                        mbar_wait(bar_bempty(slot), ph ^ 1u);
                        mbar_expect_tx(bar_bfull(slot), WP_CHUNK);
                        bulk_load_1d(ring0 + (uint32_t)slot * WP_CHUNK, p.wtap2 + (size_t)(nh * nchunks + ck) * WP_CHUNK, WP_CHUNK, bar_bfull(slot));
                        if (++slot == p.b_slots) { slot = 0; ph ^= 1u; }
                    }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: group g drains tile g of every pair =====================
        const int ew_all = warp - 4, g = ew_all >> 3, ew = ew_all & 7;
        const int q4 = warp & 3;                                   // TMEM lane quarter
        const int cb = ew < 4 ? 0 : 64;                            // this warp's 64 of the pass's 128 columns
        unsigned ovf = 0;
        int it = 0;
        for (int pair = blockIdx.x; pair < p.num_pairs; pair += gridDim.x, ++it) {
            const int tile = 2 * pair + g;
            const int q = 128 * tile + q4 * 32 + lane;             // stream pixel of this lane's row
            const int cy = (int)__umulhi((unsigned)q, p.rP_magic), x = q - cy * p.rP;
            const int n = (int)__umulhi((unsigned)cy, p.period_magic), y = cy - n * p.period;
            const bool inside = tile < p.num_tiles && cy < p.canvas_rows && y < p.H && x < p.W;
            int8_t *dst = p.out + (((size_t)n * p.H + y) * p.W + x) * p.cs_out + cb;
            for (int nh = 0; nh < 2; ++nh) {
                const int pp = 2 * it + nh, set = pp & 1;
                mbar_wait(bar_tfull(set, g), (uint32_t)(pp >> 1) & 1u);
                if (ew_all == 0 && nh == 0) WP_STAMP(6);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t)set * 256u + (uint32_t)g * 128u + (uint32_t)cb + ((uint32_t)(q4 * 32) << 16);
                const int ch0 = nh * 128 + cb;                     // first output channel of this warp's columns (dst already points at cb)
                int va[16], vb[16];
                tmem_ld16(taddr, va);
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    tmem_ld_wait();
                    tmem_ld16(taddr + c0 + 16, vb);
                    uint4 w = p.q.activ ? requant16<EPI, true>(va, s_bias, ch0 + c0, p, ovf, inside) : requant16<EPI, false>(va, s_bias, ch0 + c0, p, ovf, inside);
                    if (inside) *reinterpret_cast<uint4 *>(dst + nh * 128 + c0) = w;
                    tmem_ld_wait();
                    if (c0 + 32 < 64) tmem_ld16(taddr + c0 + 32, va);
                    w = p.q.activ ? requant16<EPI, true>(vb, s_bias, ch0 + c0 + 16, p, ovf, inside) : requant16<EPI, false>(vb, s_bias, ch0 + c0 + 16, p, ovf, inside);
                    if (inside) *reinterpret_cast<uint4 *>(dst + nh * 128 + c0 + 16) = w;
                }
                tc_fence_before();
                mbar_arrive(bar_tempty(set, g));
                if (ew_all == 0 && nh == 1) WP_STAMP(7);
            }
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
static bool wsp_enabled()
{
    static const bool on = [] { const char *e = getenv("YOLO_B200_WS_PAIR"); return e ? atoi(e) != 0 : true; }();
    return on;
}

static bool wsp_plan(const ConvArgs &a, WpParams *p)
{
    if (!wsp_enabled() || !a.wimg_tap2) return false;
    if ((a.cs_in != 128 && a.cs_in != 256) || a.cs_out != 256 || a.q.pool) return false;
    if (a.W % 8 == 0 || a.W + 1 > 64) return false;                       // the narrow maps conv_ws.cu runs as flattened rasters
    if ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) != 0) return false;
    if ((long long)a.n * (a.H + 1) * (a.W + 1) >= (1ll << 31) - 256) return false;
    if ((long long)a.n * (a.H + 2) * (a.H + 2) >= (1ll << 32)) return false;   // exactness range of the multiply-high divisions
    memset(p, 0, sizeof *p);
    p->n_img = a.n; p->H = a.H; p->W = a.W; p->npl = a.cs_in / 128;
    p->period = a.H + 1;
    p->period_magic = (unsigned)(((1ull << 32) + (unsigned)p->period - 1) / (unsigned)p->period);
    p->canvas_rows = a.n * p->period;
    p->rP = a.W + 1;
    p->rP_magic = (unsigned)(((1ull << 32) + (unsigned)p->rP - 1) / (unsigned)p->rP);
    p->raster_rows = (3 * p->rP + 127) / p->rP + 1;
    p->num_tiles = (int)(((long long)p->canvas_rows * p->rP + 127) / 128);
    p->num_pairs = (p->num_tiles + 1) / 2;
    p->plane_bytes = (128u + (uint32_t)(p->raster_rows * p->rP) * 128u + 1023u) & ~1023u;
    const uint32_t budget = 227u * 1024u, tail = (uint32_t)a.cs_out * 4u + 512u + 1024u;
    if (WP_PBUF * p->plane_bytes + 3 * WP_CHUNK + tail > budget) return false;
    int slots = (int)((budget - tail - WP_PBUF * p->plane_bytes) / WP_CHUNK);
    p->b_slots = slots > WP_MAX_BSLOTS ? WP_MAX_BSLOTS : slots;
    p->off_plane = (uint32_t)p->b_slots * WP_CHUNK;                       // 16 KB multiples: 1024-byte aligned
    p->off_bias = p->off_plane + WP_PBUF * p->plane_bytes;
    p->off_bar = (p->off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    p->cs_out = a.cs_out; p->q = a.q; p->wtap2 = a.wimg_tap2; p->bias_sh = a.bias_sh; p->out = a.out; p->ovf = a.ovf;
    return true;
}

bool conv3x3_wsp_supported(const ConvArgs &a)
{
    WpParams p;
    return wsp_plan(a, &p);
}

typedef CUresult (*WpEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int EPI>
static cudaError_t launch_wsp(WpParams &p, const WpMaps &maps, cudaStream_t st, int sm_count)
{
#ifdef YB_WS_TIMELINE
    {
        static long long *dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 32 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 32 * 8 * sizeof(long long), st);
        p.dbg = dbg;
    }
#endif
    const uint32_t smem_bytes = p.off_bar + 8u * (2 * WP_PBUF + 2 * WP_MAX_BSLOTS + 10) + 1024u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_wsp_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    const int grid = p.num_pairs < sm_count ? p.num_pairs : sm_count;
    { cudaError_t le = launch_pdl(conv3x3_wsp_kernel<EPI>, dim3(grid), dim3(WP_THREADS), smem_bytes, st, p, maps); if (le != cudaSuccess) return le; }
#ifdef YB_WS_TIMELINE
    {
        long long h[32 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.dbg, sizeof h, cudaMemcpyDeviceToHost);
        const long long t0 = h[3] ? h[3] : h[0];
        printf("WSP timeline npl=%d b_slots=%d pairs=%d (cycles since first stamp): pair | mma: tempty_ok planes_ok issued | prod: empty_ok issued | epi: tfull_ok done\n", p.npl, p.b_slots, p.num_pairs);
        for (int i = 0; i < 8; ++i)
            printf("  %2d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0, h[i*8+6]-t0, h[i*8+7]-t0);
    }
#endif
    return cudaGetLastError();
}

cudaError_t conv3x3_wsp(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    WpParams p;
    if (!wsp_plan(a, &p)) return cudaErrorInvalidValue;
    static WpEncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        enc = (WpEncodeTiledFn)fn;
    }
    WpMaps maps;
    memset(&maps, 0, sizeof maps);
    cuuint64_t dims[4] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n };
    cuuint64_t strides[3] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.cs_in * a.W, (cuuint64_t)a.cs_in * a.W * a.H };
    cuuint32_t es[4] = { 1, 1, 1, 1 };
    for (int i = 0; i < 5; ++i) {
        cuuint32_t box[4] = { 128, (cuuint32_t)p.rP, (cuuint32_t)(1 << i), 1 };
        if (enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE:      return launch_wsp<EPI_F_RNE>(p, maps, st, sm_count);
    case EPI_F_RNE_NOHI: return launch_wsp<EPI_F_RNE_NOHI>(p, maps, st, sm_count);
    case EPI_P:          return launch_wsp<EPI_P>(p, maps, st, sm_count);
    default:             return launch_wsp<EPI_GENERIC>(p, maps, st, sm_count);
    }
}

}  // namespace yb
