// conv_ws2.cu — CTA-pair (cta_group::2) tcgen05 convolution for the deep, narrow layers whose weights do not fit shared memory
// (slim_yolo_v2 conv5 / conv6 / conv7: 128 / 256 -> 256 channels on 26 x 26 maps; conv_normal, c_embedding/yolo_forward.c:575-770,
// models/slim_yolo_v2.py:283-316).
//
// These layers are bound by the SHARED-MEMORY port of an SM: an M = 128, N = 256, K = 32 MMA reads 12 KB of operands, and the
// streamed weights (295 / 590 KB per 128-pixel tile) are written through the same port (conv_ws.cu: 11.8 k cycles of port time
// per tile against 9.4 k cycles of math; conv_wsp.cu shares each weight chunk between two tiles of one CTA and lands at
// 10.9 k).  Here the two SMs of a TPC form a CTA pair: each CTA owns ONE 128-pixel raster tile (the A operand, its halo planes)
// and HALF of the output channels' weights (the B operand, 128 of the 256 rows of every chunk); one thread of the leader CTA
// issues tcgen05.mma.cta_group::2 with M = 256, N = 256: both tensor cores run, each reading its own A tile and its own half
// of B, and the accumulators land in each CTA's own TMEM (128 lanes x 256 columns).  Per tile and SM: 72 x 8 KB of operand
// reads + 295 KB of weights + 55 KB of halo = 7.2 k cycles of port time, below the 9.4 k cycles of math.
//
// Protocol (all barriers sit at the same shared-memory offsets in both CTAs):
//   * each CTA's producers fill their own plane stage / weight slot and complete their LOCAL full barrier; in the peer CTA a
//     relay thread forwards every completion to the leader's `peer` barrier with a remote mbarrier.arrive (mapa), so the
//     issuing thread waits for (own full, peer's full);
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arrives on the SAME barrier in both CTAs: slot / stage release and
//     "accumulator complete" reach each CTA's own producers and epilogue warps;
//   * the epilogue warps of both CTAs drain their own TMEM; the leader's tmem-empty barrier also counts one relayed arrival
//     from the peer, so a buffer is re-used only when both halves are drained.
// Both CTAs place their halo rows so that the tile's stream origin sits at the SAME offset of the plane buffer (the TMA
// destination is shifted by the tile's raster offset), because one A descriptor serves both.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <cstdio>
#include <cstdlib>

namespace yb {

#ifdef YB_WS_TIMELINE
#define W2_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it < 32 && lane == 0) p.dbg[it * 8 + (slot)] = clock64(); } while (0)
#else
#define W2_STAMP(slot) do { } while (0)
#endif

constexpr int W2_THREADS = 640;     // warp 0 MMA issuer (leader) / weight relay (peer), warp 1 halo TMA, warp 2 weight copies, warp 3 relays (peer), warps 4-19 epilogue
constexpr int W2_MAX_BSLOTS = 8;
constexpr uint32_t W2_CHUNK = 128u * 128u;    // this CTA's half of one (tap, 128-channel plane) weight chunk

struct W2Params {
    int n_img, H, W;
    int npl;                     // 128-channel planes of the input (1 or 2)
    int period;                  // H + 1 canvas rows per image
    unsigned period_magic;
    int canvas_rows;
    int rP;                      // W + 1 pixels per raster row
    unsigned rP_magic;
    int raster_rows;
    int num_tiles, num_pairs;
    uint32_t plane_bytes;        // one 128-channel plane of a stage
    uint32_t stage_bytes;        // npl planes
    int b_slots;
    uint32_t off_stage, off_bias, off_bar;
    int cs_out;
    LayerQ q;
    EpiConst k;
    const uint8_t *wtap2;        // [half][plane][tap] chunks of W2_CHUNK bytes (the image conv_wsp.cu uses)
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
    long long *dbg;
};

struct W2Maps { CUtensorMap m[5]; CUtensorMap px; };          // row boxes of 1, 2, 4, 8, 16 rows; one pixel

template <int EPI>
__global__ void __launch_bounds__(W2_THREADS, 1) conv3x3_ws2_kernel(const W2Params p, const __grid_constant__ W2Maps maps)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_count_x();

    const uint32_t ring0 = base;
    const uint32_t stage0 = base + p.off_stage;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    auto bar_pfull = [&](int s) { return bar0 + 8u * s; };
    auto bar_ppeer = [&](int s) { return bar0 + 8u * (2 + s); };
    auto bar_pempty = [&](int s) { return bar0 + 8u * (4 + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (6 + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (8 + b); };
    auto bar_bfull = [&](int i) { return bar0 + 8u * (10 + i); };
    auto bar_bpeer = [&](int i) { return bar0 + 8u * (10 + W2_MAX_BSLOTS + i); };
    auto bar_bempty = [&](int i) { return bar0 + 8u * (10 + 2 * W2_MAX_BSLOTS + i); };
    const uint32_t tmem_slot = bar0 + 8u * (10 + 3 * W2_MAX_BSLOTS);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (10 + 3 * W2_MAX_BSLOTS));

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_pfull(s), 1); mbar_init(bar_ppeer(s), 1); mbar_init(bar_pempty(s), 1);
            mbar_init(bar_tfull(s), 1);
            mbar_init(bar_tempty(s), rank == 0 ? 257 : 256);      // the leader also counts the peer's relayed "drained"
        }
        for (int i = 0; i < W2_MAX_BSLOTS; ++i) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bpeer(i), 1); mbar_init(bar_bempty(i), 1); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < p.cs_out; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;
    }
    __syncthreads();
    cluster_sync_all();                                           // both CTAs' barriers exist before anything arrives remotely
    if (warp == 0) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();
    const int nchunks = 9 * p.npl;

    if (warp == 0 && rank == 0) {
        // ===================== MMA issuer: one thread of the leader CTA drives both tensor cores =====================
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);   // M = 256 (2 x 128), N = 256
        const uint32_t ahi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);    // SBO = 8 pixels of 128 B, swizzle 128B
        const uint32_t bhi = (8u * 8u) | (1u << 14);                          // SBO = 8 K chunks x 128 B between 8-channel groups
        uint32_t tapoff[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) tapoff[t] = (uint32_t)((p.rP - 1) + (t / 3) * p.rP + (t % 3)) * 8u;   // from the plane start, 16-byte units
        int bslot = 0;
        uint32_t bph = 0;
        int it = 0;
        for (int pair = cid; pair < p.num_pairs; pair += ncl, ++it) {
            const int buf = it & 1, s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(bar_tempty(buf), ph ^ 1u);
            W2_STAMP(0);
            mbar_wait(bar_pfull(s), ph);
            mbar_wait(bar_ppeer(s), ph);
            W2_STAMP(1);
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)buf * 256u;
                int sl = bslot;
                uint32_t sph = bph;
                for (int pl = 0; pl < p.npl; ++pl) {
                    const uint32_t a16 = (stage0 + (uint32_t)s * p.stage_bytes + (uint32_t)pl * p.plane_bytes) >> 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(bar_bfull(sl), sph);
                        mbar_wait(bar_bpeer(sl), sph);
                        tc_fence_after();
                        const uint32_t b16 = (ring0 + (uint32_t)sl * W2_CHUNK) >> 4;
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) {
                            const uint32_t alo = (a16 + tapoff[tap] + (uint32_t)c2 * 2u) | (1u << 16);
                            const uint32_t blo = (b16 + (uint32_t)c2 * 16u) | (8u << 16);
                            if (c2 == 0 && pl == 0 && tap == 0) umma2_i8_lohi<false>(d, alo, ahi, blo, bhi, idesc);
                            else umma2_i8_lohi<true>(d, alo, ahi, blo, bhi, idesc);
                        }
                        umma2_commit_mc(bar_bempty(sl));
                        if (pl == p.npl - 1 && tap == 8) { umma2_commit_mc(bar_pempty(s)); umma2_commit_mc(bar_tfull(buf)); }
                        if (++sl == p.b_slots) { sl = 0; sph ^= 1u; }
                    }
                }
            }
            __syncwarp();
            bslot += nchunks;
            while (bslot >= p.b_slots) { bslot -= p.b_slots; bph ^= 1u; }
            W2_STAMP(2);
        }
    } else if (warp == 0) {
        // ===================== peer CTA: forwards "my half of the weight chunk has landed" to the leader =====================
        if (lane == 0) {
            int sl = 0;
            uint32_t sph = 0;
            for (int pair = cid; pair < p.num_pairs; pair += ncl)
                for (int ck = 0; ck < nchunks; ++ck) {
                    mbar_wait(bar_bfull(sl), sph);
                    mbar_arrive_remote(bar_bpeer(sl), 0);
                    if (++sl == p.b_slots) { sl = 0; sph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ===================== halo producer (one lane): this CTA's tile =====================
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)p.rP * 128u;
            int it = 0;
            for (int pair = cid; pair < p.num_pairs; pair += ncl, ++it) {
                const int tile = 2 * pair + (int)rank, s = it & 1;
                mbar_wait(bar_pempty(s), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                W2_STAMP(3);
                if (tile >= p.num_tiles) { mbar_arrive(bar_pfull(s)); continue; }       // odd tile count: the pair's second tile is a dummy
                const int cy0 = (int)__umulhi((unsigned)(128 * tile), p.rP_magic);
                const int toff = 128 * tile - cy0 * p.rP;
                mbar_expect_tx(bar_pfull(s), (uint32_t)p.npl * ((uint32_t)p.raster_rows * row_bytes + 128u));
                for (int pl = 0; pl < p.npl; ++pl) {
                    // rows start (rP - toff) pixels into the plane: the tile's stream origin is then at pixel rP - 1 for every tile
                    const uint32_t dst = stage0 + (uint32_t)s * p.stage_bytes + (uint32_t)pl * p.plane_bytes + (uint32_t)(p.rP - toff) * 128u;
                    tma_load_4d(dst - 128u, &maps.px, bar_pfull(s), 128 * pl, p.W, 0, 0);    // the pixel in front: out of bounds = zero
                    int r = 0, cy = cy0 - 1;
                    while (r < p.raster_rows) {
                        const int n = cy < 0 ? 0 : (int)__umulhi((unsigned)cy, p.period_magic);
                        int y = cy - n * p.period;
                        int run = min(p.raster_rows - r, p.period - y);
                        while (run > 0) {
                            const int lg = run >= 16 ? 4 : run >= 8 ? 3 : run >= 4 ? 2 : run >= 2 ? 1 : 0, h = 1 << lg;
                            tma_load_4d(dst + (uint32_t)r * row_bytes, &maps.m[lg], bar_pfull(s), 128 * pl, 0, y, n);
                            r += h; y += h; cy += h; run -= h;
                        }
                    }
                }
                W2_STAMP(4);
            }
        }
    } else if (warp == 2) {
        // ===================== weight producer (one lane): this CTA's half of every chunk =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            const uint8_t *src0 = p.wtap2 + (size_t)rank * nchunks * W2_CHUNK;
            for (int pair = cid; pair < p.num_pairs; pair += ncl)
                for (int ck = 0; ck < nchunks; ++ck) {
                    mbar_wait(bar_bempty(slot), ph ^ 1u);
                    mbar_expect_tx(bar_bfull(slot), W2_CHUNK);
                    bulk_load_1d(ring0 + (uint32_t)slot * W2_CHUNK, src0 + (size_t)ck * W2_CHUNK, W2_CHUNK, bar_bfull(slot));
                    if (++slot == p.b_slots) { slot = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 3) {
        // ===================== peer CTA: relays of "my halo stage is loaded" (lane 0) and "my accumulator is drained" (lane 1) =====================
        if (rank == 1 && lane < 2) {
            int it = 0;
            for (int pair = cid; pair < p.num_pairs; pair += ncl, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                if (lane == 0) { mbar_wait(bar_pfull(s), ph); mbar_arrive_remote(bar_ppeer(s), 0); }
                else { mbar_wait(bar_tempty(s), ph); mbar_arrive_remote(bar_tempty(s), 0); }
            }
        }
    } else {
        // ===================== epilogue: group g drains accumulator buffer g (tiles it = g, g + 2, ...) of this CTA's TMEM =====================
        const int ew_all = warp - 4, g = ew_all >> 3, ew = ew_all & 7;
        const int q4 = warp & 3;
        const int cb = ew < 4 ? 0 : 128;                           // this warp's 128 of the 256 columns
        unsigned ovf = 0;
        int it = g;
        for (int pair = cid + g * ncl; pair < p.num_pairs; pair += 2 * ncl, it += 2) {
            const int tile = 2 * pair + (int)rank, buf = g;
            const int q = 128 * tile + q4 * 32 + lane;
            const int cy = (int)__umulhi((unsigned)q, p.rP_magic), x = q - cy * p.rP;
            const int n = (int)__umulhi((unsigned)cy, p.period_magic), y = cy - n * p.period;
            const bool inside = tile < p.num_tiles && cy < p.canvas_rows && y < p.H && x < p.W;
            int8_t *dst = p.out + (((size_t)n * p.H + y) * p.W + x) * p.cs_out + cb;
            mbar_wait(bar_tfull(buf), (uint32_t)(it >> 1) & 1u);
            if (ew_all == 0) W2_STAMP(6);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * 256u + (uint32_t)cb + ((uint32_t)(q4 * 32) << 16);
            int va[16], vb[16];
            tmem_ld16(taddr, va);
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                tmem_ld_wait();
                tmem_ld16(taddr + c0 + 16, vb);
                uint4 w = p.q.activ ? requant16<EPI, true>(va, s_bias, cb + c0, p, ovf, inside) : requant16<EPI, false>(va, s_bias, cb + c0, p, ovf, inside);
                if (inside) *reinterpret_cast<uint4 *>(dst + c0) = w;
                tmem_ld_wait();
                if (c0 + 32 < 128) tmem_ld16(taddr + c0 + 32, va);
                w = p.q.activ ? requant16<EPI, true>(vb, s_bias, cb + c0 + 16, p, ovf, inside) : requant16<EPI, false>(vb, s_bias, cb + c0 + 16, p, ovf, inside);
                if (inside) *reinterpret_cast<uint4 *>(dst + c0 + 16) = w;
            }
            tc_fence_before();
            mbar_arrive(bar_tempty(buf));
            if (ew_all == 0) W2_STAMP(7);
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                           // nothing may still arrive in the peer's shared memory / TMEM
    if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static bool ws2_enabled()
{
    static const bool on = [] { const char *e = getenv("YOLO_B200_WS_CTAPAIR"); return e ? atoi(e) != 0 : true; }();
    return on;
}

static bool ws2_plan(const ConvArgs &a, W2Params *p, int sm_count)
{
    if (!ws2_enabled() || !a.wimg_tap2 || sm_count < 2) return false;
    if ((a.cs_in != 128 && a.cs_in != 256) || a.cs_out != 256 || a.q.pool) return false;
    if (a.W % 8 == 0 || a.W + 1 > 64) return false;
    if ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) != 0) return false;
    if ((long long)a.n * (a.H + 1) * (a.W + 1) >= (1ll << 31) - 256) return false;
    if ((long long)a.n * (a.H + 2) * (a.H + 2) >= (1ll << 32)) return false;
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
    p->plane_bytes = ((uint32_t)((p->raster_rows + 1) * p->rP + 2) * 128u + 1023u) & ~1023u;
    p->stage_bytes = (uint32_t)p->npl * p->plane_bytes;
    const uint32_t budget = 227u * 1024u, tail = (uint32_t)a.cs_out * 4u + 512u + 1024u;
    if (2 * p->stage_bytes + 3 * W2_CHUNK + tail > budget) return false;
    int slots = (int)((budget - tail - 2 * p->stage_bytes) / W2_CHUNK);
    p->b_slots = slots > W2_MAX_BSLOTS ? W2_MAX_BSLOTS : slots;
    p->off_stage = (uint32_t)p->b_slots * W2_CHUNK;
    p->off_bias = p->off_stage + 2 * p->stage_bytes;
    p->off_bar = (p->off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    p->cs_out = a.cs_out; p->q = a.q; p->wtap2 = a.wimg_tap2; p->bias_sh = a.bias_sh; p->out = a.out; p->ovf = a.ovf;
    return true;
}

bool conv3x3_ws2_supported(const ConvArgs &a, int sm_count)
{
    W2Params p;
    return ws2_plan(a, &p, sm_count);
}

typedef CUresult (*W2EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int EPI>
static cudaError_t launch_ws2(W2Params &p, const W2Maps &maps, cudaStream_t st, int sm_count)
{
#ifdef YB_WS_TIMELINE
    {
        static long long *dbg = nullptr;
        if (!dbg) cudaMalloc(&dbg, 32 * 8 * sizeof(long long));
        cudaMemsetAsync(dbg, 0, 32 * 8 * sizeof(long long), st);
        p.dbg = dbg;
    }
#endif
    const uint32_t smem_bytes = p.off_bar + 8u * (10 + 3 * W2_MAX_BSLOTS + 2) + 1024u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_ws2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    int ctas = 2 * p.num_pairs < sm_count ? 2 * p.num_pairs : sm_count;
    ctas &= ~1;                                                   // whole CTA pairs
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)ctas, 1, 1);
    cfg.blockDim = dim3(W2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    static const bool pdl_on = [] { const char *e = getenv("YOLO_B200_PDL"); return e ? atoi(e) != 0 : true; }();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_on ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_ws2_kernel<EPI>, p, maps);
    if (e != cudaSuccess) return e;
#ifdef YB_WS_TIMELINE
    {
        long long h[32 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.dbg, sizeof h, cudaMemcpyDeviceToHost);
        const long long t0 = h[3] ? h[3] : h[0];
        printf("WS2 timeline npl=%d b_slots=%d pairs=%d (cycles since first stamp): tile | mma: tempty_ok planes_ok issued | prod: empty_ok issued | epi: tfull_ok done\n", p.npl, p.b_slots, p.num_pairs);
        for (int i = 0; i < 12; ++i)
            printf("  %2d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, h[i*8]-t0, h[i*8+1]-t0, h[i*8+2]-t0, h[i*8+3]-t0, h[i*8+4]-t0, h[i*8+6]-t0, h[i*8+7]-t0);
    }
#endif
    return cudaGetLastError();
}

cudaError_t conv3x3_ws2(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    W2Params p;
    if (!ws2_plan(a, &p, sm_count)) return cudaErrorInvalidValue;
    static W2EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        enc = (W2EncodeTiledFn)fn;
    }
    W2Maps maps;
    memset(&maps, 0, sizeof maps);
    cuuint64_t dims[4] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n };
    cuuint64_t strides[3] = { (cuuint64_t)a.cs_in, (cuuint64_t)a.cs_in * a.W, (cuuint64_t)a.cs_in * a.W * a.H };
    cuuint32_t es[4] = { 1, 1, 1, 1 };
    for (int i = 0; i < 6; ++i) {
        cuuint32_t box[4] = { 128, (cuuint32_t)(i < 5 ? p.rP : 1), (cuuint32_t)(i < 5 ? (1 << i) : 1), 1 };
        if (enc(i < 5 ? &maps.m[i] : &maps.px, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)a.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    switch (epi_mode_for(a, &p.k)) {
    case EPI_F_RNE:      return launch_ws2<EPI_F_RNE>(p, maps, st, sm_count);
    case EPI_F_RNE_NOHI: return launch_ws2<EPI_F_RNE_NOHI>(p, maps, st, sm_count);
    case EPI_P:          return launch_ws2<EPI_P>(p, maps, st, sm_count);
    default:             return launch_ws2<EPI_GENERIC>(p, maps, st, sm_count);
    }
}

}  // namespace yb
