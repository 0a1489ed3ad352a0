// conv_ws3.cu — the CTA-pair (cta_group::2) kernel of conv_ws2.cu generalised to the WIDE 3x3 layers of yolo_v2 / darknet19
// (256 ... 1280 -> 512 / 1024 channels on 26 x 26 and 13 x 13 maps: backbone/darknet.py:78-108, models/yolo_v2.py:29-40,175).
//
// Same pair protocol as conv_ws2.cu (each CTA of a TPC pair owns one 128-pixel raster tile and half of every weight chunk,
// the leader issues tcgen05.mma.cta_group::2 with M = 256, N = 256; completions relayed with remote mbarrier arrivals,
// releases by multicast commits).  What changes:
//   * the unit of work is (tile pair, SLICE of 256 output channels): 13 x 13 maps at batch 64 are only 49 tile pairs, the
//     slices make 196 units for the 74 CTA pairs; a unit's accumulator is one 256-column TMEM buffer, two buffers alternate;
//   * a tile's 128-channel halo planes (4 .. 10 of them, ~25 KB each) do not fit shared memory together, so they are streamed
//     through an in-order ring like the weights: the K loop runs plane-major and a plane buffer is released (multicast
//     commit) after its nine taps;
//   * weights: 16 KB chunks [slice][half][plane][tap] (the host packs them at load).
// Per unit and SM the shared-memory port carries 36 x planes x 8 KB of operand reads + 9 x planes x 16 KB of weights + the
// planes, about 0.78 of the 36 x planes x 131 cycles of math: the layers become MMA-bound where conv_umma.cu (one TMA box per
// tap and channel chunk, every CTA re-reading all of B) reaches a third of the tensor peak.
#include "kernels.h"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <cstdio>
#include <cstdlib>

namespace yb {

constexpr int W3_THREADS = 640;     // warp 0 MMA issuer (leader) / weight relay (peer), warp 1 halo TMA, warp 2 weight copies, warp 3 relays (peer), warps 4-19 epilogue
constexpr int W3_MAX_BSLOTS = 8;
constexpr int W3_MAX_PBUF = 8;
constexpr uint32_t W3_CHUNK = 128u * 128u;

struct W3Params {
    int n_img, H, W;
    int npl;                     // 128-channel planes of the input (cs_in / 128)
    int nslices;                 // cs_out / 256
    int period;                  // H + 1 canvas rows per image
    unsigned period_magic;
    int canvas_rows;
    int rP;                      // W + 1 pixels per raster row
    unsigned rP_magic;
    int raster_rows;
    int num_tiles, num_pairs, num_units;
    unsigned slices_magic;       // ceil(2^32 / nslices)
    uint32_t plane_bytes;
    int p_bufs, b_slots;
    uint32_t off_plane, off_bias, off_bar;
    int cs_out;
    LayerQ q;
    EpiConst k;
    const uint8_t *wtap3;        // [slice][half][plane][tap] chunks of W3_CHUNK bytes
    const int *bias_sh;
    int8_t *out;
    unsigned *ovf;
};

struct W3Maps { CUtensorMap m[5]; CUtensorMap px; };

template <int EPI>
__global__ void __launch_bounds__(W3_THREADS, 1) conv3x3_ws3_kernel(const W3Params p, const __grid_constant__ W3Maps maps)
{
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = (int)cluster_id_x(), ncl = (int)cluster_count_x();

    const uint32_t ring0 = base;
    const uint32_t plane0 = base + p.off_plane;
    int *s_bias = reinterpret_cast<int *>(base_ptr + p.off_bias);
    const uint32_t bar0 = base + p.off_bar;
    auto bar_pfull = [&](int i) { return bar0 + 8u * i; };
    auto bar_ppeer = [&](int i) { return bar0 + 8u * (W3_MAX_PBUF + i); };
    auto bar_pempty = [&](int i) { return bar0 + 8u * (2 * W3_MAX_PBUF + i); };
    auto bar_bfull = [&](int i) { return bar0 + 8u * (3 * W3_MAX_PBUF + i); };
    auto bar_bpeer = [&](int i) { return bar0 + 8u * (3 * W3_MAX_PBUF + W3_MAX_BSLOTS + i); };
    auto bar_bempty = [&](int i) { return bar0 + 8u * (3 * W3_MAX_PBUF + 2 * W3_MAX_BSLOTS + i); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (3 * W3_MAX_PBUF + 3 * W3_MAX_BSLOTS + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (3 * W3_MAX_PBUF + 3 * W3_MAX_BSLOTS + 2 + b); };
    const uint32_t tmem_slot = bar0 + 8u * (3 * W3_MAX_PBUF + 3 * W3_MAX_BSLOTS + 4);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + p.off_bar + 8u * (3 * W3_MAX_PBUF + 3 * W3_MAX_BSLOTS + 4));

    if (threadIdx.x == 0) {
        for (int i = 0; i < W3_MAX_PBUF; ++i) { mbar_init(bar_pfull(i), 1); mbar_init(bar_ppeer(i), 1); mbar_init(bar_pempty(i), 1); }
        for (int i = 0; i < W3_MAX_BSLOTS; ++i) { mbar_init(bar_bfull(i), 1); mbar_init(bar_bpeer(i), 1); mbar_init(bar_bempty(i), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull(b), 1); mbar_init(bar_tempty(b), rank == 0 ? 257 : 256); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < p.cs_out; i += blockDim.x) {
        const int b = p.bias_sh[i];
        s_bias[i] = (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) ? __float_as_int((float)b) : b;
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();
    const int nchunks = 9 * p.npl;                                // weight chunks per unit

    if (warp == 0 && rank == 0) {
        // ===================== MMA issuer (leader CTA) =====================
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
        const uint32_t ahi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t bhi = (8u * 8u) | (1u << 14);
        uint32_t tapoff[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) tapoff[t] = (uint32_t)((p.rP - 1) + (t / 3) * p.rP + (t % 3)) * 8u;
        int bslot = 0, pidx = 0;
        uint32_t bph = 0, pph = 0;
        int it = 0;
        for (int unit = cid; unit < p.num_units; unit += ncl, ++it) {
            const int buf = it & 1;
            mbar_wait(bar_tempty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)buf * 256u;
                int sl = bslot, pi = pidx;
                uint32_t sph = bph, ph = pph;
                for (int pl = 0; pl < p.npl; ++pl) {
                    mbar_wait(bar_pfull(pi), ph);
                    mbar_wait(bar_ppeer(pi), ph);
                    const uint32_t a16 = (plane0 + (uint32_t)pi * p.plane_bytes) >> 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(bar_bfull(sl), sph);
                        mbar_wait(bar_bpeer(sl), sph);
                        tc_fence_after();
                        const uint32_t b16 = (ring0 + (uint32_t)sl * W3_CHUNK) >> 4;
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) {
                            const uint32_t alo = (a16 + tapoff[tap] + (uint32_t)c2 * 2u) | (1u << 16);
                            const uint32_t blo = (b16 + (uint32_t)c2 * 16u) | (8u << 16);
                            if (c2 == 0 && pl == 0 && tap == 0) umma2_i8_lohi<false>(d, alo, ahi, blo, bhi, idesc);
                            else umma2_i8_lohi<true>(d, alo, ahi, blo, bhi, idesc);
                        }
                        umma2_commit_mc(bar_bempty(sl));
                        if (tap == 8) umma2_commit_mc(bar_pempty(pi));                  // last use of this plane buffer
                        if (pl == p.npl - 1 && tap == 8) umma2_commit_mc(bar_tfull(buf));
                        if (++sl == p.b_slots) { sl = 0; sph ^= 1u; }
                    }
                    if (++pi == p.p_bufs) { pi = 0; ph ^= 1u; }
                }
            }
            __syncwarp();
            bslot += nchunks;
            while (bslot >= p.b_slots) { bslot -= p.b_slots; bph ^= 1u; }
            pidx += p.npl;
            while (pidx >= p.p_bufs) { pidx -= p.p_bufs; pph ^= 1u; }
        }
    } else if (warp == 0) {
        // ===================== peer CTA: forwards "my half of the weight chunk has landed" =====================
        if (lane == 0) {
            int sl = 0;
            uint32_t sph = 0;
            for (int unit = cid; unit < p.num_units; unit += ncl)
                for (int ck = 0; ck < nchunks; ++ck) {
                    mbar_wait(bar_bfull(sl), sph);
                    mbar_arrive_remote(bar_bpeer(sl), 0);
                    if (++sl == p.b_slots) { sl = 0; sph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ===================== halo producer (one lane): the planes of this CTA's tile, unit after unit =====================
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)p.rP * 128u;
            int pi = 0;
            uint32_t ph = 0;
            for (int unit = cid; unit < p.num_units; unit += ncl) {
                const int pair = p.nslices == 1 ? unit : (int)__umulhi((unsigned)unit, p.slices_magic);
                const int tile = 2 * pair + (int)rank;
                const int cy0 = (int)__umulhi((unsigned)(128 * tile), p.rP_magic);
                const int toff = 128 * tile - cy0 * p.rP;
                for (int pl = 0; pl < p.npl; ++pl) {
                    mbar_wait(bar_pempty(pi), ph ^ 1u);
                    if (tile >= p.num_tiles) mbar_arrive(bar_pfull(pi));                  // odd tile count: dummy second tile
                    else {
                        mbar_expect_tx(bar_pfull(pi), (uint32_t)p.raster_rows * row_bytes + 128u);
                        const uint32_t dst = plane0 + (uint32_t)pi * p.plane_bytes + (uint32_t)(p.rP - toff) * 128u;
                        tma_load_4d(dst - 128u, &maps.px, bar_pfull(pi), 128 * pl, p.W, 0, 0);   // the pixel in front: out of bounds = zero
                        int r = 0, cy = cy0 - 1;
                        while (r < p.raster_rows) {
                            const int n = cy < 0 ? 0 : (int)__umulhi((unsigned)cy, p.period_magic);
                            int y = cy - n * p.period;
                            int run = min(p.raster_rows - r, p.period - y);
                            while (run > 0) {
                                const int lg = run >= 16 ? 4 : run >= 8 ? 3 : run >= 4 ? 2 : run >= 2 ? 1 : 0, h = 1 << lg;
                                tma_load_4d(dst + (uint32_t)r * row_bytes, &maps.m[lg], bar_pfull(pi), 128 * pl, 0, y, n);
                                r += h; y += h; cy += h; run -= h;
                            }
                        }
                    }
                    if (++pi == p.p_bufs) { pi = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== weight producer (one lane): this CTA's half of every chunk of the unit's slice =====================
        if (lane == 0) {
            int slot = 0;
            uint32_t ph = 0;
            for (int unit = cid; unit < p.num_units; unit += ncl) {
                const int pair = p.nslices == 1 ? unit : (int)__umulhi((unsigned)unit, p.slices_magic);
                const int slice = unit - pair * p.nslices;
                const uint8_t *src0 = p.wtap3 + (size_t)(slice * 2 + (int)rank) * nchunks * W3_CHUNK;
                for (int ck = 0; ck < nchunks; ++ck) {
                    mbar_wait(bar_bempty(slot), ph ^ 1u);
                    mbar_expect_tx(bar_bfull(slot), W3_CHUNK);
                    bulk_load_1d(ring0 + (uint32_t)slot * W3_CHUNK, src0 + (size_t)ck * W3_CHUNK, W3_CHUNK, bar_bfull(slot));
                    if (++slot == p.b_slots) { slot = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== peer CTA: relays of "my plane is loaded" (lane 0) and "my accumulator is drained" (lane 1) =====================
        if (rank == 1 && lane == 0) {
            int pi = 0;
            uint32_t ph = 0;
            for (int unit = cid; unit < p.num_units; unit += ncl)
                for (int pl = 0; pl < p.npl; ++pl) {
                    mbar_wait(bar_pfull(pi), ph);
                    mbar_arrive_remote(bar_ppeer(pi), 0);
                    if (++pi == p.p_bufs) { pi = 0; ph ^= 1u; }
                }
        } else if (rank == 1 && lane == 1) {
            int it = 0;
            for (int unit = cid; unit < p.num_units; unit += ncl, ++it) {
                mbar_wait(bar_tempty(it & 1), (uint32_t)(it >> 1) & 1u);
                mbar_arrive_remote(bar_tempty(it & 1), 0);
            }
        }
    } else {
        // ===================== epilogue: group g drains accumulator buffer g (units it = g, g + 2, ...) =====================
        const int ew_all = warp - 4, g = ew_all >> 3, ew = ew_all & 7;
        const int q4 = warp & 3;
        const int cb = ew < 4 ? 0 : 128;
        unsigned ovf = 0;
        int it = g;
        for (int unit = cid + g * ncl; unit < p.num_units; unit += 2 * ncl, it += 2) {
            const int pair = p.nslices == 1 ? unit : (int)__umulhi((unsigned)unit, p.slices_magic);
            const int slice = unit - pair * p.nslices;
            const int tile = 2 * pair + (int)rank, buf = g;
            const int q = 128 * tile + q4 * 32 + lane;
            const int cy = (int)__umulhi((unsigned)q, p.rP_magic), x = q - cy * p.rP;
            const int n = (int)__umulhi((unsigned)cy, p.period_magic), y = cy - n * p.period;
            const bool inside = tile < p.num_tiles && cy < p.canvas_rows && y < p.H && x < p.W;
            const int ch0 = 256 * slice + cb;
            int8_t *dst = p.out + (((size_t)n * p.H + y) * p.W + x) * p.cs_out + ch0;
            mbar_wait(bar_tfull(buf), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * 256u + (uint32_t)cb + ((uint32_t)(q4 * 32) << 16);
            int va[16], vb[16];
            tmem_ld16(taddr, va);
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                tmem_ld_wait();
                tmem_ld16(taddr + c0 + 16, vb);
                uint4 w = p.q.activ ? requant16<EPI, true>(va, s_bias, ch0 + c0, p, ovf, inside) : requant16<EPI, false>(va, s_bias, ch0 + c0, p, ovf, inside);
                if (inside) *reinterpret_cast<uint4 *>(dst + c0) = w;
                tmem_ld_wait();
                if (c0 + 32 < 128) tmem_ld16(taddr + c0 + 32, va);
                w = p.q.activ ? requant16<EPI, true>(vb, s_bias, ch0 + c0 + 16, p, ovf, inside) : requant16<EPI, false>(vb, s_bias, ch0 + c0 + 16, p, ovf, inside);
                if (inside) *reinterpret_cast<uint4 *>(dst + c0 + 16) = w;
            }
            tc_fence_before();
            mbar_arrive(bar_tempty(buf));
        }
        if (p.q.contract == CONTRACT_P) {
            ovf = __reduce_add_sync(0xffffffffu, ovf);
            if (lane == 0 && ovf) atomicAdd(p.ovf, ovf);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static bool ws3_enabled()
{
    static const bool on = [] { const char *e = getenv("YOLO_B200_WS_WIDE"); return e ? atoi(e) != 0 : true; }();
    return on;
}

static bool ws3_plan(const ConvArgs &a, W3Params *p, int sm_count)
{
    if (!ws3_enabled() || !a.wimg_tap3 || sm_count < 2 || a.taps != 9) return false;
    if (a.cs_in % 128 || a.cs_in < 128 || a.cs_in > 2048 || a.cs_out % 256 || a.cs_out < 256 || a.cs_out > 2048 || a.q.pool) return false;
    if (a.W % 8 == 0 || a.W + 1 > 64) return false;
    if ((((uintptr_t)a.in | (uintptr_t)a.out) & 15) != 0) return false;
    if ((long long)a.n * (a.H + 1) * (a.W + 1) >= (1ll << 31) - 256) return false;
    if ((long long)a.n * (a.H + 2) * (a.H + 2) >= (1ll << 32)) return false;
    memset(p, 0, sizeof *p);
    p->n_img = a.n; p->H = a.H; p->W = a.W; p->npl = a.cs_in / 128; p->nslices = a.cs_out / 256;
    p->period = a.H + 1;
    p->period_magic = (unsigned)(((1ull << 32) + (unsigned)p->period - 1) / (unsigned)p->period);
    p->canvas_rows = a.n * p->period;
    p->rP = a.W + 1;
    p->rP_magic = (unsigned)(((1ull << 32) + (unsigned)p->rP - 1) / (unsigned)p->rP);
    p->raster_rows = (3 * p->rP + 127) / p->rP + 1;
    p->num_tiles = (int)(((long long)p->canvas_rows * p->rP + 127) / 128);
    p->num_pairs = (p->num_tiles + 1) / 2;
    if ((long long)p->num_pairs * p->nslices >= (1ll << 28)) return false;
    p->num_units = p->num_pairs * p->nslices;
    p->slices_magic = (unsigned)(((1ull << 32) + (unsigned)p->nslices - 1) / (unsigned)p->nslices);
    p->plane_bytes = ((uint32_t)((p->raster_rows + 1) * p->rP + 2) * 128u + 1023u) & ~1023u;
    const uint32_t budget = 227u * 1024u, tail = (uint32_t)a.cs_out * 4u + 1024u + 1024u;
    // at least two plane buffers and four weight slots; planes first (up to 4), the rest to the weight ring
    if (2 * p->plane_bytes + 4 * W3_CHUNK + tail > budget) return false;
    int pb = (int)((budget - tail - 6 * W3_CHUNK) / p->plane_bytes);
    if (pb < 2) pb = 2;
    if (pb > 4) pb = 4;
    p->p_bufs = pb;
    int slots = (int)((budget - tail - (uint32_t)pb * p->plane_bytes) / W3_CHUNK);
    p->b_slots = slots > W3_MAX_BSLOTS ? W3_MAX_BSLOTS : slots;
    if (p->b_slots < 3) return false;
    p->off_plane = (uint32_t)p->b_slots * W3_CHUNK;
    p->off_bias = p->off_plane + (uint32_t)p->p_bufs * p->plane_bytes;
    p->off_bar = (p->off_bias + (uint32_t)a.cs_out * 4u + 15u) & ~15u;
    p->cs_out = a.cs_out; p->q = a.q; p->wtap3 = a.wimg_tap3; p->bias_sh = a.bias_sh; p->out = a.out; p->ovf = a.ovf;
    return true;
}

bool conv3x3_ws3_supported(const ConvArgs &a, int sm_count)
{
    W3Params p;
    return ws3_plan(a, &p, sm_count);
}

typedef CUresult (*W3EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int EPI>
static cudaError_t launch_ws3(W3Params &p, const W3Maps &maps, cudaStream_t st, int sm_count)
{
    const uint32_t smem_bytes = p.off_bar + 8u * (3 * W3_MAX_PBUF + 3 * W3_MAX_BSLOTS + 6) + 1024u;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_ws3_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
    }
    int ctas = 2 * p.num_units < sm_count ? 2 * p.num_units : sm_count;
    ctas &= ~1;
    static const bool pdl_on = [] { const char *e = getenv("YOLO_B200_PDL"); return e ? atoi(e) != 0 : true; }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)ctas, 1, 1);
    cfg.blockDim = dim3(W3_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_on ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, conv3x3_ws3_kernel<EPI>, p, maps);
}

cudaError_t conv3x3_ws3(const ConvArgs &a, cudaStream_t st, int sm_count)
{
    if (a.n == 0) return cudaSuccess;
    W3Params p;
    if (!ws3_plan(a, &p, sm_count)) return cudaErrorInvalidValue;
    static W3EncodeTiledFn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        enc = (W3EncodeTiledFn)fn;
    }
    W3Maps maps;
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
    case EPI_F_RNE:      return launch_ws3<EPI_F_RNE>(p, maps, st, sm_count);
    case EPI_F_RNE_NOHI: return launch_ws3<EPI_F_RNE_NOHI>(p, maps, st, sm_count);
    case EPI_P:          return launch_ws3<EPI_P>(p, maps, st, sm_count);
    default:             return launch_ws3<EPI_GENERIC>(p, maps, st, sm_count);
    }
}

}  // namespace yb
