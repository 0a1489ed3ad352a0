// head.cu — YOLOv2 detection head on the device: decode + score + threshold, then per-frame greedy NMS.
//
// Replaces get_boxes/conf_sort/NMS (c_embedding/yolo_forward.c:1052-1147) and
// decode_boxes/postprocess/nms (models/slim_yolo_v2.py:111-210,330-358).
//
// Floating-point expressions are written with explicit round-to-nearest intrinsics so the compiler cannot contract
// them into FMAs: the reference evaluates each product/sum separately in fp32 (NumPy / torch CPU), and NMS decisions
// compare against the threshold bit-for-bit.
#include "kernels.h"
#include "ptx.cuh"
#include <math.h>
#include <cstdio>
#include <cstdlib>

namespace yb {

__device__ __forceinline__ float sigmoid_py(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
// sigmoid() as written in the C driver: 1/(exp(x)+1) = sigma(-x) (yolo_forward.c:965-968), double math
__device__ __forceinline__ float sigmoid_c(float x) { return (float)(1 / (exp((double)x) + 1)); }

// Python head of one anchor (slim_yolo_v2.py:330-358), split so that the NMS kernel can decode in place: the score first, the
// box only for anchors that pass the confidence threshold.  Same expressions as head_decode_kernel (explicit rn operations).
__device__ __forceinline__ float decode_py_score(const HeadArgs &a, const int8_t *p, int an, int &best)
{
    const int8_t *pc = p + a.A + an * a.C;
    const float inv = ldexpf(1.0f, -a.sa_pred);
    const float obj = sigmoid_py(p[an] * inv);
    float m = -INFINITY;
    for (int c = 0; c < a.C; ++c) m = fmaxf(m, pc[c] * inv);
    float sum = 0.f;
    for (int c = 0; c < a.C; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(pc[c] * inv, m)));
    best = 0;
    float score = -1.f;
    for (int c = 0; c < a.C; ++c) {
        const float s = __fmul_rn(__fdiv_rn(expf(__fsub_rn(pc[c] * inv, m)), sum), obj);
        if (s > score) { score = s; best = c; }
    }
    return score;
}
__device__ __forceinline__ float4 decode_py_box(const HeadArgs &a, const int8_t *p, int an, int row, int col)
{
    const int8_t *pb = p + a.A * (1 + a.C) + an * 4;
    const float inv = ldexpf(1.0f, -a.sa_pred);
    const float st = (float)a.stride;
    const float cx = __fmul_rn(__fadd_rn(sigmoid_py(pb[0] * inv), (float)col), st);
    const float cy = __fmul_rn(__fadd_rn(sigmoid_py(pb[1] * inv), (float)row), st);
    const float bw = __fmul_rn(__fmul_rn(expf(pb[2] * inv), a.anchors[an][0]), st);
    const float bh = __fmul_rn(__fmul_rn(expf(pb[3] * inv), a.anchors[an][1]), st);
    const float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
    const float iw = (float)a.in_w, ih = (float)a.in_h;
    float4 box;
    box.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(cx, hw), iw), 0.f), 1.f);
    box.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(cy, hh), ih), 0.f), 1.f);
    box.z = fminf(fmaxf(__fdiv_rn(__fadd_rn(cx, hw), iw), 0.f), 1.f);
    box.w = fminf(fmaxf(__fdiv_rn(__fadd_rn(cy, hh), ih), 0.f), 1.f);
    return box;
}

// One thread per (frame, cell, anchor).  Channel order of a cell (slim_yolo_v2.py:337-341, yolo_forward.c:1273):
// A conf | A*C class scores (anchor-major) | A*4 box terms (anchor-major).
__global__ void __launch_bounds__(256) head_decode_kernel(HeadArgs a)
{
    pdl_launch_dependents();
    pdl_wait();
    const int N = a.gh * a.gw * a.A;
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= a.n * N) return;
    int f = gid / N, idx = gid % N;
    int cell = idx / a.A, an = idx % a.A;
    int row = cell / a.gw, col = cell % a.gw;
    const int8_t *p = a.pred + ((size_t)f * a.gh * a.gw + cell) * a.cs;
    const int8_t *pc = p + a.A + an * a.C;
    const int8_t *pb = p + a.A * (1 + a.C) + an * 4;
    float score; int best; float4 box;
    if (a.head_mode == YOLO_B200_HEAD_PYTHON) {
        score = decode_py_score(a, p, an, best);
        box = decode_py_box(a, p, an, row, col);
    } else {
        // C head on its well-defined subset (see include/yolo_b200.h, YOLO_B200_HEAD_C)
        const double sc = exp2((double)a.sa_pred);
        float conf = sigmoid_c((float)(p[an] / sc));
        float c0 = (float)exp((double)(float)(pc[0] / sc));
        float c1 = (float)exp((double)(float)(pc[1] / sc));
        float sum = __fadd_rn(__fadd_rn(0.f, c0), c1);
        c0 = __fdiv_rn(c0, sum); c1 = __fdiv_rn(c1, sum);
        best = c0 >= c1 ? 0 : 1;
        score = __fmul_rn(conf, best ? c1 : c0);
        float tx = (float)(pb[0] / sc), ty = (float)(pb[1] / sc), tw = (float)(pb[2] / sc), th = (float)(pb[3] / sc);
        float st = (float)a.stride;
        float xc = __fmul_rn(__fadd_rn(sigmoid_c(tx), (float)col), st);
        float yc = __fmul_rn(__fadd_rn(sigmoid_c(ty), (float)row), st);
        float bw = (float)((double)a.anchors[an][0] * exp((double)tw) * (double)a.stride);
        float bh = (float)((double)a.anchors[an][0] * exp((double)th) * (double)a.stride);   // anchor WIDTH, :1044
        float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
        box.x = (float)(int)__fsub_rn(xc, hw); box.z = (float)(int)__fadd_rn(xc, hw);
        box.y = (float)(int)__fsub_rn(yc, hh); box.w = (float)(int)__fadd_rn(yc, hh);
    }
    a.scores[gid] = score; a.cls[gid] = best; a.boxes[gid] = box;
}

cudaError_t head_decode(const HeadArgs &a, cudaStream_t st)
{
    int total = a.n * a.gh * a.gw * a.A;
    if (total == 0) return cudaSuccess;
    { cudaError_t le = launch_pdl(head_decode_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a); if (le != cudaSuccess) return le; }
    return cudaGetLastError();
}

// ---- NMS ------------------------------------------------------------------------------------------
//
// One CTA of 512 threads per frame, two CTAs per SM (so that one frame's serial phases overlap the other's pair loops).  Candidates over the threshold are compacted and sorted once by
// (class, score desc, tie-break) with a shared-memory bitonic sort.  Greedy suppression then runs per class segment
// (python head; the C head is class-agnostic = one segment) in chunks of 128 sorted candidates, against a COMPACTED
// list of the candidates kept so far (which is also the output list); four barriers + one per chunk:
//   a. every candidate of the chunk is tested against the kept list; 4 threads share one candidate, each taking a part
//      of the list / of the index cells; then the chunk is squeezed to the candidates that are still alive;
//   b. every candidate records, as bit-rows, which EARLIER candidates of the chunk would suppress it; the triangle of
//      pairs is cut into (candidate word, scanned half-word) items dealt round-robin to the 16 warps;
//   c. one warp resolves the chunk word by word (ballots; rank order only inside a 32-candidate word and only for the
//      candidates that have a possibly-surviving predecessor there) and leaves per-word survivor counts;
//   d. the chunk's survivors are appended to the kept list (in place: kept count <= processed count; no scan).
// Python head with thresh > 1e-6: the kept list also carries a SPATIAL INDEX so that step a visits only boxes that can
// matter.  IoU > t implies (i) area ratio within [t, 1/t] and (ii) |dcx| < (1-t)/(1+t) max(w_i,w_j), likewise in y
// (from inter > t/(1+t) (a_i + a_j): the x-overlap must exceed t/(1+t) times the sum of the widths).  Kept boxes are therefore filed, as linked lists, under
// (area bucket by binary exponent, 8x8 centre cell); a candidate walks only the lists of compatible buckets and of the
// cells its window covers (with slack for rounding; every visited pair still goes through the screen and the exact test,
// so pruning can only skip pairs that could not suppress).  Candidates with a degenerate area scan the whole list: the
// reference's 0/0 = NaN quirk makes zero-area boxes suppress each other at any distance.
// The decisions are those of sequential greedy NMS: a candidate is dropped iff a KEPT higher-ranked candidate of its
// segment overlaps it beyond the threshold.  Most pairs are rejected without a division: disjoint boxes, boxes whose
// area ratio already bounds the IoU below the threshold, and quotients clearly away from the threshold; only
// borderline pairs evaluate the reference's exact fp32 expression.

constexpr int NMS_THREADS = 512;
constexpr int NMS_CHUNK = 128;
constexpr int NMS_SLICES = NMS_THREADS / NMS_CHUNK;
constexpr int NB = NMS_CHUNK / 32;                           // candidate words per chunk
constexpr int NMS_MAX_CLASSES = 64;
constexpr int NMS_BUCKETS = 12;          // area buckets by binary exponent: bucket b = areas in [2^(b-11), 2^(b-10)), bucket 0 also everything smaller
constexpr int NMS_GRID = 8;              // centre cells per axis over [0,1]
constexpr unsigned NMS_END = 0xffffu;

struct NmsSmem {
    union {
        unsigned long long key[HEAD_MAX_CAND];                     // sort phase
        struct {                                                   // suppression phase (keys no longer needed)
            unsigned short idx[HEAD_MAX_CAND];                     // anchor index of sorted position / kept entry
            unsigned mask[NMS_CHUNK][NMS_CHUNK / 32];              // intra-chunk suppression rows
            float area[HEAD_MAX_CAND];                             // box areas (same order as box[])
        } s2;
    } u;
    float4 box[HEAD_MAX_CAND];                                     // sorted candidates; kept list compacted in place
    unsigned char cls[HEAD_MAX_CAND];
    unsigned keepmap[HEAD_MAX_CAND / 32];                          // by anchor index (python mode output order)
    unsigned chunk_dead[2][NMS_CHUNK / 32];                       // per chunk: dropped / no candidate (double-buffered)
    int warp_sums[NMS_THREADS / 32 + 1];
    // spatial index of the kept list (python head): linked lists per (area bucket, centre cell)
    unsigned ghead[NMS_BUCKETS * NMS_GRID * NMS_GRID];             // first kept entry of the list, NMS_END = empty
    unsigned short gnext[HEAD_MAX_CAND];                           // next entry of the same list
    float gw[NMS_BUCKETS], gh[NMS_BUCKETS];                        // widest / tallest kept box of the bucket (-1 = empty)
    unsigned long long wtot[NMS_THREADS / 32];                     // C head, conf_sort emulation: per-warp prefix maxima
    unsigned tie_min;                                              // C head: lowest score (bits) that occurs twice
};

// block-wide exclusive scan of a 0/1 flag; returns this thread's offset, *total = sum over the block
__device__ __forceinline__ int block_scan_flag(bool flag, int *warp_sums, int *total)
{
    unsigned b = __ballot_sync(0xffffffffu, flag);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int off = __popc(b & ((1u << lane) - 1));
    if (lane == 0) warp_sums[wid] = __popc(b);
    __syncthreads();
    int v = lane < NMS_THREADS / 32 ? warp_sums[lane] : 0;       // NMS_THREADS/32 <= 32 warps: one per lane
    int tot = __reduce_add_sync(0xffffffffu, v);
    int base = __reduce_add_sync(0xffffffffu, lane < wid ? v : 0);
    __syncthreads();
    *total = tot;
    return base + off;
}

__device__ __forceinline__ float area_py(float4 a) { return __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)); }

// python: overlap as in slim_yolo_v2.py:159-169, suppress when NOT (ovr <= thresh): the reference's fp32 expression,
// every operation rounded separately.  Only pairs the cheap screen below cannot decide get here.
__device__ __noinline__ bool suppress_py_exact(float4 a, float4 b, float thresh)
{
    float w = fmaxf(1e-28f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    float h = fmaxf(1e-28f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    float inter = __fmul_rn(w, h);
    float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_py(a), area_py(b)), inter));
    return !(ovr <= thresh);
}

// Screen: inter/(ai+aj-inter) > t  <=>  inter > t/(1+t) * (ai+aj).  The pair loops store ca = t/(1+t)*(1-1e-5) * area per box
// and evaluate, branch-free, max(iw,0)*ih > ca_i + ca_j: false means "certainly kept" (the 1e-5 slack covers every
// rounding difference, and the reference's 1e-28 clamps only matter for degenerate boxes, which the screen passes on as
// they make the right-hand side ~0); true sends the pair to suppress_py_exact.
__device__ __forceinline__ bool screen_py(float4 a, float ca, float4 b, float cb)
{
    const float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    const float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    return fmaxf(iw, 0.f) * ih >= ca + cb;
}

// C: integer boxes, overlap()/box_intersection()/box_union() of yolo_forward.c:1000-1036, suppress iou >= thresh
__device__ __forceinline__ bool suppress_c(float4 a, float4 b, float thresh)
{
    int ax1 = (int)a.x, ay1 = (int)a.y, ax2 = (int)a.z, ay2 = (int)a.w;
    int bx1 = (int)b.x, by1 = (int)b.y, bx2 = (int)b.z, by2 = (int)b.w;
    int ow = (ax2 - ax1 + bx2 - bx1) - (max(ax2, bx2) - min(ax1, bx1));
    int oh = (ay2 - ay1 + by2 - by1) - (max(ay2, by2) - min(ay1, by1));
    if ((ow <= 0 || oh <= 0) && thresh > 0.f) return false;       // iou = 0/uni = 0 (or NaN): never >= a positive threshold
    int inter = (ow <= 0 || oh <= 0) ? 0 : ow * oh;
    int uni = (ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter;
    float iou = __fdiv_rn((float)inter, (float)uni);
    return iou >= thresh;
}

// One pair: does box a (kept / higher ranked) suppress box b?  FAST: screen first (python head, thresh > 1e-6).
template <bool PY, bool FAST>
__device__ __forceinline__ bool suppresses(float4 a, float ca, float4 b, float cb, float thresh)
{
    if (!PY) return suppress_c(a, b, thresh);
    if (FAST && !screen_py(a, ca, b, cb)) return false;
    return suppress_py_exact(a, b, thresh);
}

// Four boxes of the sorted/kept array against box bj: bit k of the result = "entry i4+k may matter" (screen, or simply
// valid when there is no screen).  vm masks the entries that belong to the range.  Branch-free: the four screens are
// independent instruction streams; the caller branches once per block.
template <bool PY, bool FAST>
__device__ __forceinline__ unsigned screen4(const float4 *box, const float *carea, int i4, unsigned vm, float4 bj, float cj)
{
    if (!(PY && FAST)) return vm;
    const float4 c4 = *reinterpret_cast<const float4 *>(carea + i4);
    const float4 b0 = box[i4], b1 = box[i4 + 1], b2 = box[i4 + 2], b3 = box[i4 + 3];
    unsigned m = (unsigned)screen_py(b0, c4.x, bj, cj) | ((unsigned)screen_py(b1, c4.y, bj, cj) << 1) |
                 ((unsigned)screen_py(b2, c4.z, bj, cj) << 2) | ((unsigned)screen_py(b3, c4.w, bj, cj) << 3);
    return m & vm;
}

// Area bucket: floor(log_{1/t'}(area)) counted down from area = 1, t' = the slackened threshold.  Two boxes whose area
// ratio is at least t' are then at most ONE bucket apart (bscale = 1 / log2(1/t'); the error of log2f is orders of
// magnitude below the slack built into t').  Bucket 0 also takes everything smaller, the last one areas >= 1/t'.
__device__ __forceinline__ int nms_bucket(float area, float bscale)
{
    const float l = log2f(fmaxf(area, 1e-30f)) * bscale;                  // <= 0 for areas <= 1
    return min(max((int)floorf(l) + (NMS_BUCKETS - 1), 0), NMS_BUCKETS - 1);
}
__device__ __forceinline__ int nms_cell(float c) { return min(max((int)(c * (float)NMS_GRID), 0), NMS_GRID - 1); }


// conf_sort (yolo_forward.c:1114-1126) is a selection sort BY SWAPPING with a strict '>': pass i walks j = i+1.. and swaps
// src[j] into slot i whenever it beats the current occupant, so slot i ends with the first maximum of the tail and every
// strict prefix maximum ("record") of the tail moves to the slot of the NEXT record.  The result is sorted by score, but
// the order among EQUAL scores depends on that swap history, and NMS is order dependent.  This reproduces it exactly:
// elements below the lowest repeated score are all distinct, never take part in a swap chain that moves a larger element
// (a record chain only links elements of increasing score) and end in sorted order anyway, so the passes are replayed in
// parallel (one prefix-max scan + record shift per pass) on the sub-sequence S of candidates at or above that score, in
// their original (anchor) order.  On entry key[0..m) is sorted (score desc, anchor asc); on exit the first |S| keys are in
// conf_sort's order.  Nothing happens (one reduction) when all scores are distinct.
constexpr int NMS_PER = HEAD_MAX_CAND / NMS_THREADS;
__device__ void conf_sort_tie_order(NmsSmem &s, int m, const float *scores, int N, float conf_thresh)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s.tie_min = 0xffffffffu;
    __syncthreads();
    unsigned mymin = 0xffffffffu;
    for (int i = tid; i + 1 < m; i += NMS_THREADS) {
        const unsigned a = (unsigned)(s.u.key[i] >> 12), b = (unsigned)(s.u.key[i + 1] >> 12);     // score bits (class field is 0)
        if (a == b) mymin = min(mymin, a);
    }
    mymin = __reduce_min_sync(0xffffffffu, mymin);
    if (lane == 0 && mymin != 0xffffffffu) atomicMin(&s.tie_min, mymin);
    __syncthreads();
    const unsigned vmin = s.tie_min;
    if (vmin == 0xffffffffu) return;                               // block-uniform
    // S in anchor order: (score bits << 32 | anchor index); box[] is not in use yet
    unsigned long long *E = reinterpret_cast<unsigned long long *>(s.box);
    int T1 = 0;
    for (int base = 0; base < N; base += NMS_THREADS) {
        const int i = base + tid;
        const float sc = i < N ? scores[i] : -1.f;
        const bool in = i < N && sc > conf_thresh && __float_as_uint(sc) >= vmin;
        int tot;
        const int off = block_scan_flag(in, s.warp_sums, &tot);
        if (in) E[T1 + off] = ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned)i;
        T1 += tot;
    }
    __syncthreads();
    const int per = (T1 + NMS_THREADS - 1) / NMS_THREADS;          // <= NMS_PER; thread t owns slots [t per, (t + 1) per)
    const int p0 = tid * per;
    constexpr unsigned long long HI = 0xffffffff00000000ull;
    for (int i = 0; i + 1 < T1; ++i) {
        // scan key = score bits << 32 | ~slot: the maximum is the highest score at its FIRST slot
        unsigned long long loc = 0ull;                             // scores are > 0: every real key is > 0
#pragma unroll
        for (int q = 0; q < NMS_PER; ++q) {
            const int p = p0 + q;
            if (q < per && p >= i && p < T1) loc = max(loc, (E[p] & HI) | (unsigned long long)(0xffffffffu - (unsigned)p));
        }
        unsigned long long inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc = max(inc, o);
        }
        if (lane == 31) s.wtot[wid] = inc;
        unsigned long long run = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) run = 0ull;
        __syncthreads();
        unsigned long long top = 0ull;
#pragma unroll
        for (int w = 0; w < NMS_THREADS / 32; ++w) {
            const unsigned long long v = s.wtot[w];
            top = max(top, v);
            if (w < wid) run = max(run, v);
        }
        // own slots: a record (strictly above everything before it in the tail) takes the previous record's element;
        // slot i takes the tail's first maximum
        unsigned long long val[NMS_PER];
        unsigned rec = 0u;
#pragma unroll
        for (int q = 0; q < NMS_PER; ++q) {
            const int p = p0 + q;
            val[q] = 0ull;
            if (q < per && p >= i && p < T1) {
                const unsigned long long e = E[p];
                if (p == i) {
                    const unsigned g = 0xffffffffu - (unsigned)top;
                    if (g != (unsigned)i) { val[q] = E[g]; rec |= 1u << q; }
                    run = (e & HI) | (unsigned long long)(0xffffffffu - (unsigned)p);
                } else if ((unsigned)(e >> 32) > (unsigned)(run >> 32)) {
                    val[q] = E[0xffffffffu - (unsigned)run]; rec |= 1u << q;
                    run = (e & HI) | (unsigned long long)(0xffffffffu - (unsigned)p);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NMS_PER; ++q)
            if ((rec >> q) & 1u) E[p0 + q] = val[q];
        __syncthreads();
    }
    for (int k = tid; k < T1; k += NMS_THREADS) {
        const unsigned long long e = E[k];
        s.u.key[k] = ((e >> 32) << 12) | (unsigned long long)(HEAD_MAX_CAND - 1 - (unsigned)e);
    }
    __syncthreads();
}

#ifdef YB_NMS_TIMELINE
#define NMS_T(k) do { if (tid == 0) { long long c_ = clock64(); tacc[k] += c_ - tlast; tlast = c_; } } while (0)
#else
#define NMS_T(k) do { } while (0)
#endif

template <bool PY, bool FAST>
__global__ void __launch_bounds__(NMS_THREADS, 2) head_nms_kernel(HeadArgs a)
{
#ifdef YB_NMS_TIMELINE
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    int dbg_s4 = 0, dbg_ex = 0;
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem &s = *reinterpret_cast<NmsSmem *>(smem_raw);
    const int f = blockIdx.x;
    const int N = a.gh * a.gw * a.A;
    const float *scores = a.scores + (size_t)f * N;
    const int *cls = a.cls + (size_t)f * N;
    const float4 *boxes = a.boxes + (size_t)f * N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float thresh = a.nms_thresh;
    const float cfac = FAST ? thresh / (1.f + thresh) * 0.99999f : 1.f;     // screen factor (see screen_py)
    constexpr bool GRID = PY && FAST;
    // spatial-index bounds with slack: IoU > t needs area ratio >= t (bucket distance <= grid_d) and, per axis, centre
    // distance < (1 - t)/(1 + t) * max extent:  inter > t/(1+t) (a_i + a_j) and ih <= min(h) give
    // iw > t/(1+t) (w_i + w_j), and iw <= (w_i + w_j)/2 - |dcx|
    const float t_lo = fminf(thresh * 0.999f, 0.97f);                     // (buckets no finer than 3 % steps)
    const float bscale = GRID ? 1.f / log2f(1.f / t_lo) : 1.f;
    constexpr int grid_d = 1;
    const float inv_t = 1.f / t_lo * 1.0001f;                            // a suppressor is at most this much wider / taller
    const float grid_q = fmaxf(1.f - t_lo, 0.f) / (1.f + t_lo) * 1.0001f;

    // 1. threshold + compaction (python: score >= conf, slim_yolo_v2.py:190; C: score > conf, yolo_forward.c:1077).
    //    key = class | score bits | tie-break: python sorts per class, ties -> higher anchor index first (reversed stable
    //    ascending argsort); C is class-agnostic, ties -> the order conf_sort's swap history leaves (conf_sort_tie_order).
    int m = 0;
    for (int base = 0; base < N; base += NMS_THREADS) {
        int i = base + tid;
        float sc = i < N ? scores[i] : -1.f;
        bool cand = i < N && (PY ? sc >= a.conf_thresh : sc > a.conf_thresh);
        int tot;
        int off = block_scan_flag(cand, s.warp_sums, &tot);
        if (cand) {
            unsigned long long cf = PY ? (unsigned long long)(NMS_MAX_CLASSES - 1 - cls[i]) : 0ull;
            unsigned tb = PY ? (unsigned)i : (unsigned)(HEAD_MAX_CAND - 1 - i);
            s.u.key[m + off] = (cf << 44) | ((unsigned long long)__float_as_uint(sc) << 12) | tb;
        }
        m += tot;
    }
    for (int i = tid; i < HEAD_MAX_CAND / 32; i += NMS_THREADS) s.keepmap[i] = 0;
    int P = 1;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += NMS_THREADS) s.u.key[i] = 0ull;
    __syncthreads();

    NMS_T(0);
    // 2. bitonic sort, descending.  Two consecutive strides (j, j/2) of a stage are fused: a thread takes the four keys
    //    i, i+j/2, i+j, i+3j/2 through both compare-exchange levels in registers (half the barriers and shared-memory
    //    passes of the textbook loop); a stage with an odd number of strides ends with one plain pass.
    for (int k = 2; k <= P; k <<= 1) {
        int j = k >> 1;
        for (; j >= 2; j >>= 2) {
            const int h = j >> 1;
            for (int t = tid; t < P / 4; t += NMS_THREADS) {
                const int i = ((t & ~(h - 1)) << 2) | (t & (h - 1));             // bits h and j clear
                unsigned long long a0 = s.u.key[i], a1 = s.u.key[i + h], a2 = s.u.key[i + j], a3 = s.u.key[i + j + h];
                const bool desc = (i & k) == 0;                                 // k > j: the same for the four keys
                auto cx = [&](unsigned long long &x, unsigned long long &y) {
                    const bool sw = desc ? x < y : x > y;
                    const unsigned long long tx = sw ? y : x, ty = sw ? x : y;
                    x = tx; y = ty;
                };
                cx(a0, a2); cx(a1, a3);                                         // stride j
                cx(a0, a1); cx(a2, a3);                                         // stride j / 2
                s.u.key[i] = a0; s.u.key[i + h] = a1; s.u.key[i + j] = a2; s.u.key[i + j + h] = a3;
            }
            __syncthreads();
        }
        if (j == 1) {
            for (int t = tid; t < P / 2; t += NMS_THREADS) {
                const int i = t << 1;
                const unsigned long long x = s.u.key[i], y = s.u.key[i + 1];
                const bool desc = (i & k) == 0;
                if (desc ? x < y : x > y) { s.u.key[i] = y; s.u.key[i + 1] = x; }
            }
            __syncthreads();
        }
    }

    if (!PY) conf_sort_tie_order(s, m, scores, N, a.conf_thresh);
    NMS_T(1);
    // 3. sorted order -> anchor index, box, area, class (keys are dead after this; their storage is reused)
    constexpr int PER = HEAD_MAX_CAND / NMS_THREADS;
    unsigned short my_idx[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        int i = tid + r * NMS_THREADS;
        unsigned tb = i < m ? (unsigned)(s.u.key[i] & 0xfffu) : 0u;
        my_idx[r] = (unsigned short)(PY ? tb : (HEAD_MAX_CAND - 1 - tb));
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        int i = tid + r * NMS_THREADS;
        if (i < m) {
            int idx = my_idx[r];
            float4 b = boxes[idx];
            s.u.s2.idx[i] = (unsigned short)idx;
            s.box[i] = b;
            s.u.s2.area[i] = cfac * area_py(b);
            s.cls[i] = (unsigned char)cls[idx];
        }
    }
    __syncthreads();
    NMS_T(2);
    // 4. greedy suppression, segment by segment, chunk by chunk
    const int cj_local = tid & (NMS_CHUNK - 1);      // candidate within the chunk
    const int slice = tid / NMS_CHUNK;               // warp-uniform
    int out_count = 0;                               // C head: kept so far (= output position)
    const int nseg = PY ? a.C : 1;
    for (int sg = 0; sg < nseg; ++sg) {
        // python head: the sort put the classes in ascending order; segment of class sg = [first cls >= sg, first cls > sg)
        // (binary searches on the sorted class array; every thread computes the same bounds)
        int seg_b = 0, seg_e = m;
        if (PY) {
            int lo = 0, hi = m;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (s.cls[mid] < sg) lo = mid + 1; else hi = mid; }
            seg_b = lo; hi = m;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (s.cls[mid] <= sg) lo = mid + 1; else hi = mid; }
            seg_e = lo;
        }
        if (seg_b == seg_e) continue;

        if (GRID) {                                  // empty spatial index for this segment
            for (int i = tid; i < NMS_BUCKETS * NMS_GRID * NMS_GRID; i += NMS_THREADS) s.ghead[i] = NMS_END;
            if (tid < NMS_BUCKETS) { s.gw[tid] = -1.f; s.gh[tid] = -1.f; }
        }
        // per-chunk scratch is cleared one chunk ahead (dead words are double-buffered), so a chunk needs four barriers
        if (tid < 2 * NB) (&s.chunk_dead[0][0])[tid] = 0;
        for (int i = tid; i < NMS_CHUNK * NB; i += NMS_THREADS) (&s.u.s2.mask[0][0])[i] = 0;
        __syncthreads();
        int K = 0;                                   // kept in this segment: entries [seg_b, seg_b + K)
        int par = 0;                                 // which set of dead words this chunk uses
        // chunks start on 4-aligned positions (screen4 reads 4-aligned blocks); the first may begin before the segment
        for (int cs = seg_b & ~3; cs < seg_e; cs += NMS_CHUNK, par ^= 1) {
            NMS_T(9);
            unsigned *dead_w = s.chunk_dead[par];
            const int hi = min(cs + NMS_CHUNK, seg_e);             // end of the chunk's candidates
            const int j = cs + cj_local;
            const bool have = j >= seg_b && j < hi;
            // everything of candidate j that 4d needs is read here: the kept list grows into this chunk's storage
            const float4 bj = have ? s.box[j] : make_float4(0, 0, 0, 0);
            const float aj = have ? s.u.s2.area[j] : 0.f;
            const unsigned short my = have ? s.u.s2.idx[j] : 0;
            const unsigned char mc = have ? s.cls[j] : 0;
            if (tid < NB) s.chunk_dead[par ^ 1][tid] = 0;          // for the next chunk (last read before this chunk began)
            // 4a. against the kept list of this segment: slice 0 takes the first half, slice 1 the second, in 4-aligned
            //     blocks; every lane of a warp reads the same kept boxes (shared-memory broadcasts)
            bool dead = false;
            const float area_j = GRID ? area_py(bj) : 0.f;
            if (GRID && have && K > 0 && area_j >= 1e-20f) {
                // walk only the lists that can hold a suppressor of j; the slices take alternate cells
                const float wj = bj.z - bj.x, hj = bj.w - bj.y;
                const float cxj = 0.5f * (bj.x + bj.z), cyj = 0.5f * (bj.y + bj.w);
                const int bk = nms_bucket(area_j, bscale);
                const float wcap = wj * inv_t, hcap = hj * inv_t;
                int cnt = 0;
                for (int b = max(bk - grid_d, 0); b <= min(bk + grid_d, NMS_BUCKETS - 1) && !dead; ++b) {
                    const float wm = s.gw[b];
                    if (wm < 0.f) continue;                                      // nothing kept in this bucket yet
                    // centre window: the partner's extent is bounded by the bucket's maximum AND by extent_j / t
                    // (IoU > t needs both the width and the height ratio above t)
                    const float rx = grid_q * fmaxf(wj, fminf(wm, wcap)) + 1e-6f, ry = grid_q * fmaxf(hj, fminf(s.gh[b], hcap)) + 1e-6f;
                    const int x0 = nms_cell(cxj - rx), x1 = nms_cell(cxj + rx), y0 = nms_cell(cyj - ry), y1 = nms_cell(cyj + ry);
                    for (int yc = y0; yc <= y1 && !dead; ++yc) {
                        if (((cnt++) & (NMS_SLICES - 1)) != slice) continue;     // the slices take alternate cell rows
                        for (int xc = x0; xc <= x1 && !dead; ++xc) {
                            unsigned e = s.ghead[(b * NMS_GRID + yc) * NMS_GRID + xc];
                            while (e != NMS_END) {
                                const unsigned nx = s.gnext[e];                  // fetched alongside the box, not after the test
                                if (suppresses<PY, FAST>(s.box[e], s.u.s2.area[e], bj, aj, thresh)) { dead = true; break; }
                                e = nx;
                            }
                        }
                    }
                }
            } else if (have && K > 0) {
                const int lo = seg_b, khi = seg_b + K;
                const int base4 = lo & ~3;
                const int half = (((khi - base4 + NMS_SLICES - 1) / NMS_SLICES) + 3) & ~3;      // one part per slice
                const int beg = base4 + slice * half, end = min(khi, beg + half);
                for (int i4 = beg; i4 < end; i4 += 4) {
                    unsigned vm = 0xfu;                                         // which of the 4 entries belong to the list
                    if (i4 < lo) vm &= 0xfu << (lo - i4);
                    if (i4 + 4 > khi) vm &= 0xfu >> (i4 + 4 - khi);
                    unsigned mk = screen4<PY, FAST>(s.box, s.u.s2.area, i4, vm, bj, aj);
                    if (mk) {                                                    // rare: exact evaluation of the flagged pairs
                        while (mk) {
                            const int k = __ffs(mk) - 1;
                            mk &= mk - 1;
                            if (PY ? suppress_py_exact(s.box[i4 + k], bj, thresh) : suppress_c(s.box[i4 + k], bj, thresh)) dead = true;
                        }
                        if (dead) break;
                    }
                }
            }
            unsigned db = __ballot_sync(0xffffffffu, dead || !have);
            if (lane == 0 && db) atomicOr(&dead_w[cj_local >> 5], db);
            __syncthreads();                                                    // dead_w = dropped by the kept list (or no candidate)
            NMS_T(3);
            // 4a'. squeeze the candidates the kept list dropped out of the chunk (in place, rank order preserved; offsets
            // from the dead words, no scan): the pairwise work of 4b is quadratic in what is left.  From here on the chunk
            // is entries [cs, cs + n1); a writer may overwrite another candidate's slot, whose owner read it at the top.
            int n1 = 0;
            {
                int before = 0;
#pragma unroll
                for (int v = 0; v < NB; ++v) {
                    const int pc = __popc(~dead_w[v]);
                    if (v < (cj_local >> 5)) before += pc;
                    n1 += pc;
                }
                const unsigned dwj = dead_w[cj_local >> 5];
                if (tid < NMS_CHUNK && !((dwj >> lane) & 1u)) {
                    const int d = cs + before + __popc(~dwj & ((1u << lane) - 1u));
                    s.box[d] = bj; s.u.s2.area[d] = aj; s.u.s2.idx[d] = my; s.cls[d] = mc;
                }
            }
            __syncthreads();
            // the squeezed entry this thread appends in 4d if it survives (read before anything is appended)
            const bool have2 = tid < n1;
            const float4 b2 = have2 ? s.box[cs + tid] : make_float4(0, 0, 0, 0);
            const float a2 = have2 ? s.u.s2.area[cs + tid] : 0.f;
            const unsigned short my2 = have2 ? s.u.s2.idx[cs + tid] : 0;
            const unsigned char mc2 = have2 ? s.cls[cs + tid] : 0;
            const int nb1 = (n1 + 31) >> 5, hi1 = cs + n1;
            NMS_T(8);
            // 4b. predecessor rows: pred[c][w] = which candidates of word w of this chunk (ranked before c) would suppress c.
            // Work item = (candidate word wb, scanned word w <= wb, half h of w = four blocks of 4 entries):
            // nb1 (nb1 + 1) items, dealt round-robin to the 16 warps (the triangle of pairs is spread evenly whatever nb1
            // is).  The lanes of a warp are the 32 candidates of wb and read the same entries (shared-memory broadcasts);
            // hits are rare and go to the zero-initialised rows with one atomic OR per item.
            {
                const int nitems = nb1 * (nb1 + 1);
#pragma unroll 1
                for (int it = wid; it < nitems; it += NMS_THREADS / 32) {
                    const int wi = it >> 1, h = it & 1;
                    int wb = 0, wbase = 0;
                    while (wbase + wb + 1 <= wi) { wbase += wb + 1; ++wb; }     // wi -> (wb, w): wi = wb (wb + 1) / 2 + w
                    const int w = wi - wbase;
                    const int i0 = cs + 32 * w + 16 * h;
                    if (i0 >= hi1) continue;
                    const int c = wb * 32 + lane;                               // candidate inside the chunk
                    const bool act = c < n1;
                    // entries of this half that are ranked before c and inside the chunk
                    unsigned vm16 = w < wb ? 0xffffu : (((1u << lane) - 1u) >> (16 * h)) & 0xffffu;
                    if (i0 + 16 > hi1) vm16 &= 0xffffu >> (i0 + 16 - hi1);
                    if (!act) vm16 = 0u;
                    if (!__any_sync(0xffffffffu, vm16 != 0u)) continue;
                    const float4 bc = act ? s.box[cs + c] : make_float4(0, 0, 0, 0);
                    const float ac = act ? s.u.s2.area[cs + c] : 0.f;
                    unsigned bits = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i4 = i0 + 4 * q;
                        if (i4 < hi1) {                                          // warp-uniform
                            unsigned mk = screen4<PY, FAST>(s.box, s.u.s2.area, i4, (vm16 >> (4 * q)) & 0xfu, bc, ac);
#ifdef YB_NMS_TIMELINE
                            dbg_s4++; dbg_ex += __popc(__ballot_sync(0xffffffffu, mk != 0u)) ? 1 : 0;
#endif
                            while (mk) {
                                const int k = __ffs(mk) - 1;
                                mk &= mk - 1;
                                if (PY ? suppress_py_exact(s.box[i4 + k], bc, thresh) : suppress_c(s.box[i4 + k], bc, thresh)) bits |= 1u << (4 * q + k);
                            }
                        }
                    }
                    if (bits) atomicOr(&s.u.s2.mask[c][w], bits << (16 * h));
                }
            }
            NMS_T(4);
            __syncthreads();
            NMS_T(5);
            // 4c. one warp resolves the chunk word by word: lane l <-> candidate 32 w + l.  A candidate survives iff no
            // SURVIVING predecessor suppresses it; predecessors in earlier words are final, those in the same word are
            // resolved in rank order, but only for the candidates that have a possibly-surviving predecessor there.
            // It also leaves, per word, the number of survivors before it (4d needs no scan).
            if (wid == 0) {
                unsigned alive_w[NB];
                int run = 0;
#pragma unroll
                for (int w = 0; w < NB; ++w) {
                    alive_w[w] = 0u;
                    if (w < nb1) {                                               // warp-uniform
                        const int c = 32 * w + lane;
                        // rows of this candidate for all earlier words at once (independent loads), then one test
                        unsigned hit = 0u;
#pragma unroll
                        for (int v = 0; v < w; ++v) hit |= s.u.s2.mask[c & (NMS_CHUNK - 1)][v] & alive_w[v];
                        const bool pre = c < n1 && hit == 0u;
                        const unsigned inw = pre ? s.u.s2.mask[c][w] : 0u;
                        const unsigned cand = __ballot_sync(0xffffffffu, pre);
                        // in-word fixpoint: `und` = undecided candidates, `aw` = survivors so far.  An undecided candidate
                        // with a surviving predecessor is dropped; one whose predecessors are all decided survives.  The
                        // lowest undecided candidate is always decided, and typically most are after two or three rounds.
                        unsigned und = __ballot_sync(0xffffffffu, (inw & cand) != 0u);
                        unsigned aw = cand & ~und;
                        while (und) {
                            const bool mine = (und >> lane) & 1u;
                            const bool drop = mine && (inw & aw) != 0u;
                            const bool keep = mine && !drop && (inw & und) == 0u;
                            const unsigned d = __ballot_sync(0xffffffffu, drop), k = __ballot_sync(0xffffffffu, keep);
                            aw |= k; und &= ~(d | k);
                        }
                        alive_w[w] = aw;
                    }
                    if (lane == 0) { dead_w[w] = ~alive_w[w]; s.warp_sums[w] = run; }
                    run += __popc(alive_w[w]);
                }
                if (lane == 0) s.warp_sums[NB] = run;
            }
            __syncthreads();
            NMS_T(6);
            // 4d. append the survivors to the kept list (the squeezed entries have been in registers since before 4b)
            const unsigned dw = tid < NMS_CHUNK ? dead_w[cj_local >> 5] : 0xffffffffu;
            const bool alive = !((dw >> lane) & 1u);
            const int tot = s.warp_sums[NB];
            if (alive) {
                const int d = seg_b + K + s.warp_sums[cj_local >> 5] + __popc(~dw & ((1u << lane) - 1u));
                s.box[d] = b2; s.u.s2.area[d] = a2; s.u.s2.idx[d] = my2; s.cls[d] = mc2;
                if (PY) atomicOr(&s.keepmap[my2 >> 5], 1u << (my2 & 31));
                if (GRID) {                                                      // file it in the spatial index
                    const int b = nms_bucket(area_py(b2), bscale);
                    const int cell = (b * NMS_GRID + nms_cell(0.5f * (b2.y + b2.w))) * NMS_GRID + nms_cell(0.5f * (b2.x + b2.z));
                    s.gnext[d] = (unsigned short)atomicExch(&s.ghead[cell], (unsigned)d);
                    atomicMax(reinterpret_cast<int *>(&s.gw[b]), __float_as_int(b2.z - b2.x));      // widths are >= 0: integer order = float order
                    atomicMax(reinterpret_cast<int *>(&s.gh[b]), __float_as_int(b2.w - b2.y));
                }
            }
            // rows of the next chunk (this chunk's were last read in 4c)
            for (int i = tid; i < NMS_CHUNK * NB; i += NMS_THREADS) (&s.u.s2.mask[0][0])[i] = 0;
            K += tot;
            __syncthreads();
            NMS_T(7);
        }
        out_count = K;
    }

    // 5. output
#ifdef YB_NMS_TIMELINE
    if (tid == 0 && f < 2) printf("NMS frame %d m=%d, warp0: screen4 %d with-exact %d: compact %lld sort %lld gather %lld | 4a %lld squeeze %lld 4b %lld wait %lld 4c %lld 4d %lld other %lld\n", f, m, dbg_s4, dbg_ex, tacc[0], tacc[1], tacc[2], tacc[3], tacc[8], tacc[4], tacc[5], tacc[6], tacc[7], tacc[9]);
#endif
    yolo_b200_det *dets = a.dets + (size_t)f * a.max_det;
    if (PY) {
        // ascending anchor order (np.where(keep > 0), slim_yolo_v2.py:205)
        __syncthreads();
        int cnt = 0;
        for (int base = 0; base < N; base += NMS_THREADS) {
            int i = base + tid;
            bool k = i < N && ((s.keepmap[i >> 5] >> (i & 31)) & 1u);
            int tot;
            int off = block_scan_flag(k, s.warp_sums, &tot);
            if (k && cnt + off < a.max_det) {
                float4 b = boxes[i];
                yolo_b200_det d; d.x1 = b.x; d.y1 = b.y; d.x2 = b.z; d.y2 = b.w; d.score = scores[i]; d.cls = cls[i];
                d.anchor_index = i; d.pad_ = 0;
                dets[cnt + off] = d;
            }
            cnt += tot;
        }
        if (tid == 0) a.counts[f] = cnt;
    } else {
        // descending score order (conf_sort, yolo_forward.c:1114-1126): the kept list itself
        for (int i = tid; i < out_count && i < a.max_det; i += NMS_THREADS) {
            float4 b = s.box[i];
            int idx = s.u.s2.idx[i];
            yolo_b200_det d; d.x1 = b.x; d.y1 = b.y; d.x2 = b.z; d.y2 = b.w;
            d.score = scores[idx]; d.cls = s.cls[i]; d.anchor_index = idx; d.pad_ = 0;
            dets[i] = d;
        }
        if (tid == 0) a.counts[f] = out_count;
    }
}


// ---- grid NMS: the python head without a sort ----------------------------------------------------------------
//
// Greedy NMS keeps a candidate iff no KEPT candidate of its class that precedes it (higher score; equal score: higher anchor
// index) overlaps it beyond the threshold.  That fixed point depends only on the ORDER RELATION between overlapping pairs, so
// no global sort is needed, and the python head returns its detections in anchor order anyway (slim_yolo_v2.py:205):
//   1. every candidate (score >= conf) is filed by a counting sort under (area bucket, 16 x 16 centre cell): the same
//      conservative pruning as the sorted kernel above (IoU > t => area ratio within [t, 1/t] => at most one bucket apart;
//      centre distance per axis < (1-t)/(1+t) of the larger extent), but over ALL candidates at once and in CONTIGUOUS bins,
//      so a window row is one range of the bin-ordered arrays;
//   2. one thread per candidate, in bin order (the lanes of a warp are neighbours: similar ranges), walks those ranges as ONE
//      flat loop (a per-lane state machine over bucket / row / entry, so lanes with different windows never wait for each
//      other's inner loops) with the branch-free screen, and queues the pairs (candidate, possible predecessor) that pass;
//   3. the queue is drained with all lanes busy: the reference's exact fp32 IoU decides each pair; a confirmed pair is an
//      EDGE (predecessor -> candidate) and bumps the candidate's predecessor count;
//   4. edge-parallel propagation rounds: candidates without predecessors are kept; an edge whose source is kept drops its
//      target, an edge whose source is dropped releases it (count - 1; kept at zero); used edges are retired.  No pointer
//      chasing, ~8 rounds on the dense random-init frames;
//   5. kept anchors are flagged in a bitmap whose prefix counts give every kept candidate its slot in ascending anchor order.
// One CTA of 1024 threads per frame.  If the pairs do not fit the queue (adversarial inputs: thousands of mutually overlapping
// boxes) the frame falls back to pull rounds that re-walk the windows; degenerate (zero-area) candidates scan everything
// (the reference's 0/0 = NaN quirk).  The decisions are exactly those of the sequential algorithm.
constexpr int GN_THREADS = 1024;
constexpr int GN_G = 16;
constexpr int GN_CELLS = NMS_BUCKETS * GN_G * GN_G;
constexpr int GN_PER = HEAD_MAX_CAND / GN_THREADS;
constexpr int GN_RMAX = 2 * (GN_G >> 1);                    // ranges per candidate: two buckets x row groups (row pairs at least)
enum { GN_UNKNOWN = 0, GN_KEPT = 1, GN_DEAD = 2 };
constexpr unsigned GN_RETIRED = 0xffffffffu;

struct GnSmem {                       // byte offsets into dynamic shared memory (host: gn_layout)
    uint32_t box, score, npred, idx, cls, state, ofs, keepmap, misc, rng, pool, pool_cap, total;
};
__host__ __device__ inline GnSmem gn_layout(int N, uint32_t budget)
{
    const uint32_t n = (uint32_t)((N + 3) & ~3);
    GnSmem L;
    L.box = 0; L.score = L.box + 16u * n; L.npred = L.score + 4u * n; L.idx = L.npred + 4u * n;
    L.cls = L.idx + 2u * n; L.state = L.cls + n;
    L.ofs = (L.state + n + 15u) & ~15u;
    L.keepmap = (L.ofs + 2u * (GN_CELLS + 1) + 15u) & ~15u;
    L.misc = (L.keepmap + 4u * ((n + 31u) / 32u + 1u) + 15u) & ~15u;
    L.rng = L.misc + 512u;                                      // per-lane range lists of the walk: [warps][32][GN_RMAX + 1]
    L.pool = L.rng + 4u * (GN_THREADS / 32) * 32u * (GN_RMAX + 1);
    const uint32_t min_pool = 4u * GN_CELLS;                    // the bin counters live in the pool area while the bins are built
    uint32_t room = budget > L.pool ? budget - L.pool : 0u;
    if (room < min_pool) room = min_pool;
    L.pool_cap = room / 4u;
    L.total = L.pool + 4u * L.pool_cap;
    return L;
}

__device__ __forceinline__ int gn_cell(float c) { return min(max((int)(c * (float)GN_G), 0), GN_G - 1); }
// Bin order: (bucket, row GROUP, column, row inside the group).  A window's rows y0..y1 and columns x0..x1 are then one
// contiguous range of entries per row group instead of one per row: longer ranges, fewer range changes in the walk.
#ifndef GN_YSH
#define GN_YSH 1
#endif
constexpr int GN_YG = GN_G >> GN_YSH;                       // row groups
__device__ __forceinline__ int gn_bin(int b, int y, int x) { return (((b * GN_YG + (y >> GN_YSH)) * GN_G + x) << GN_YSH) + (y & ((1 << GN_YSH) - 1)); }
__device__ __forceinline__ int gn_group_first(int b, int yg, int x) { return ((b * GN_YG + yg) * GN_G + x) << GN_YSH; }

struct GnView {
    float4 *box; unsigned *score; unsigned *npred; unsigned short *idx; unsigned char *cls; volatile unsigned char *state;
    unsigned short *ofs; unsigned *pool; unsigned pool_cap; unsigned *pool_cnt; const float *gw, *gh;
    float thresh, cfac, bscale, inv_t, grid_q;
    int m;
#ifdef YB_NMS_TIMELINE
    unsigned *dbg;      // [4]: visits, queued pairs, -, ranges
#endif
};
#ifdef YB_NMS_TIMELINE
#define GN_COUNT(k) (dbgc[k]++)
#else
#define GN_COUNT(k) do { } while (0)
#endif

// Window of a candidate: which (bucket, row) ranges of the bin-ordered arrays can hold a box that overlaps it beyond the threshold
struct GnWindow { float wj, hj, cxj, cyj, wcap, hcap; int b, b_hi, x0, x1, yc, y1; };
__device__ __forceinline__ void gn_window_init(const GnView &v, float4 bj, float area_j, GnWindow &w)
{
    w.wj = bj.z - bj.x; w.hj = bj.w - bj.y;
    w.cxj = 0.5f * (bj.x + bj.z); w.cyj = 0.5f * (bj.y + bj.w);
    const int bk = nms_bucket(area_j, v.bscale);
    w.wcap = w.wj * v.inv_t; w.hcap = w.hj * v.inv_t;
    w.b = max(bk - 1, 0) - 1; w.b_hi = min(bk + 1, NMS_BUCKETS - 1);
    w.yc = 0; w.y1 = -1; w.x0 = 0; w.x1 = 0;
}
// next non-empty-bucket row range [p, end); false when the window is exhausted
__device__ __forceinline__ bool gn_window_next(const GnView &v, GnWindow &w, int &p, int &end)
{
    if (++w.yc > w.y1) {
        float wm;
        do { if (++w.b > w.b_hi) return false; wm = v.gw[w.b]; } while (wm < 0.f);          // skip buckets with nothing filed
        const float rx = v.grid_q * fmaxf(w.wj, fminf(wm, w.wcap)) + 1e-6f, ry = v.grid_q * fmaxf(w.hj, fminf(v.gh[w.b], w.hcap)) + 1e-6f;
        w.x0 = gn_cell(w.cxj - rx); w.x1 = gn_cell(w.cxj + rx); w.yc = gn_cell(w.cyj - ry) >> GN_YSH; w.y1 = gn_cell(w.cyj + ry) >> GN_YSH;
    }
    p = v.ofs[gn_group_first(w.b, w.yc, w.x0)]; end = v.ofs[gn_group_first(w.b, w.yc, w.x1 + 1)];     // (x1 + 1 = 16 is the next group's first bin)
    return true;
}

// Fallback only (queue overflow): the window of candidate j evaluated against the current states.
// Returns bit 0 = a KEPT predecessor exists, bit 1 = an UNDECIDED one.
__device__ __noinline__ unsigned gn_pull(const GnView &v, int j)
{
    const float4 bj = v.box[j];
    const unsigned sj = v.score[j], ij = v.idx[j], cj = v.cls[j];
    const float area_j = area_py(bj), caj = v.cfac * area_j;
    unsigned res = 0u;
    int p = 0, end = 0;
    GnWindow w;
    const bool degenerate = !(area_j >= 1e-20f);
    if (degenerate) end = v.m; else gn_window_init(v, bj, area_j, w);
    for (;;) {
        if (p >= end) { if (degenerate || !gn_window_next(v, w, p, end)) break; continue; }
        const float4 be = v.box[p];
        const unsigned se = v.score[p];
        if (p != j && v.cls[p] == cj && (se > sj || (se == sj && (unsigned)v.idx[p] > ij)) &&
            screen_py(be, v.cfac * area_py(be), bj, caj) && suppress_py_exact(be, bj, v.thresh)) {
            const unsigned st = v.state[p];
            res |= st == GN_KEPT ? 1u : st == GN_UNKNOWN ? 2u : 0u;
        }
        ++p;
    }
    return res;
}

__global__ void __launch_bounds__(GN_THREADS, 1) head_nms_grid_kernel(HeadArgs a, GnSmem L)
{
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.x;
    const int N = a.gh * a.gw * a.A;
    const float *scores = a.scores + (size_t)f * N;
    const int *cls = a.cls + (size_t)f * N;
    const float4 *boxes = a.boxes + (size_t)f * N;
    const int tid = threadIdx.x, lane = tid & 31;
    float4 *sbox = reinterpret_cast<float4 *>(smem_raw + L.box);
    unsigned *sscore = reinterpret_cast<unsigned *>(smem_raw + L.score);
    unsigned *npred = reinterpret_cast<unsigned *>(smem_raw + L.npred);
    unsigned short *sidx = reinterpret_cast<unsigned short *>(smem_raw + L.idx);
    unsigned char *scls = smem_raw + L.cls;
    volatile unsigned char *state = smem_raw + L.state;
    unsigned short *ofs = reinterpret_cast<unsigned short *>(smem_raw + L.ofs);
    unsigned *keepmap = reinterpret_cast<unsigned *>(smem_raw + L.keepmap);
    int *warp_sums = reinterpret_cast<int *>(smem_raw + L.misc);                    // [33]
    float *gw = reinterpret_cast<float *>(smem_raw + L.misc + 160), *gh = gw + NMS_BUCKETS;
    unsigned *pool_cnt = reinterpret_cast<unsigned *>(smem_raw + L.misc + 160 + 8 * NMS_BUCKETS);
    unsigned *pool = reinterpret_cast<unsigned *>(smem_raw + L.pool);
    unsigned *cnt = pool;                                                           // bin counters while the bins are built

    const float thresh = a.nms_thresh;
    const float t_lo = fminf(thresh * 0.999f, 0.97f);
    const float bscale = 1.f / log2f(1.f / t_lo);
#ifdef YB_NMS_TIMELINE
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    int rounds = 0;
    unsigned dbgc[4] = {0, 0, 0, 0};
#endif

    // 1. candidates -> (bucket, cell), counts, per-bucket extents
    for (int i = tid; i < GN_CELLS; i += GN_THREADS) cnt[i] = 0u;
    const int kwords = (N + 31) / 32;
    for (int i = tid; i <= kwords; i += GN_THREADS) keepmap[i] = 0u;
    if (tid < NMS_BUCKETS) { gw[tid] = -1.f; gh[tid] = -1.f; }
    if (tid == 0) { pool_cnt[0] = 0u; pool_cnt[1] = 0u; }
    __syncthreads();
    float4 mybox[GN_PER]; float mysc[GN_PER]; int mybin[GN_PER]; unsigned myrank[GN_PER];
    unsigned mycls = 0u;                                                            // fused decode: class of candidate r in byte r
    static_assert(GN_PER <= 4, "one class byte per candidate");
#pragma unroll
    for (int r = 0; r < GN_PER; ++r) {
        const int i = tid + r * GN_THREADS;
        mybin[r] = -1; myrank[r] = 0u; mysc[r] = 0.f; mybox[r] = make_float4(0, 0, 0, 0);
        if (i < N) {
            // fused decode (a.fused_decode): the prediction map is decoded here, the score first and the box only for anchors
            // that pass the threshold; otherwise head_decode_kernel's scratch is read
            const int cell = i / a.A, an = i - cell * a.A;
            const int8_t *pp = a.pred + ((size_t)f * a.gh * a.gw + cell) * a.cs;
            int best = 0;
            const float sc = a.fused_decode ? decode_py_score(a, pp, an, best) : scores[i];
            if (sc >= a.conf_thresh) {
                const int row = cell / a.gw;
                const float4 b = a.fused_decode ? decode_py_box(a, pp, an, row, cell - row * a.gw) : boxes[i];
                mycls |= (unsigned)best << (8 * r);
                const int bk = nms_bucket(area_py(b), bscale);
                const int cell = gn_bin(bk, gn_cell(0.5f * (b.y + b.w)), gn_cell(0.5f * (b.x + b.z)));
                mybox[r] = b; mysc[r] = sc; mybin[r] = cell;
                myrank[r] = atomicAdd(&cnt[cell], 1u);
                // extents are >= 0: integer order = float order; the plain read only skips atomics that cannot raise the maximum
                const float bw = fmaxf(b.z - b.x, 0.f), bh = fmaxf(b.w - b.y, 0.f);
                if (bw > *reinterpret_cast<volatile float *>(&gw[bk])) atomicMax(reinterpret_cast<int *>(&gw[bk]), __float_as_int(bw));
                if (bh > *reinterpret_cast<volatile float *>(&gh[bk])) atomicMax(reinterpret_cast<int *>(&gh[bk]), __float_as_int(bh));
            }
        }
    }
    __syncthreads();
    NMS_T(0);
    // 2. exclusive scan of the bin counts (3 bins per thread) -> bin offsets
    int m;
    {
        constexpr int BPT = GN_CELLS / GN_THREADS;
        static_assert(GN_CELLS % GN_THREADS == 0, "bins per thread");
        unsigned c[BPT], sum = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) { c[k] = cnt[tid * BPT + k]; sum += c[k]; }
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) warp_sums[tid >> 5] = (int)inc;
        __syncthreads();
        const int v = warp_sums[lane];
        const unsigned wbase = (unsigned)__reduce_add_sync(0xffffffffu, lane < (tid >> 5) ? v : 0);
        m = __reduce_add_sync(0xffffffffu, v);
        unsigned run = wbase + inc - sum;
#pragma unroll
        for (int k = 0; k < BPT; ++k) { ofs[tid * BPT + k] = (unsigned short)run; run += c[k]; }
        if (tid == GN_THREADS - 1) ofs[GN_CELLS] = (unsigned short)run;
    }
    __syncthreads();
    NMS_T(1);
    // 3. scatter into bin order
#pragma unroll
    for (int r = 0; r < GN_PER; ++r) {
        if (mybin[r] >= 0) {
            const int i = tid + r * GN_THREADS;
            const int p = (int)ofs[mybin[r]] + (int)myrank[r];
            sbox[p] = mybox[r]; sscore[p] = __float_as_uint(mysc[r]); sidx[p] = (unsigned short)i; scls[p] = a.fused_decode ? (unsigned char)(mycls >> (8 * r)) : (unsigned char)cls[i];
            npred[p] = 0u;
        }
    }
    __syncthreads();            // the counters (pool area) are dead from here on
    NMS_T(2);

    GnView v;
    v.box = sbox; v.score = sscore; v.npred = npred; v.idx = sidx; v.cls = scls; v.state = state; v.ofs = ofs;
    v.pool = pool; v.pool_cap = L.pool_cap; v.pool_cnt = pool_cnt; v.gw = gw; v.gh = gh;
    v.thresh = thresh; v.cfac = thresh / (1.f + thresh) * 0.99999f; v.bscale = bscale;
    v.inv_t = 1.f / t_lo * 1.0001f; v.grid_q = fmaxf(1.f - t_lo, 0.f) / (1.f + t_lo) * 1.0001f; v.m = m;
    // 4. candidate pairs.  Every unordered pair is looked at ONCE, from its member with the lower bin-ordered position: a
    //    candidate scans only entries BEHIND it (the rest of its own row group, the later row groups of its window in its own
    //    bucket, and its window in the next bucket) and the queued pair is oriented by the order relation.  (A window always
    //    contains every box that overlaps its owner beyond the threshold, so the lower member finds the higher one.)
    //    Degenerate candidates scan everything; pairs with exactly one degenerate member are recorded by that member.
    //    Warps take chunks of 32 bin-ordered candidates from a counter.  Converged prologue: every lane loads its candidate
    //    and writes the list of its non-empty entry ranges (begin | end << 16) to its own strip of shared memory; the walk is
    //    then ONE flat loop per lane (visit an entry; when the range is used up, fetch the next one: a handful of predicated
    //    instructions), so lanes with different windows never wait for each other's inner loops.
    {
        unsigned *chunk_next = pool_cnt + 1;
        unsigned *myrng = reinterpret_cast<unsigned *>(smem_raw + L.rng) + ((tid >> 5) * 32 + lane) * (GN_RMAX + 1);
        for (;;) {
            int c = 0;
            if (lane == 0) c = (int)atomicAdd(chunk_next, 1u);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (32 * c >= m) break;
            c = (m + 31) / 32 - 1 - c;                           // last chunks first: the large-box buckets at the end of the bin order walk the widest
                                                                 // windows, so they should not be the ones left for the tail (0.150 -> 0.1475 ms per 256 dense frames)
            const int j = 32 * c + lane;
            const bool valid = j < m;
            const float4 bj = valid ? sbox[j] : make_float4(0, 0, 0, 0);
            const unsigned sj = valid ? sscore[j] : 0u, ij = valid ? sidx[j] : 0u, cj = valid ? scls[j] : 0u;
            const float area_j = area_py(bj), caj = v.cfac * area_j;
            const bool degenerate = !(area_j >= 1e-20f);
            int nr = 0;
            if (valid && degenerate) myrng[nr++] = (unsigned)m << 16;               // degenerate: everything, once
            else if (valid) {
                const int bk = nms_bucket(area_j, bscale);
                const float wj = bj.z - bj.x, hj = bj.w - bj.y;
                const float cxj = 0.5f * (bj.x + bj.z), cyj = 0.5f * (bj.y + bj.w);
                const float wcap = wj * v.inv_t, hcap = hj * v.inv_t;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int b = bk + k;
                    if (b >= NMS_BUCKETS) continue;
                    const float wm = gw[b];
                    if (wm < 0.f) continue;                                         // nothing filed in this bucket
                    const float rx = v.grid_q * fmaxf(wj, fminf(wm, wcap)) + 1e-6f, ry = v.grid_q * fmaxf(hj, fminf(gh[b], hcap)) + 1e-6f;
                    const int x0 = gn_cell(cxj - rx), x1 = gn_cell(cxj + rx);
                    const int g0 = (k == 0 ? gn_cell(cyj) : gn_cell(cyj - ry)) >> GN_YSH;   // own bucket: from the candidate's own row group on
                    const int g1 = gn_cell(cyj + ry) >> GN_YSH;
                    for (int g = g0; g <= g1; ++g) {
                        int beg = ofs[gn_group_first(b, g, x0)];
                        const int end = ofs[gn_group_first(b, g, x1 + 1)];
                        if (k == 0 && g == g0) beg = j + 1;                         // behind the candidate itself
                        if (beg < end) myrng[nr++] = (unsigned)beg | ((unsigned)end << 16);
                        GN_COUNT(3);
                    }
                }
            }
            __syncwarp();
            int r = 0, p = 0, end = 0;
            if (nr > 0) { const unsigned w0 = myrng[0]; p = w0 & 0xffffu; end = w0 >> 16; }
            while (r < nr) {
                const float4 be = sbox[p];
                const float area_e = area_py(be);
                GN_COUNT(0);
                if (screen_py(be, v.cfac * area_e, bj, caj)) {
                    // not itself; same class; a pair with ONE degenerate member belongs to that member's scan
                    const bool mine = degenerate ? (p > j || area_e >= 1e-20f) : area_e >= 1e-20f;
                    if (p != j && mine && scls[p] == cj) {
                        GN_COUNT(1);
                        const unsigned se = sscore[p];
                        const bool p_first = se > sj || (se == sj && (unsigned)sidx[p] > ij);      // p precedes j?
                        const unsigned slot = atomicAdd(pool_cnt, 1u);
                        if (slot < L.pool_cap) pool[slot] = p_first ? ((unsigned)j | ((unsigned)p << 16)) : ((unsigned)p | ((unsigned)j << 16));
                    }
                }
                if (++p >= end) {
                    ++r;
                    const unsigned wn = myrng[r];                                   // (one slot past the last range exists)
                    p = wn & 0xffffu; end = wn >> 16;
                }
            }
            __syncwarp();
        }
    }
    NMS_T(3);
    __syncthreads();
    NMS_T(4);
    const unsigned npairs = *pool_cnt;
    if (npairs <= L.pool_cap) {
        // 5. exact test of the queued pairs (all lanes busy).  Every warp owns one contiguous segment of the queue and
        //    squeezes it in place (ballot offsets) to the pairs that are still needed: first to the confirmed pairs = EDGES
        //    (target | source << 16), then, pass by pass below, to the edges that are still undecided, so a pass costs what
        //    is left, not what was queued.
        const int wid = tid >> 5;
        const unsigned seg_beg = (unsigned)(((unsigned long long)npairs * (unsigned)wid) / (GN_THREADS / 32));
        unsigned live_end = (unsigned)(((unsigned long long)npairs * (unsigned)(wid + 1)) / (GN_THREADS / 32));
        {
            unsigned wr = seg_beg;
            for (unsigned e0 = seg_beg; e0 < live_end; e0 += 32) {
                const unsigned e = e0 + lane;
                bool keep = false;
                unsigned wd = 0u;
                if (e < live_end) {
                    wd = pool[e];
                    const unsigned j = wd & 0xffffu, p = wd >> 16;
                    keep = suppress_py_exact(sbox[p], sbox[j], thresh);
                    if (keep) atomicAdd(&npred[j], 1u);
                }
                const unsigned kb = __ballot_sync(0xffffffffu, keep);
                if (keep) pool[wr + __popc(kb & ((1u << lane) - 1u))] = wd;
                wr += __popc(kb);
            }
            live_end = wr;
        }
        __syncthreads();
        for (int j = tid; j < m; j += GN_THREADS) state[j] = npred[j] ? GN_UNKNOWN : GN_KEPT;
        __syncthreads();
        NMS_T(7);
        // 6. edge-parallel propagation rounds.  States are read while other threads update them: a value is either
        //    "undecided" or final, so acting on it is always right; each round decides at least the first undecided candidate.
        //    An edge whose source is kept drops its target; one whose source is dropped releases it (count - 1, kept at zero);
        //    used edges leave the segment.  Several passes per barrier: other warps' decisions are visible without one.
        int again;
        do {
            int progress = 0;
#pragma unroll 1
            for (int pass = 0; pass < 3 && live_end > seg_beg; ++pass) {
                unsigned wr = seg_beg;
                for (unsigned e0 = seg_beg; e0 < live_end; e0 += 32) {
                    const unsigned e = e0 + lane;
                    bool keep = false;
                    unsigned wd = 0u;
                    if (e < live_end) {
                        wd = pool[e];
                        const unsigned j = wd & 0xffffu;
                        const unsigned sp = state[wd >> 16], st = state[j];
                        if (st == GN_UNKNOWN) {
                            if (sp == GN_KEPT) { state[j] = GN_DEAD; progress = 1; }
                            else if (sp == GN_DEAD) { if (atomicSub(&npred[j], 1u) == 1u) state[j] = GN_KEPT; progress = 1; }
                            else keep = true;
                        }
                    }
                    const unsigned kb = __ballot_sync(0xffffffffu, keep);
                    if (keep) pool[wr + __popc(kb & ((1u << lane) - 1u))] = wd;
                    wr += __popc(kb);
                }
                live_end = wr;
            }
            again = __syncthreads_or(progress);
#ifdef YB_NMS_TIMELINE
            ++rounds;
#endif
        } while (again);
    } else {
        // the pairs did not fit: pull rounds, every undecided candidate re-walks its window
        for (int j = tid; j < m; j += GN_THREADS) state[j] = GN_UNKNOWN;
        __syncthreads();
        int again;
        do {
            int unk = 0;
            for (int j = tid; j < m; j += GN_THREADS) {
                if (state[j] != GN_UNKNOWN) continue;
                const unsigned res = gn_pull(v, j);
                if (res & 1u) state[j] = GN_DEAD; else if (!(res & 2u)) state[j] = GN_KEPT; else unk = 1;
            }
            again = __syncthreads_or(unk);
        } while (again);
    }
    NMS_T(5);
    // 7. kept anchors in ascending anchor order (np.where(keep > 0), slim_yolo_v2.py:205): bitmap by anchor index, exclusive
    //    prefix counts of its words (<= 128 words: four per lane of one warp), every kept candidate then knows its slot
    for (int j = tid; j < m; j += GN_THREADS)
        if (state[j] == GN_KEPT) { const unsigned i = sidx[j]; atomicOr(&keepmap[i >> 5], 1u << (i & 31)); }
    __syncthreads();
    unsigned *kprefix = reinterpret_cast<unsigned *>(smem_raw + L.ofs);             // the bin offsets are dead: [kwords + 1] prefix counts
    if (tid < 32) {
        constexpr int WPL = HEAD_MAX_CAND / 32 / 32;                                // 4 words per lane
        unsigned c[WPL], sum = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) { const int wi = lane * WPL + k; c[k] = wi < kwords ? __popc(keepmap[wi]) : 0u; sum += c[k]; }
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        unsigned run = inc - sum;
#pragma unroll
        for (int k = 0; k < WPL; ++k) { const int wi = lane * WPL + k; if (wi <= kwords) kprefix[wi] = run; run += c[k]; }
        if (lane == 31) { a.counts[f] = (int)inc; if (kwords == HEAD_MAX_CAND / 32) kprefix[kwords] = inc; }
    }
    __syncthreads();
    yolo_b200_det *dets = a.dets + (size_t)f * a.max_det;
    for (int j = tid; j < m; j += GN_THREADS) {
        if (state[j] != GN_KEPT) continue;
        const unsigned i = sidx[j];
        const unsigned slot = kprefix[i >> 5] + __popc(keepmap[i >> 5] & ((1u << (i & 31)) - 1u));
        if (slot < (unsigned)a.max_det) {
            uint4 *dst = reinterpret_cast<uint4 *>(dets + slot);                   // 32-byte records, 16-byte aligned: two vector stores
            const float4 b = sbox[j];
            dst[0] = make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
            dst[1] = make_uint4(sscore[j], (unsigned)scls[j], i, 0u);
        }
    }
    NMS_T(6);
#ifdef YB_NMS_TIMELINE
    for (int k = 0; k < 4; ++k) atomicAdd(&reinterpret_cast<unsigned *>(smem_raw + L.misc + 272)[k], dbgc[k]);
    __syncthreads();
    if (tid == 0 && f < 2) {
        const unsigned *dg = reinterpret_cast<unsigned *>(smem_raw + L.misc + 272);
        printf("gridNMS frame %d m=%d: visits %u ranges %u queued pairs %u rounds %d | cycles: bin %lld scan %lld scatter %lld | walk(thread0) %lld wait %lld | exact %lld rounds %lld | output %lld\n",
               f, m, dg[0], dg[3], npairs, rounds, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[7], tacc[5], tacc[6]);
    }
#endif
}

template <bool PY, bool FAST>
static cudaError_t nms_attrs()
{
    cudaError_t e = cudaFuncSetAttribute(head_nms_kernel<PY, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem));
    if (e != cudaSuccess) return e;
    // two frames per SM: ask for the full shared-memory carve-out
    return cudaFuncSetAttribute(head_nms_kernel<PY, FAST>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t head_init(void)
{
    cudaError_t e = nms_attrs<true, true>();
    if (e == cudaSuccess) e = nms_attrs<true, false>();
    if (e == cudaSuccess) e = nms_attrs<false, false>();
    if (e == cudaSuccess) e = cudaFuncSetAttribute(head_nms_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    return e;
}

// ---- detection packing (multi-GPU collection: only the filled part of the fixed-capacity lists travels) ----
// offsets[f] = sum of min(counts[.], max_det) over the frames before f (offsets[n] = total); packed[offsets[f] + i] = dets[f][i].
// One CTA: the prefix over <= a few thousand frames is one block scan per 1024 frames; the copy is 32 bytes per record.
__global__ void __launch_bounds__(1024) pack_offsets_kernel(const int32_t *counts, int n, int max_det, int32_t *offsets)
{
    __shared__ int warp_sums[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int f = base + tid;
        const int c = f < n ? min(max(counts[f], 0), max_det) : 0;
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        const int v = warp_sums[lane];
        const int wbase = __reduce_add_sync(0xffffffffu, lane < wid ? v : 0);
        const int tot = __reduce_add_sync(0xffffffffu, v);
        if (f < n) offsets[f] = carry + wbase + inc - c;
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = carry;
}
__global__ void __launch_bounds__(256) pack_copy_kernel(const yolo_b200_det *dets, const int32_t *offsets, int n, int max_det, yolo_b200_det *packed)
{
    // one CTA per frame slice: 16-byte halves of the 32-byte records, coalesced both ways
    const int f = blockIdx.x;
    const int beg = offsets[f], cnt = offsets[f + 1] - beg;
    const uint4 *src = reinterpret_cast<const uint4 *>(dets + (size_t)f * max_det);
    uint4 *dst = reinterpret_cast<uint4 *>(packed + beg);
    for (int i = threadIdx.x; i < 2 * cnt; i += blockDim.x) dst[i] = src[i];
}
cudaError_t pack_detections(const yolo_b200_det *dets, const int32_t *counts, int n, int max_det, yolo_b200_det *packed, int32_t *offsets, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    pack_offsets_kernel<<<1, 1024, 0, st>>>(counts, n, max_det, offsets);
    pack_copy_kernel<<<n, 256, 0, st>>>(dets, offsets, n, max_det, packed);
    return cudaGetLastError();
}

static bool nms_grid_selected(const HeadArgs &a)
{
    // python head: the sort-free grid kernel (YOLO_B200_NMS_SORTED=1 keeps the sorted, chunked kernel for comparison)
    static const bool sorted_nms = [] { const char *e = getenv("YOLO_B200_NMS_SORTED"); return e && atoi(e) != 0; }();
    return a.head_mode == YOLO_B200_HEAD_PYTHON && a.nms_thresh > 1e-6f && !sorted_nms;
}

bool head_nms_fuses_decode(const HeadArgs &a)
{
    static const bool fuse = [] { const char *e = getenv("YOLO_B200_FUSE_DECODE"); return e ? atoi(e) != 0 : true; }();
    return fuse && nms_grid_selected(a) && a.C <= 255;
}

cudaError_t head_nms(const HeadArgs &a, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    if (a.gh * a.gw * a.A > HEAD_MAX_CAND || a.C > NMS_MAX_CLASSES) return cudaErrorInvalidValue;
    if (nms_grid_selected(a)) {
        const GnSmem L = gn_layout(a.gh * a.gw * a.A, 226u * 1024u);
        { cudaError_t le = launch_pdl(head_nms_grid_kernel, dim3((unsigned)a.n), dim3(GN_THREADS), L.total, st, a, L); if (le != cudaSuccess) return le; }
        return cudaGetLastError();
    }
    if (a.head_mode != YOLO_B200_HEAD_PYTHON) head_nms_kernel<false, false><<<a.n, NMS_THREADS, sizeof(NmsSmem), st>>>(a);
    else if (a.nms_thresh > 1e-6f) head_nms_kernel<true, true><<<a.n, NMS_THREADS, sizeof(NmsSmem), st>>>(a);
    else head_nms_kernel<true, false><<<a.n, NMS_THREADS, sizeof(NmsSmem), st>>>(a);
    return cudaGetLastError();
}

}  // namespace yb
