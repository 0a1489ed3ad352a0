/*
 * oracle/ref_int8.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's fixed-point slim_yolo_v2 forward pass.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may
 * build, load or call this file.  The product (libyolo_b200.so) never links or calls it.
 *
 * Parity pinning status:
 *   - Contract P (PyTorch fake-quant) is PINNED: tests/golden/ holds feature maps and detections
 *     produced by the unmodified reference module models/slim_yolo_v2.py (oracle/gen_golden.py),
 *     and tests/test_oracle.py checks this file against them bit-for-bit.
 *   - The RGB444 front end and the C head helpers are PINNED against the reference's own
 *     functions compiled from c_embedding/yolo_forward.c (oracle/Makefile -> oracle/_ref/).
 *   - Contract F's shift rounding is "parity unpinned": the MAC/shift/saturate arithmetic of the
 *     C path lives in FPGA RTL that is not in the reference repository and the reference has no
 *     tests or golden vectors.  The shift programme (yolo_forward.c:233-257) is pinned; the
 *     rounding mode is a parameter (RNE default).
 *
 * All feature maps are int8 NHWC with an explicit channel stride.  Sums of products are exact in
 * int32; everything after the accumulation is done in int64 so nothing depends on wrap-around.
 *
 * Reference lines followed (relative to the reference checkout):
 *   layer list / flags ......... c_embedding/yolo_forward.c:1202-1262, models/slim_yolo_v2.py:58-87
 *   3x3, stride 1, zero pad 1 .. utils/modules.py:20-24
 *   leaky slope 1/8 ............ utils/modules.py:25
 *   shift programme ............ c_embedding/yolo_forward.c:233-257
 *   16-bit accumulator bound ... models/slim_yolo_v2.py:222-227
 *   activation fake-quant ...... models/slim_yolo_v2.py:16-38
 *   op order conv,leaky,q,pool . models/slim_yolo_v2.py:218-231
 *   head split / decode ........ models/slim_yolo_v2.py:111-143,330-350
 *   postprocess / nms .......... models/slim_yolo_v2.py:145-210
 *   RGB444 quantiser ........... c_embedding/yolo_forward.c:57-85
 *   C head ..................... c_embedding/yolo_forward.c:965-1147
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { ROUND_RNE = 0, ROUND_FLOOR = 1, ROUND_HALF_UP = 2 };
enum { CONTRACT_F = 0, CONTRACT_P = 1 };

/* ---- shifts --------------------------------------------------------------------------- */

/* Right shift of x by n >= 0 with the given rounding of the discarded fraction. */
static int64_t shr_round(int64_t x, int n, int mode)
{
    if (n <= 0) return x;
    int64_t fl = x >> n;                         /* floor (arithmetic shift) */
    int64_t rem = x - (fl << n);                 /* 0 <= rem < 2^n */
    int64_t half = (int64_t)1 << (n - 1);
    switch (mode) {
    case ROUND_FLOOR:   return fl;
    case ROUND_HALF_UP: return fl + (rem >= half ? 1 : 0);
    default: /* RNE */
        if (rem > half) return fl + 1;
        if (rem < half) return fl;
        return fl + (fl & 1);                    /* tie -> even */
    }
}

/* sh(x, n, dir): dir == 1 is a left shift, dir == 0 a rounding right shift
 * (direction bits of set_offset, yolo_forward.c:242-254). */
static int64_t sh(int64_t x, int n, int dir, int mode)
{
    return dir ? x * ((int64_t)1 << n) : shr_round(x, n, mode);
}

static int64_t clampi(int64_t x, int64_t lo, int64_t hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ---- one output element ----------------------------------------------------------------- */

typedef struct {
    int sa_i, sw, sb, retune, sa_o;   /* exponents (yolo_forward.c:32-35) */
    int activ;
    int contract, round_mode;
} elem_params;

/* Contract F: t = sh(acc,iofs) + sh(b,bofs); sat16; leaky >>3; (pool); sat8(sh(t,oofs)).
 * All steps are monotone, so the max-pool may be applied anywhere after the accumulation. */
static int64_t requant_F(int64_t acc, int b, const elem_params *p)
{
    /* set_quantize_scale, yolo_forward.c:235-254 */
    int iofs = p->sa_i + p->sw - p->retune, idir = 0;
    int bofs = p->sb - p->retune, bdir = 0;
    int oofs = p->retune - p->sa_o, odir = 0;
    if (iofs < 0) { idir = 1; iofs = -iofs; }
    if (bofs < 0) { bdir = 1; bofs = -bofs; }
    if (oofs < 0) { odir = 1; oofs = -oofs; }
    int64_t t = sh(acc, iofs, idir, p->round_mode) + sh(b, bofs, bdir, p->round_mode);
    t = clampi(t, -32768, 32767);                       /* 16-bit accumulator, slim_yolo_v2.py:223 */
    if (p->activ && t < 0) t = shr_round(t, 3, p->round_mode);   /* leaky 0.125, modules.py:25 */
    int64_t o = sh(t, oofs, odir, p->round_mode);
    return clampi(o, -128, 127);
}

/* Contract P: y = acc*2^-(sa_i+sw) + b*2^-sb; leaky; o = RNE(y*2^sa_o).  The reference never clamps
 * (slim_yolo_v2.py:35): the value is returned unclamped; the int8 STORE saturates (after the pool, as the
 * reference pools its unclamped values, :229-231) and the caller counts how many stored elements that hit. */
static int64_t requant_P(int64_t acc, int b, const elem_params *p)
{
    int ea = p->sa_i + p->sw, eb = p->sb;
    int E = ea > eb ? ea : eb;
    int64_t num = acc * ((int64_t)1 << (E - ea)) + (int64_t)b * ((int64_t)1 << (E - eb));
    int s = E - p->sa_o;
    if (p->activ && num < 0) s += 3;
    return s >= 0 ? shr_round(num, s, ROUND_RNE) : num * ((int64_t)1 << (-s));
}

/* ---- one layer ---------------------------------------------------------------------------- */

/* in : [n][h][w][cs_in]   int8, first cin channels used
 * wgt: [cout][3][3][cin]  int8 (OHWI)
 * out: [n][h'][w'][cs_out] int8, channels >= cout written as 0
 * The 2x2/2 max-pool is applied to the requantised values, as the reference does (quantise, then pool,
 * slim_yolo_v2.py:229-231); contract P values are pooled unclamped (the reference never clamps) and saturate only
 * when stored.  Odd trailing rows/columns are dropped (nn.MaxPool2d(2,2) floor mode).
 * *overflow counts STORED elements that had to be saturated (contract P). */
ORACLE_API int oracle_conv_layer(const int8_t *in, int n, int h, int w, int cs_in, int cin,
                                 const int8_t *wgt, const int8_t *bias, int cout, int cs_out,
                                 int sa_i, int sw, int sb, int retune, int sa_o,
                                 int activ, int pool, int contract, int round_mode,
                                 int8_t *out, int64_t *overflow)
{
    elem_params p = { sa_i, sw, sb, retune, sa_o, activ, contract, round_mode };
    int oh = pool ? h / 2 : h, ow = pool ? w / 2 : w;
    int32_t *full = (int32_t *)malloc(sizeof(int32_t) * (size_t)h * w * cout);
    if (!full) return -1;
    for (int img = 0; img < n; ++img) {
        const int8_t *x = in + (size_t)img * h * w * cs_in;
        /* rows are independent; OpenMP only spreads them over host threads (bench cpu_baseline) */
        #pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y) {
            for (int xx = 0; xx < w; ++xx)
                for (int co = 0; co < cout; ++co) {
                    int32_t acc = 0;                 /* |acc| <= 9*cin*128*128 < 2^31 for cin <= 14000 */
                    for (int kh = 0; kh < 3; ++kh) {
                        int iy = y + kh - 1;
                        if (iy < 0 || iy >= h) continue;           /* zero padding 1 */
                        for (int kw = 0; kw < 3; ++kw) {
                            int ix = xx + kw - 1;
                            if (ix < 0 || ix >= w) continue;
                            const int8_t *a = x + ((size_t)iy * w + ix) * cs_in;
                            const int8_t *k = wgt + (((size_t)co * 3 + kh) * 3 + kw) * cin;
                            for (int ci = 0; ci < cin; ++ci) acc += (int32_t)a[ci] * (int32_t)k[ci];
                        }
                    }
                    int64_t o = contract == CONTRACT_P ? requant_P(acc, bias[co], &p) : requant_F(acc, bias[co], &p);
                    full[((size_t)y * w + xx) * cout + co] = (int32_t)clampi(o, INT32_MIN / 2, INT32_MAX / 2);
                }
        }
        int8_t *o8 = out + (size_t)img * oh * ow * cs_out;
        for (int y = 0; y < oh; ++y)
            for (int xx = 0; xx < ow; ++xx) {
                int8_t *dst = o8 + ((size_t)y * ow + xx) * cs_out;
                for (int co = 0; co < cs_out; ++co) {
                    if (co >= cout) { dst[co] = 0; continue; }
                    int64_t m;
                    if (!pool) m = full[((size_t)y * w + xx) * cout + co];
                    else {
                        m = INT32_MIN;
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                int v = full[((size_t)(2 * y + dy) * w + (2 * xx + dx)) * cout + co];
                                if (v > m) m = v;
                            }
                    }
                    if (m < -128 || m > 127) { if (overflow) (*overflow)++; m = clampi(m, -128, 127); }
                    dst[co] = (int8_t)m;
                }
            }
    }
    free(full);
    return 0;
}

/* Raw int32 accumulators of one layer (for kernel bring-up tests): acc[n][h][w][cout]. */
ORACLE_API int oracle_conv_acc(const int8_t *in, int n, int h, int w, int cs_in, int cin,
                               const int8_t *wgt, int cout, int32_t *acc_out)
{
    for (int img = 0; img < n; ++img)
        for (int y = 0; y < h; ++y)
            for (int xx = 0; xx < w; ++xx)
                for (int co = 0; co < cout; ++co) {
                    int64_t acc = 0;
                    for (int kh = 0; kh < 3; ++kh)
                        for (int kw = 0; kw < 3; ++kw) {
                            int iy = y + kh - 1, ix = xx + kw - 1;
                            if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
                            const int8_t *a = in + (((size_t)img * h + iy) * w + ix) * cs_in;
                            const int8_t *k = wgt + (((size_t)co * 3 + kh) * 3 + kw) * cin;
                            for (int ci = 0; ci < cin; ++ci) acc += (int64_t)a[ci] * k[ci];
                        }
                    acc_out[(((size_t)img * h + y) * w + xx) * cout + co] = (int32_t)acc;
                }
    return 0;
}

/* Scalar helpers exported for property tests of the kernels' epilogue. */
ORACLE_API int oracle_requant(int64_t acc, int b, int sa_i, int sw, int sb, int retune, int sa_o,
                              int activ, int contract, int round_mode)
{
    elem_params p = { sa_i, sw, sb, retune, sa_o, activ, contract, round_mode };
    return (int)clampi(contract == CONTRACT_P ? requant_P(acc, b, &p) : requant_F(acc, b, &p), -128, 127);
}

/* The shift programme set_quantize_scale() hands to set_offset() (yolo_forward.c:233-257):
 * out6 = {iofs, idir, bofs, bdir, oofs, odir}. */
ORACLE_API void oracle_shift_programme(int sa_i, int sw, int sb, int retune, int sa_o, int out6[6])
{
    int v[3] = { sa_i + sw - retune, sb - retune, retune - sa_o };
    for (int i = 0; i < 3; ++i) {
        int d = 0, o = v[i];
        if (o < 0) { d = 1; o = -o; }
        out6[2 * i] = o; out6[2 * i + 1] = d;
    }
}

/* ---- input quantisers ------------------------------------------------------------------- */

/* pixel_norm_quantize, yolo_forward.c:57-85, for one 12-bit 0x0BGR code.  The channel is masked
 * but NOT shifted down (R in 0..15, G in {0,16,..,240}, B in {0,256,..,3840}); float/double mix
 * and the truncating (char) conversion are as in the reference.  out3 = R,G,B. */
ORACLE_API void oracle_rgb444_pixel(int pixel_0bgr, int sa, int8_t out3[3])
{
    static const int    mask[3] = { 0x000f, 0x00f0, 0x0f00 };
    static const double mean[3] = { 0.485, 0.456, 0.406 };
    static const double stdv[3] = { 0.229, 0.224, 0.225 };
    double s = pow(2.0, (double)sa);
    for (int c = 0; c < 3; ++c) {
        float v = (float)(pixel_0bgr & mask[c]);
        v = (float)(v / 255.);
        v = (float)(v - mean[c]);
        v = (float)(v / stdv[c]);
        out3[c] = (int8_t)(v * s);          /* truncation toward zero, :68 */
    }
}

/* frames: uint16 [n][h][w] -> int8 [n][h][w][4] = R,G,B,0 (the word camera_to_inpBuf writes, :95-116) */
ORACLE_API void oracle_quantize_rgb444(const uint16_t *frames, size_t npix, int sa, int8_t *nhwc4)
{
    for (size_t i = 0; i < npix; ++i) {
        int8_t q[3];
        oracle_rgb444_pixel((int16_t)frames[i], sa, q);
        nhwc4[4 * i + 0] = q[0]; nhwc4[4 * i + 1] = q[1]; nhwc4[4 * i + 2] = q[2]; nhwc4[4 * i + 3] = 0;
    }
}

/* a_tracker_in.quantize_activation with a frozen scale 2^sa: round-half-even(x * 2^sa)
 * (slim_yolo_v2.py:33-35).  float NCHW [n][3][h][w] -> int8 NHWC4. */
ORACLE_API void oracle_quantize_f32(const float *nchw, int n, int h, int w, int sa, int8_t *nhwc4,
                                    int64_t *overflow)
{
    float s = ldexpf(1.0f, sa);
    size_t plane = (size_t)h * w;
    for (int img = 0; img < n; ++img)
        for (size_t i = 0; i < plane; ++i) {
            for (int c = 0; c < 3; ++c) {
                float r = nearbyintf(nchw[((size_t)img * 3 + c) * plane + i] * s);
                if (r < -128.f || r > 127.f) { if (overflow) (*overflow)++; r = r < 0 ? -128.f : 127.f; }
                nhwc4[((size_t)img * plane + i) * 4 + c] = (int8_t)r;
            }
            nhwc4[((size_t)img * plane + i) * 4 + 3] = 0;
        }
}

/* ---- detection head, Python semantics ----------------------------------------------------- */

typedef struct {
    float x1, y1, x2, y2, score;
    int32_t cls, anchor_index, pad_;
} oracle_det;

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* Decode every anchor of one frame (slim_yolo_v2.py:333-350).
 * pred: [gh][gw][cs] int8, channel order: A conf | A*C cls (anchor-major) | A*4 box (anchor-major).
 * boxes: [gh*gw*A][4] normalised+clamped, scores/cls: best class score and index. */
ORACLE_API void oracle_decode_python(const int8_t *pred, int gh, int gw, int cs, int A, int C,
                                     int sa_pred, const float *anchors /*[A][2]*/, int stride,
                                     int in_h, int in_w, float *boxes, float *scores, int32_t *cls)
{
    float inv = ldexpf(1.0f, -sa_pred);
    for (int cell = 0; cell < gh * gw; ++cell) {
        const int8_t *p = pred + (size_t)cell * cs;
        int row = cell / gw, col = cell % gw;
        for (int a = 0; a < A; ++a) {
            float obj = sigmoidf_(p[a] * inv);
            const int8_t *pc = p + A + a * C;
            float m = -INFINITY;
            for (int c = 0; c < C; ++c) { float v = pc[c] * inv; if (v > m) m = v; }
            float sum = 0.f, e[64];
            for (int c = 0; c < C; ++c) { e[c] = expf(pc[c] * inv - m); sum += e[c]; }
            int best = 0; float bs = -1.f;
            for (int c = 0; c < C; ++c) { float s = e[c] / sum * obj; if (s > bs) { bs = s; best = c; } }
            const int8_t *pb = p + A * (1 + C) + a * 4;
            float cx = (sigmoidf_(pb[0] * inv) + (float)col) * (float)stride;
            float cy = (sigmoidf_(pb[1] * inv) + (float)row) * (float)stride;
            float bw = expf(pb[2] * inv) * anchors[2 * a + 0] * (float)stride;
            float bh = expf(pb[3] * inv) * anchors[2 * a + 1] * (float)stride;
            float b4[4] = { (cx - bw / 2) / in_w, (cy - bh / 2) / in_h, (cx + bw / 2) / in_w, (cy + bh / 2) / in_h };
            int idx = cell * A + a;
            for (int k = 0; k < 4; ++k) boxes[4 * idx + k] = fminf(fmaxf(b4[k], 0.f), 1.f);
            scores[idx] = bs; cls[idx] = best;
        }
    }
}

typedef struct { float s; int i; } sort_key;
static int cmp_desc(const void *a, const void *b)
{
    const sort_key *x = (const sort_key *)a, *y = (const sort_key *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return y->i - x->i;      /* ties: higher index first (= reversed stable ascending argsort) */
}

/* postprocess + nms (slim_yolo_v2.py:145-210): threshold score >= conf, per-class greedy NMS keeping
 * ovr <= thresh, result in ascending anchor order.  Returns the number kept (<= max_det written). */
ORACLE_API int oracle_nms_python(const float *boxes, const float *scores, const int32_t *cls, int N, int C,
                                 float conf_thresh, float nms_thresh, oracle_det *dets, int max_det)
{
    int *cand = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    char *keep = (char *)calloc((size_t)(N > 0 ? N : 1), 1);
    sort_key *keys = (sort_key *)malloc(sizeof(sort_key) * (size_t)(N > 0 ? N : 1));
    for (int c = 0; c < C; ++c) {
        int m = 0;
        for (int i = 0; i < N; ++i)
            if (scores[i] >= conf_thresh && cls[i] == c) { keys[m].s = scores[i]; keys[m].i = i; ++m; }
        qsort(keys, (size_t)m, sizeof(sort_key), cmp_desc);
        for (int i = 0; i < m; ++i) cand[i] = keys[i].i;
        char *dead = (char *)calloc((size_t)(m > 0 ? m : 1), 1);
        for (int i = 0; i < m; ++i) {
            if (dead[i]) continue;
            int bi = cand[i]; keep[bi] = 1;
            float ax1 = boxes[4 * bi], ay1 = boxes[4 * bi + 1], ax2 = boxes[4 * bi + 2], ay2 = boxes[4 * bi + 3];
            float aarea = (ax2 - ax1) * (ay2 - ay1);
            for (int j = i + 1; j < m; ++j) {
                if (dead[j]) continue;
                int bj = cand[j];
                float bx1 = boxes[4 * bj], by1 = boxes[4 * bj + 1], bx2 = boxes[4 * bj + 2], by2 = boxes[4 * bj + 3];
                float barea = (bx2 - bx1) * (by2 - by1);
                float ww = fmaxf(1e-28f, fminf(ax2, bx2) - fmaxf(ax1, bx1));
                float hh = fmaxf(1e-28f, fminf(ay2, by2) - fmaxf(ay1, by1));
                float inter = ww * hh;
                float ovr = inter / (aarea + barea - inter);
                if (!(ovr <= nms_thresh)) dead[j] = 1;
            }
        }
        free(dead);
    }
    int cnt = 0;
    for (int i = 0; i < N; ++i)
        if (keep[i]) {
            if (cnt < max_det) {
                oracle_det *d = &dets[cnt];
                d->x1 = boxes[4 * i]; d->y1 = boxes[4 * i + 1]; d->x2 = boxes[4 * i + 2]; d->y2 = boxes[4 * i + 3];
                d->score = scores[i]; d->cls = cls[i]; d->anchor_index = i; d->pad_ = 0;
            }
            ++cnt;
        }
    free(cand); free(keep); free(keys);
    return cnt;
}

/* ---- detection head, C semantics (well-defined subset of yolo_forward.c:965-1147) -------------
 * Restated with correct channel addressing and de-quantised tx..th (the literal code feeds raw int8
 * into exp(), :1079-1085, and drifts its pointers, :1078-1107 — both recorded in oracle/DEVIATIONS.md).
 * Kept as written: sigma(-x) (:965-968), softmax over 2 classes (:976-987), argmax with ties -> class 0
 * (:989-997), score > thresh strictly (:1077), h decoded with the anchor WIDTH (:1044), corners
 * truncated toward zero to int pixels (:1045-1048), descending selection sort with strict '>'
 * (:1114-1126), class-agnostic NMS suppressing iou >= thresh on integer boxes (:1000-1036,1128-1147). */
static float c_sigmoid(float x) { return (float)(1 / (exp((double)x) + 1)); }

/* conf_sort: selection sort by swapping, strict '>' (yolo_forward.c:1114-1126).  The order it leaves among EQUAL scores
 * depends on the swap history; NMS (:1128-1147) is order dependent, so the sort is restated literally. */
static void c_conf_sort(oracle_det *b, int m)
{
    for (int i = 0; i < m - 1; ++i)
        for (int j = i + 1; j < m; ++j)
            if (b[j].score > b[i].score) { oracle_det t = b[j]; b[j] = b[i]; b[i] = t; }
}

/* NMS (:1128-1147): class-agnostic, suppress iou >= thresh, on integer boxes (box_iou :1000-1036). dead[] must be zeroed. */
static void c_nms(const oracle_det *b, int m, float nms_thresh, char *dead)
{
    for (int i = 0; i < m; ++i) {
        if (dead[i]) continue;
        for (int j = i + 1; j < m; ++j) {
            int ax1 = (int)b[i].x1, ax2 = (int)b[i].x2, ay1 = (int)b[i].y1, ay2 = (int)b[i].y2;
            int bx1 = (int)b[j].x1, bx2 = (int)b[j].x2, by1 = (int)b[j].y1, by2 = (int)b[j].y2;
            /* overlap(): sum of sides minus hull (:1000-1004) */
            int ow = (ax2 - ax1 + bx2 - bx1) - ((ax2 > bx2 ? ax2 : bx2) - (ax1 <= bx1 ? ax1 : bx1));
            int oh = (ay2 - ay1 + by2 - by1) - ((ay2 > by2 ? ay2 : by2) - (ay1 <= by1 ? ay1 : by1));
            int inter = (ow <= 0 || oh <= 0) ? 0 : ow * oh;
            int uni = (ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter;
            float iou = (float)inter / (float)uni;
            if (iou >= nms_thresh) dead[j] = 1;
        }
    }
}

/* The sort + NMS of the restated C head on caller-provided boxes {x_max, x_min, y_max, y_min} and scores (the same argument
 * convention as oracle/tierA_shim.c: tierA_sort_nms, so tests can lay the two side by side, ties included): order[i] =
 * original index at sorted position i, suppressed[i] = its flag.  Returns the number kept. */
ORACLE_API int oracle_conf_sort_nms(int n, const int *boxes, const float *conf, float thresh, int *order, int *suppressed)
{
    oracle_det *b = (oracle_det *)malloc(sizeof(oracle_det) * (size_t)(n > 0 ? n : 1));
    char *dead = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n; ++i) {
        b[i].x2 = (float)boxes[4 * i]; b[i].x1 = (float)boxes[4 * i + 1]; b[i].y2 = (float)boxes[4 * i + 2]; b[i].y1 = (float)boxes[4 * i + 3];
        b[i].score = conf[i]; b[i].cls = 0; b[i].anchor_index = i; b[i].pad_ = 0;
    }
    c_conf_sort(b, n);
    c_nms(b, n, thresh, dead);
    int kept = 0;
    for (int i = 0; i < n; ++i) { order[i] = b[i].anchor_index; suppressed[i] = dead[i]; kept += !dead[i]; }
    free(dead); free(b);
    return kept;
}


ORACLE_API int oracle_head_c(const int8_t *pred, int gh, int gw, int cs, int A,
                             int sa_pred, const float *anchors, int stride,
                             float conf_thresh, float nms_thresh, oracle_det *dets, int max_det)
{
    int N = gh * gw * A, m = 0;
    oracle_det *b = (oracle_det *)malloc(sizeof(oracle_det) * (size_t)(N > 0 ? N : 1));
    double sc = pow(2.0, (double)sa_pred);
    for (int cell = 0; cell < gh * gw; ++cell) {
        const int8_t *p = pred + (size_t)cell * cs;
        int row = cell / gw, col = cell % gw;
        for (int a = 0; a < A; ++a) {
            float conf = c_sigmoid((float)(p[a] / sc));
            float cl[2] = { (float)(p[A + 2 * a] / sc), (float)(p[A + 2 * a + 1] / sc) };
            float sum = 0;
            for (int i = 0; i < 2; ++i) { cl[i] = (float)exp((double)cl[i]); sum += cl[i]; }
            for (int i = 0; i < 2; ++i) cl[i] = cl[i] / sum;
            int c = cl[0] >= cl[1] ? 0 : 1;
            conf = conf * cl[c];
            if (!(conf > conf_thresh)) continue;
            const int8_t *pb = p + A * 3 + a * 4;
            float tx = (float)(pb[0] / sc), ty = (float)(pb[1] / sc), tw = (float)(pb[2] / sc), th = (float)(pb[3] / sc);
            float xc = (c_sigmoid(tx) + col) * stride;
            float yc = (c_sigmoid(ty) + row) * stride;
            float bw = (float)(anchors[2 * a] * exp((double)tw) * stride);
            float bh = (float)(anchors[2 * a] * exp((double)th) * stride);   /* anchor WIDTH, :1044 */
            oracle_det *d = &b[m++];
            d->x1 = (float)(int)(xc - bw / 2); d->x2 = (float)(int)(xc + bw / 2);
            d->y1 = (float)(int)(yc - bh / 2); d->y2 = (float)(int)(yc + bh / 2);
            d->score = conf; d->cls = c; d->anchor_index = cell * A + a; d->pad_ = 0;
        }
    }
    c_conf_sort(b, m);
    char *dead = (char *)calloc((size_t)(m > 0 ? m : 1), 1);
    c_nms(b, m, nms_thresh, dead);
    int cnt = 0;
    for (int i = 0; i < m; ++i) {
        if (dead[i]) continue;
        if (cnt < max_det) dets[cnt] = b[i];
        ++cnt;
    }
    free(dead); free(b);
    return cnt;
}

/* Number of OpenMP threads the oracle's loops use (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, and the
 * OpenMP runtime may have read it before this library was loaded: bench.py sets the count explicitly).  Returns the maximum. */
#ifdef _OPENMP
#include <omp.h>
#endif
ORACLE_API int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ---- whole network ------------------------------------------------------------------------ */

typedef struct {
    int num_layers;
    int cin[32], cout[32], activ[32], pool[32];
    int sw[32], sb[32], sa[33], retune[32];
    int contract, round_mode;
} oracle_net;

/* Runs all layers from an int8 NHWC4 input.  weights[l]: OHWI int8, biases[l]: int8.
 * layer_out[l] (optional, may be NULL entries) receives [n][h_l][w_l][cs(cout_l)] with cs = round-up-16.
 * Returns 0, and the last layer's grid in *gh,*gw. */
ORACLE_API int oracle_backbone(const oracle_net *net, const int8_t *const *weights, const int8_t *const *biases,
                               const int8_t *nhwc4, int n, int h, int w, int8_t *const *layer_out,
                               int *gh, int *gw, int64_t *overflow)
{
    const int8_t *cur = nhwc4; int cs_in = 4; int8_t *owned = NULL;
    for (int l = 0; l < net->num_layers; ++l) {
        int cs_out = (net->cout[l] + 15) / 16 * 16;
        int oh = net->pool[l] ? h / 2 : h, ow = net->pool[l] ? w / 2 : w;
        int8_t *out = (int8_t *)malloc((size_t)n * oh * ow * cs_out);
        if (!out) return -1;
        int rc = oracle_conv_layer(cur, n, h, w, cs_in, net->cin[l], weights[l], biases[l], net->cout[l], cs_out,
                                   net->sa[l], net->sw[l], net->sb[l], net->retune[l], net->sa[l + 1],
                                   net->activ[l], net->pool[l], net->contract, net->round_mode, out, overflow);
        if (rc) return rc;
        if (layer_out && layer_out[l]) memcpy(layer_out[l], out, (size_t)n * oh * ow * cs_out);
        free(owned); owned = out; cur = out; cs_in = cs_out; h = oh; w = ow;
    }
    free(owned);
    *gh = h; *gw = w;
    return 0;
}
