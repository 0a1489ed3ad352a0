/*
 * oracle/tierA_shim.c — ORACLE (test infrastructure).  Glue that lets the reference's OWN functions
 * from c_embedding/yolo_forward.c be called on the host ("tier A").
 *
 * The Makefile compiles this file with -DREF_TU='"<patched temporary copy of yolo_forward.c>"'.
 * The reference translation unit is included verbatim; nothing of it is stored in this repository.
 * What this shim supplies:
 *   - weight.h is missing from the reference checkout (.MISSING_LARGE_BLOBS:1): an empty stand-in with
 *     the ten w_conv/b_conv symbols (only their addresses are taken, yolo_forward.c:1204-1260);
 *   - the 11 accelerator intrinsics (RTL absent): no-op stubs, except set_offset which records its
 *     arguments so set_quantize_scale (yolo_forward.c:233-257) can be checked;
 *   - printf is compiled away (box_union prints ints with %f, yolo_forward.c:1024).
 * Exported wrappers call the reference functions unchanged.
 */
#include <stdint.h>
#include <stdbool.h>
#include <string.h>
#include <stdio.h>

static int g_offset_args[6];
static void set_offset(int a, int b, int c, int d, int e, int f)
{ g_offset_args[0] = a; g_offset_args[1] = b; g_offset_args[2] = c; g_offset_args[3] = d; g_offset_args[4] = e; g_offset_args[5] = f; }
static void write_input_buffer(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void write_kernel_buffer(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void write_bias_buffer(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void read_output_buffer(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void read_psram(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void write_psram(uint32_t *s, uint32_t a, int n) { (void)s; (void)a; (void)n; }
static void set_tile_info(int a, int b, int c, int d) { (void)a; (void)b; (void)c; (void)d; }
static void set_tile_detile(int a, int b, int c, int d, int e, int f, int g, int h)
{ (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; }
static void start_calculate(void) {}
static void wait_cal_done(void) {}

#undef printf
#define printf(...) ((void)0)

#include REF_TU   /* the reference translation unit (patched copy made by the Makefile) */

#define API __attribute__((visibility("default")))

API void tierA_pixel_norm_quantize(int pixel, int sa, signed char out3[3])
{
    int word = 0;
    pixel_norm_quantize((short)pixel, (char)sa, (char *)&word);
    memcpy(out3, &word, 3);
}
API void tierA_set_quantize_scale(int sa_i, int sw, int sb, int rt, int sa_o, int out6[6])
{
    set_quantize_scale((char)sa_i, (char)sw, (char)sb, (char)rt, (char)sa_o);
    memcpy(out6, g_offset_args, sizeof g_offset_args);
}
API float tierA_sigmoid(float x) { return sigmoid(x); }
API float tierA_dequantize(int q, int sa) { return dequantize((char)q, (char)sa); }
API void tierA_softmax(float v[2]) { softmax(v); }
API int tierA_cls_sort(float v[2]) { return cls_sort((int *)v); }
API float tierA_box_iou(const int a[4], const int b[4])   /* {x_max,x_min,y_max,y_min} */
{
    struct BOX A = { a[0], a[1], a[2], a[3], 0, 0, 0 }, B = { b[0], b[1], b[2], b[3], 0, 0, 0 };
    return box_iou(A, B);
}
API void tierA_decode_txtytwth(float tx, float ty, float tw, float th, int cx, int cy, int a, int out4[4])
{
    struct BOX b = decode_txtytwth(tx, ty, tw, th, (char)cx, (char)cy, (char)a);
    out4[0] = b.x_max; out4[1] = b.x_min; out4[2] = b.y_max; out4[3] = b.y_min;
}
/* conf_sort + NMS on caller-provided boxes: boxes[i] = {x_max,x_min,y_max,y_min}, conf[i];
 * on return order[i] is the original index at sorted position i and suppressed[i] its flag. */
API int tierA_sort_nms(int n, const int *boxes, const float *conf, float thresh, int *order, int *suppressed)
{
    struct BOX *b = (struct BOX *)calloc((size_t)n + 1, sizeof(struct BOX));
    for (int i = 0; i < n; ++i) {
        b[i].x_max = boxes[4 * i]; b[i].x_min = boxes[4 * i + 1]; b[i].y_max = boxes[4 * i + 2]; b[i].y_min = boxes[4 * i + 3];
        b[i].conf = conf[i]; b[i].cls = 0; b[i].supression = 0;
        /* stash the original index in cls' padding-free neighbour: use a parallel array instead */
    }
    /* keep original indices by sorting a copy of conf alongside: conf values are made unique by the caller */
    conf_sort(b, n);
    int kept = NMS(b, n, thresh) - 1;      /* NMS returns kept+1 (:1146) */
    for (int i = 0; i < n; ++i) {
        order[i] = -1;
        for (int j = 0; j < n; ++j)
            if (conf[j] == b[i].conf && boxes[4 * j] == b[i].x_max && boxes[4 * j + 1] == b[i].x_min &&
                boxes[4 * j + 2] == b[i].y_max && boxes[4 * j + 3] == b[i].y_min) { order[i] = j; break; }
        suppressed[i] = b[i].supression;
    }
    free(b);
    return kept;
}
API const signed char *tierA_tables(int which)   /* 0 scale_w, 1 scale_b, 2 scale_a, 3 retune */
{
    switch (which) { case 0: return (const signed char *)scale_w; case 1: return (const signed char *)scale_b;
                     case 2: return (const signed char *)scale_a; default: return (const signed char *)retune; }
}
API const float *tierA_anchors(void) { return &anchor_size[0][0]; }
