/*
 * oracle/tierB_shim.c — ORACLE (test infrastructure), tier B of SURVEY.md 8c: the reference's LITERAL first-layer driver
 * (first_conv, c_embedding/yolo_forward.c:269-418, with camera_to_inpBuf :87-123, load_weight :165-173, load_bias
 * :175-177, config_ctl_reg :179-197, outBuf_to_psRam :199-215, set_quantize_scale :233-257) run over a HOST MODEL of the
 * accelerator it drives.  The RTL is not in the reference, so the model below is this repository's reading of the 11
 * intrinsics' call sites; its arithmetic is Contract F with round-half-even shifts.
 *
 * Purpose: find out how far a literal run of the driver gets, and ledger every divergence from the clean restatement
 * (oracle/ref_int8.c) in oracle/DEVIATIONS.md with the reference line that causes it.  It is NOT a parity target.
 *
 * Compiled by oracle/Makefile like tierA: the reference TU is included where it lies (one-token patched temporary copy);
 * nothing of it is stored in this repository.
 *
 * Accelerator model (one "address" = one entry):
 *   input buffer : 2 x 396 entries (inp_buf 0x0 / inp_buf2 0x18c = 18 * 22, :20-21), entry = one pixel = up to 16 bytes
 *                  (write_input_buffer(src, addr, words - 1), :116 one word, :154 Tn * 8 / 32 words)
 *   kernel buffer: 2 x 9 entries (kernel_buf 0x0 / kernel_buf2 0x9, :22-23), entry = one tap = Tm x Tn bytes, [Tm][Tn]
 *                  (write_kernel_buffer(weight + tap * Cin * Cout, addr + tap, ...), :169)
 *   bias buffer  : 32 int8 (write_bias_buffer(b, addr, 7), :176)
 *   output buffer: 320 entries (16 x 20 tile), entry = one output pixel = Tm bytes (read_output_buffer, :206)
 *   PSRAM        : flat byte array (read_psram / write_psram(ptr, byte address, bits - 1), :153, :208)
 *   set_tile_info(Tn, Tm, Tc, Tr) :310, set_offset(iofs, idir, bofs, bdir, oofs, odir) :256,
 *   set_tile_detile(pad_pos, pp_i, pp_w, pp_b, first_cg, last_cg, activ, pool) :195, start_calculate() :404,
 *   wait_cal_done() :388.
 *   start_calculate(): the input tile is (Tr + 2) x (Tc + 2) entries at the ping-pong base; the sides flagged in pad_pos are
 *   zero-filled by the "hardware" (the driver never writes them, :101-113); 3x3 MACs over Tn x Tm, the three shifts,
 *   16-bit saturation, leaky >> 3, 2 x 2 max-pool inside the tile, 8-bit saturation; results land in the output buffer.
 */
#include <stdint.h>
#include <stdbool.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#define INP_ENTRIES 792
#define OUT_ENTRIES 400
#define PSRAM_BYTES (1u << 23)        /* psram_max, :29 */

static struct {
    int8_t inp[INP_ENTRIES][16];
    int8_t ker[18][512];
    int8_t bias[64];
    int8_t out[OUT_ENTRIES][32];
    int8_t *ps_mem;
    int Tn, Tm, Tc, Tr;
    int iofs, idir, bofs, bdir, oofs, odir;
    int pad_pos, pp_i, pp_w, pp_b, first_cg, last_cg, activ, pool;
    int32_t acc[OUT_ENTRIES * 4][32];
    long tiles, oob_inp, oob_out, oob_psram, zero_tiles, corner_tiles;
} M;

static int shr_rne(int x, int n)
{
    if (n <= 0) return x;
    int fl = x >> n, rem = x - (fl << n), half = 1 << (n - 1);
    if (rem != half) return fl + (rem > half);
    return fl + (fl & 1);
}
static int sh(int x, int n, int dir) { return dir ? x * (1 << n) : shr_rne(x, n); }
static int sat(int x, int lo, int hi) { return x < lo ? lo : x > hi ? hi : x; }

static void write_input_buffer(uint32_t *src, uint32_t addr, int words_m1)
{
    if (addr >= INP_ENTRIES) { M.oob_inp++; return; }
    memset(M.inp[addr], 0, 16);
    memcpy(M.inp[addr], src, 4u * (unsigned)(words_m1 + 1) > 16u ? 16u : 4u * (unsigned)(words_m1 + 1));
}
static void write_kernel_buffer(uint32_t *src, uint32_t addr, int words_m1)
{
    if (addr >= 18) return;
    unsigned n = 4u * (unsigned)(words_m1 + 1);
    memcpy(M.ker[addr], src, n > 512u ? 512u : n);
}
static void write_bias_buffer(uint32_t *src, uint32_t addr, int words_m1) { (void)addr; memcpy(M.bias, src, 4u * (unsigned)(words_m1 + 1)); }
static void read_output_buffer(uint32_t *dst, uint32_t addr, int words_m1)
{
    if (addr >= OUT_ENTRIES) { M.oob_out++; memset(dst, 0, 4u * (unsigned)(words_m1 + 1)); return; }
    memcpy(dst, M.out[addr], 4u * (unsigned)(words_m1 + 1));
}
static void read_psram(uint32_t *dst, uint32_t addr, int bits_m1)
{
    unsigned n = (unsigned)(bits_m1 + 1) / 8u;
    if ((uint64_t)addr + n > PSRAM_BYTES) { M.oob_psram++; memset(dst, 0, n); return; }
    memcpy(dst, M.ps_mem + addr, n);
}
static void write_psram(uint32_t *src, uint32_t addr, int bits_m1)
{
    unsigned n = (unsigned)(bits_m1 + 1) / 8u;
    if ((uint64_t)addr + n > PSRAM_BYTES) { M.oob_psram++; return; }
    memcpy(M.ps_mem + addr, src, n);
}
static void set_tile_info(int Tn, int Tm, int Tc, int Tr) { M.Tn = Tn; M.Tm = Tm; M.Tc = Tc; M.Tr = Tr; }
static void set_offset(int a, int b, int c, int d, int e, int f) { M.iofs = a; M.idir = b; M.bofs = c; M.bdir = d; M.oofs = e; M.odir = f; }
static void set_tile_detile(int pad, int pi, int pw, int pb, int fc, int lc, int act, int pool)
{ M.pad_pos = pad; M.pp_i = pi; M.pp_w = pw; M.pp_b = pb; M.first_cg = fc; M.last_cg = lc; M.activ = act; M.pool = pool; }
static void wait_cal_done(void) {}

#define MODEL_PAD_UP 1
#define MODEL_PAD_DOWN 2
#define MODEL_PAD_LEFT 4
#define MODEL_PAD_RIGHT 8
static int g_pad_bits[4];          /* the TU's PADDING_UP/DOWN/LEFT/RIGHT values, filled in after the include */

static void start_calculate(void)
{
    const int Tr = M.Tr, Tc = M.Tc, Tn = M.Tn, Tm = M.Tm, TRow = Tr + 2, TCol = Tc + 2;
    M.tiles++;
    if (Tr <= 0 || Tc <= 0) { M.zero_tiles++; return; }
    /* pp_i names the input buffer just filled (:352-359 fill, then :397 passes the same flag); pp_w arrives ALREADY inverted
     * (:333-339 load, invert, then :397), i.e. it names the buffer to load NEXT: compute from the other one */
    const int ibase = M.pp_i ? 0x18c : 0, kbase = M.pp_w ? 0 : 9;
    /* the hardware zero-fills the flagged sides */
    if (M.pad_pos & g_pad_bits[0]) for (int c = 0; c < TCol; ++c) memset(M.inp[ibase + c], 0, 16);
    if (M.pad_pos & g_pad_bits[1]) for (int c = 0; c < TCol; ++c) memset(M.inp[ibase + (TRow - 1) * TCol + c], 0, 16);
    if (M.pad_pos & g_pad_bits[2]) for (int r = 0; r < TRow; ++r) memset(M.inp[ibase + r * TCol], 0, 16);
    if (M.pad_pos & g_pad_bits[3]) for (int r = 0; r < TRow; ++r) memset(M.inp[ibase + r * TCol + TCol - 1], 0, 16);
    for (int r = 0; r < Tr; ++r)
        for (int c = 0; c < Tc; ++c)
            for (int m = 0; m < Tm; ++m) {
                int32_t a = M.first_cg ? 0 : M.acc[r * Tc + c][m];
                for (int t = 0; t < 9; ++t) {
                    const int8_t *px = M.inp[ibase + (r + t / 3) * TCol + (c + t % 3)];
                    const int8_t *w = &M.ker[kbase + t][m * Tn];
                    for (int n = 0; n < Tn; ++n) a += (int32_t)px[n] * (int32_t)w[n];
                }
                M.acc[r * Tc + c][m] = a;
            }
    if (!M.last_cg) return;
    const int oTr = M.pool ? Tr / 2 : Tr, oTc = M.pool ? Tc / 2 : Tc;
    for (int r = 0; r < oTr; ++r)
        for (int c = 0; c < oTc; ++c)
            for (int m = 0; m < Tm; ++m) {
                int best = -1000000;
                for (int dy = 0; dy < (M.pool ? 2 : 1); ++dy)
                    for (int dx = 0; dx < (M.pool ? 2 : 1); ++dx) {
                        const int rr = M.pool ? 2 * r + dy : r, cc = M.pool ? 2 * c + dx : c;
                        int t = sh(M.acc[rr * Tc + cc][m], M.iofs, M.idir) + sh(M.bias[m], M.bofs, M.bdir);
                        t = sat(t, -32768, 32767);
                        if (M.activ && t < 0) t = shr_rne(t, 3);
                        t = sat(sh(t, M.oofs, M.odir), -128, 127);
                        if (t > best) best = t;
                    }
                M.out[r * oTc + c][m] = (int8_t)best;
            }
}

#undef printf
#define printf(...) ((void)0)

#include REF_TU   /* the reference translation unit (patched copy made by the Makefile) */

#define API __attribute__((visibility("default")))

/* Runs the reference's first_conv (:269-418) literally on one 240 x 320 RGB444 frame with the as-shipped tables
 * (layer_index 0 of scale_a/scale_w/scale_b/retune, :32-35) and the caller's weight.h-order weights / biases, then drains
 * the last tile the way the next layer would (outBuf_to_psRam with the returned INFO, second_conv :497-499).
 * psram_out: the first out_bytes of the PSRAM page first_conv wrote (psram, :28), i.e. [120][160][16] if the driver is right.
 * stats[6]: tiles started, zero-height tiles, out-of-range input-buffer / output-buffer / PSRAM accesses, 0. */
API int tierB_first_conv(const short *frame /*[240*320], with >= 32 rows of slack after it*/, const int8_t *w_h_order /*[9][16][3]*/,
                         const int8_t *bias32 /*32 bytes*/, int8_t *psram_out, int out_bytes, long stats[6])
{
    memset(&M, 0, sizeof M);
    M.ps_mem = (int8_t *)calloc(PSRAM_BYTES, 1);
    if (!M.ps_mem) return -1;
    g_pad_bits[0] = PADDING_UP; g_pad_bits[1] = PADDING_DOWN; g_pad_bits[2] = PADDING_LEFT; g_pad_bits[3] = PADDING_RIGHT;
    bool pingpong[3] = { 0, 0, 0 };
    int save[8];
    struct INFO info = first_conv(320, 240, 16, 20, 16, 3, 18, 22, 0, (short *)frame, save, (char *)bias32, (char *)w_h_order, 1, 1, pingpong, 3);
    outBuf_to_psRam(out_buf, info.psram_addr, save, info.out_w, info.Tm, info.out_Tr, info.out_Tc);
    memcpy(psram_out, M.ps_mem, (size_t)out_bytes);
    stats[0] = M.tiles; stats[1] = M.zero_tiles; stats[2] = M.oob_inp; stats[3] = M.oob_out; stats[4] = M.oob_psram; stats[5] = 0;
    free(M.ps_mem);
    return 0;
}
