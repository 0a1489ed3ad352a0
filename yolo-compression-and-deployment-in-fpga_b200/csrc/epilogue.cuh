// epilogue.cuh — requantisation epilogue pieces shared by the tcgen05 convolution kernels: P is any parameter struct
// with members `LayerQ q` and `EpiConst k`.
#pragma once
#include <cstring>
#include <math.h>
#include "kernels.h"

namespace yb {

// ---- packed fp32 (Blackwell f32x2 pipe) and saturating pack helpers ------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// d = sat8(a) | sat8(b) << 8 | c << 16   (cvt.pack.sat: two s32 -> two saturated bytes, merged above the low half of c)
__device__ __forceinline__ unsigned pack_sat_s8(int a, int b, unsigned c)
{
    unsigned d;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(a), "r"(c));
    return d;
}
__device__ __forceinline__ unsigned pack_sat_u8(int a, int b, unsigned c)
{
    unsigned d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(a), "r"(c));
    return d;
}
#define YB_MAGIC_BITS 0x4B400000
// bits - MAGIC_BITS as a multiply-add with the run-time constant one = 1 (EpiConst::one): issues on the FMA pipe
__device__ __forceinline__ int unbias_magic(int bits, int one)
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(bits), "r"(one), "n"(-YB_MAGIC_BITS));
    return r;
}

// Contract F, round-half-even, two channels at a time.  fb = (float)sh(bias).  Returns o as an integer that is
// exact inside [-128,127] and on the correct side outside it (the pack saturates).  See requant_f_rne in common.cuh for
// the arithmetic; HI16 = false drops the upper 16-bit clamp when the host proved it cannot change the result.
// The LOWER 16-bit clamp always runs: it keeps the final float positive, which the integer view of the result needs.
// With the leaky-ReLU it is folded into the same 3-input maximum: l = rne(v / 8) is taken from the UNclamped v, and
// max(v, l, MAGIC - 4096) equals leaky(max(v, MAGIC - 32768)) because rne(./8) is monotone and leaky(-32768) = -4096
// (one FMNMX3 instead of two FMNMX; the float min/max share the half-rate ALU pipe with the integer work).
// PRE: the accumulators were started at YB_MAGIC_BITS, so their bit pattern read as fp32 is MAGIC + sum exactly
// (|sum| < 2^22); the int->float conversion disappears and the first FMA's addend becomes MAGIC * (1 - s_in), which
// gives the same real number sum * s_in + MAGIC before the FMA's single rounding.
template <bool ACT, bool HI16, bool PRE = false>
__device__ __forceinline__ int2 requant_f_rne_x2(int a0, int a1, float2 fb, const EpiConst &k)
{
    float2 v;
    if (PRE) v = ffma2(make_float2(__int_as_float(a0), __int_as_float(a1)), make_float2(k.s_in, k.s_in), make_float2(k.in_add, k.in_add));
    else v = ffma2(make_float2(__int2float_rn(a0), __int2float_rn(a1)), make_float2(k.s_in, k.s_in), make_float2(YB_MAGIC, YB_MAGIC));
    v = fadd2(v, fb);
    if (ACT) {
        const float2 l = ffma2(v, make_float2(0.125f, 0.125f), make_float2(k.leak_add, k.leak_add));
        v.x = fmaxf(fmaxf(v.x, l.x), YB_MAGIC - 4096.f); v.y = fmaxf(fmaxf(v.y, l.y), YB_MAGIC - 4096.f);
    } else {
        v.x = fmaxf(v.x, YB_MAGIC - 32768.f); v.y = fmaxf(v.y, YB_MAGIC - 32768.f);
    }
    if (HI16) { v.x = fminf(v.x, YB_MAGIC + 32767.f); v.y = fminf(v.y, YB_MAGIC + 32767.f); }
    v = ffma2(v, make_float2(k.s_out, k.s_out), make_float2(k.out_add, k.out_add));
    return make_int2(unbias_magic(__float_as_int(v.x), k.one), unbias_magic(__float_as_int(v.y), k.one));
}

// ---- the same with the constants held as packed register pairs ------------------------------------------------------
// requant_f_rne_x2 builds every f32x2 operand ({s_in, s_in}, {MAGIC, MAGIC}, ...) from the constant bank at each use (conv3_1, ncu
// source counters: 38 IMAD.MOV + 28 LDC / LDCU per 32 values).  EpiPk holds the pairs in registers for the lifetime of an epilogue
// warp; `pin` keeps the compiler from re-materialising them.  Used by conv_fs.cu (no spills at 64 registers).  In conv_ws.cu the
// same change removed ~18 % of the epilogue's instructions and did NOT make the layers faster (conv3_1 0.101 -> 0.107 ms, conv3_2
// 0.113 -> 0.110): those epilogues are paced by TMEM-load and dependency latency with two groups, not by issue slots; not kept there.
struct EpiPk { unsigned long long s_in, magic, in_add, eighth, leak_add, s_out, out_add; int one; };
__device__ __forceinline__ unsigned long long epi_pair(float v)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    asm volatile("" : "+l"(r));                                // pin
    return r;
}
__device__ __forceinline__ EpiPk epi_pack(const EpiConst &k)
{
    EpiPk e;
    e.s_in = epi_pair(k.s_in); e.magic = epi_pair(YB_MAGIC); e.in_add = epi_pair(k.in_add); e.eighth = epi_pair(0.125f);
    e.leak_add = epi_pair(k.leak_add); e.s_out = epi_pair(k.s_out); e.out_add = epi_pair(k.out_add); e.one = k.one;
    return e;
}
__device__ __forceinline__ unsigned long long ffma2_pk(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long fadd2_pk(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long pack_f2(float x, float y)
{ unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 unpack_f2(unsigned long long v)
{ float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(v)); return d; }

// Contract F / RNE, two channels, packed constants; fb = the two fp32 biases as one pair.  Same arithmetic as requant_f_rne_x2.
template <bool ACT, bool HI16, bool PRE = false>
__device__ __forceinline__ int2 requant_f_rne_x2_pk(int a0, int a1, unsigned long long fb, const EpiPk &k)
{
    unsigned long long v;
    if (PRE) v = ffma2_pk(pack_f2(__int_as_float(a0), __int_as_float(a1)), k.s_in, k.in_add);
    else v = ffma2_pk(pack_f2(__int2float_rn(a0), __int2float_rn(a1)), k.s_in, k.magic);
    v = fadd2_pk(v, fb);
    float2 t = unpack_f2(v);
    if (ACT) {
        const float2 l = unpack_f2(ffma2_pk(v, k.eighth, k.leak_add));
        t.x = fmaxf(fmaxf(t.x, l.x), YB_MAGIC - 4096.f); t.y = fmaxf(fmaxf(t.y, l.y), YB_MAGIC - 4096.f);
    } else {
        t.x = fmaxf(t.x, YB_MAGIC - 32768.f); t.y = fmaxf(t.y, YB_MAGIC - 32768.f);
    }
    if (HI16) { t.x = fminf(t.x, YB_MAGIC + 32767.f); t.y = fminf(t.y, YB_MAGIC + 32767.f); }
    const float2 o = unpack_f2(ffma2_pk(pack_f2(t.x, t.y), k.s_out, k.out_add));
    return make_int2(unbias_magic(__float_as_int(o.x), k.one), unbias_magic(__float_as_int(o.y), k.one));
}

// Contract P, two channels.  bp = bias << (E - sb).  Returns o + 128 (exact inside [0,255], on the correct side outside).
template <bool ACT>
__device__ __forceinline__ int2 requant_p_x2(int a0, int a1, int bp0, int bp1, const EpiConst &k)
{
    const float2 magic = make_float2(YB_MAGIC, YB_MAGIC);
    const float2 nf = make_float2(__int2float_rn((a0 << k.la) + bp0), __int2float_rn((a1 << k.la) + bp1));
    float2 r = ffma2(nf, make_float2(k.s_in, k.s_in), magic);
    if (ACT) {                                                             // negatives: one RNE shift by sh+3
        const float2 r2 = ffma2(nf, make_float2(k.s_in2, k.s_in2), magic);
        r.x = fmaxf(r.x, r2.x); r.y = fmaxf(r.y, r2.y);
    }
    r.x = fmaxf(r.x, YB_MAGIC - 129.f); r.y = fmaxf(r.y, YB_MAGIC - 129.f);      // keeps the float positive; -129 still reads as saturated
    return make_int2(__float_as_int(r.x) - (YB_MAGIC_BITS - 128), __float_as_int(r.y) - (YB_MAGIC_BITS - 128));
}

// Requantise 4 accumulators with per-channel words b (fp32 bias bits for the F fast paths, int otherwise) and pack them
// into one word.  EPI_F_RNE_NOHI is EPI_F_RNE without the upper 16-bit clamp.
// PRE (F fast paths only): acc[] started at YB_MAGIC_BITS, see requant_f_rne_x2.
template <int EPI, bool ACT, class P, bool PRE = false>
__device__ __forceinline__ unsigned requant4v(const int *acc, int4 b, const P &p, unsigned &ovf, bool count)
{
    static_assert(!PRE || EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI, "pre-biased accumulators: contract F fast paths only");
    if (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) {
        constexpr bool HI = EPI == EPI_F_RNE;
        const int2 lo = requant_f_rne_x2<ACT, HI, PRE>(acc[0], acc[1], make_float2(__int_as_float(b.x), __int_as_float(b.y)), p.k);
        const int2 hi = requant_f_rne_x2<ACT, HI, PRE>(acc[2], acc[3], make_float2(__int_as_float(b.z), __int_as_float(b.w)), p.k);
        return pack_sat_s8(lo.x, lo.y, pack_sat_s8(hi.x, hi.y, 0u));
    } else if (EPI == EPI_P) {
        const int2 lo = requant_p_x2<ACT>(acc[0], acc[1], b.x, b.y, p.k);
        const int2 hi = requant_p_x2<ACT>(acc[2], acc[3], b.z, b.w, p.k);
        if (count && (unsigned)(lo.x | lo.y | hi.x | hi.y) > 255u)          // some value left [0,255] (negatives set the high bits)
            ovf += ((unsigned)lo.x > 255u) + ((unsigned)lo.y > 255u) + ((unsigned)hi.x > 255u) + ((unsigned)hi.y > 255u);
        return pack_sat_u8(lo.x, lo.y, pack_sat_u8(hi.x, hi.y, 0u)) ^ 0x80808080u;
    } else {
        const int bb[4] = { b.x, b.y, b.z, b.w };
        unsigned word = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int o = requant(acc[j], bb[j], p.q);
            unsigned dummy = 0;
            o = store8(o, count ? ovf : dummy);
            word |= (unsigned)(o & 0xff) << (8 * j);
        }
        return word;
    }
}

// `bias` points at the per-channel words in shared memory; c is a multiple of 4.
template <int EPI, bool ACT, class P, bool PRE = false>
__device__ __forceinline__ unsigned requant4(const int *acc, const int *bias, int c, const P &p, unsigned &ovf, bool count)
{
    return requant4v<EPI, ACT, P, PRE>(acc, *reinterpret_cast<const int4 *>(bias + c), p, ovf, count);
}

// 16 accumulators (channels c0..c0+15) -> 16 output bytes
template <int EPI, bool ACT, class P, bool PRE = false>
__device__ __forceinline__ uint4 requant16(const int (&v)[16], const int *bias, int c0, const P &p, unsigned &ovf, bool count)
{
    uint4 w;
    w.x = requant4<EPI, ACT, P, PRE>(&v[0], bias, c0, p, ovf, count);
    w.y = requant4<EPI, ACT, P, PRE>(&v[4], bias, c0 + 4, p, ovf, count);
    w.z = requant4<EPI, ACT, P, PRE>(&v[8], bias, c0 + 8, p, ovf, count);
    w.w = requant4<EPI, ACT, P, PRE>(&v[12], bias, c0 + 12, p, ovf, count);
    return w;
}

// 4 accumulators -> one word; the F fast paths run on the packed constants, every other epilogue on requant4v.
template <int EPI, bool ACT, class P, bool PRE = false>
__device__ __forceinline__ unsigned requant4_pk(const int *acc, const int *bias, int c, const P &p, const EpiPk &k, unsigned &ovf, bool count)
{
    if (EPI == EPI_F_RNE || EPI == EPI_F_RNE_NOHI) {
        constexpr bool HI = EPI == EPI_F_RNE;
        const ulonglong2 fb = *reinterpret_cast<const ulonglong2 *>(bias + c);         // four fp32 biases = two pairs
        const int2 lo = requant_f_rne_x2_pk<ACT, HI, PRE>(acc[0], acc[1], fb.x, k);
        const int2 hi = requant_f_rne_x2_pk<ACT, HI, PRE>(acc[2], acc[3], fb.y, k);
        return pack_sat_s8(lo.x, lo.y, pack_sat_s8(hi.x, hi.y, 0u));
    } else {
        return requant4v<EPI, ACT, P, false>(acc, *reinterpret_cast<const int4 *>(bias + c), p, ovf, count);
    }
}

// host: exact contract-F arithmetic on one value of t (after the 16-bit clamp), used to prove clamps redundant
static inline long long host_shr_rne(long long x, int n)
{
    if (n <= 0) return x;
    long long fl = x >> n, rem = x - (fl << n), half = 1ll << (n - 1);
    if (rem != half) return fl + (rem > half);
    return fl + (fl & 1);
}
static inline long long host_f_tail(long long t, const LayerQ &q)
{
    if (q.activ && t < 0) t = host_shr_rne(t, 3);
    long long o = q.odir ? t * (1ll << q.oofs) : host_shr_rne(t, q.oofs);
    return o < -128 ? -128 : o > 127 ? 127 : o;
}

// Which epilogue may run: the fp32 paths need every non-saturating intermediate to be exactly representable.
static inline int epi_mode_for(const ConvArgs &a, EpiConst *k)
{
    memset(k, 0, sizeof *k);
    k->one = 1;
    const LayerQ &q = a.q;
    if (a.force_generic_epilogue) return EPI_GENERIC;
    if (q.contract == CONTRACT_F && q.round_mode == ROUND_RNE) {
        const long long bmax = a.bias_abs_max;
        const bool in_ok = q.idir ? q.iofs <= 20 : (q.iofs <= 20 && ((32769LL + bmax) << q.iofs) <= (1LL << 24));
        if (!in_ok || bmax >= (1 << 21) || q.oofs > 20 || (q.odir && q.oofs > 8)) return EPI_GENERIC;
        k->s_in = ldexpf(1.0f, q.idir ? q.iofs : -q.iofs);
        k->in_add = (float)((double)YB_MAGIC * (1.0 - (double)k->s_in));     // exact: 3 * 2^22 * (1 - 2^+-iofs), iofs <= 20
        k->leak_add = YB_MAGIC * 0.875f;
        k->s_out = ldexpf(1.0f, q.odir ? q.oofs : -q.oofs);
        k->out_add = (float)((double)YB_MAGIC * (1.0 - (double)k->s_out));
        // the upper 16-bit clamp is redundant when t = 32767 already saturates the output: the tail is monotone
        return host_f_tail(32767, q) == 127 ? EPI_F_RNE_NOHI : EPI_F_RNE;
    }
    if (q.contract == CONTRACT_P) {
        if (q.sh > 13 || q.sh < -8 || q.la > 4 || a.bias_abs_max >= (1 << 28)) return EPI_GENERIC;
        k->s_in = ldexpf(1.0f, -q.sh);
        k->s_in2 = ldexpf(1.0f, -(q.sh + 3));
        k->la = q.la;
        return EPI_P;
    }
    return EPI_GENERIC;
}


}  // namespace yb
