// common.cuh — shared device-side definitions for the fixed-point slim_yolo_v2 kernels (sm_100a).
//
// The per-element requantisation implemented here is this library's OWN implementation of the two
// arithmetic contracts; the CPU oracle under oracle/ is written independently and is
// never included or linked.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yb {

enum { CONTRACT_F = 0, CONTRACT_P = 1 };
enum { ROUND_RNE = 0, ROUND_FLOOR = 1, ROUND_HALF_UP = 2 };

// Epilogue programme of one layer, derived on the host from the exponent tables
// (set_quantize_scale, yolo_forward.c:233-257 for contract F; slim_yolo_v2.py:33-38 for contract P).
struct LayerQ {
    int contract;      // CONTRACT_*
    int round_mode;    // ROUND_* (contract F)
    int activ;         // leaky-ReLU 1/8 on negatives (utils/modules.py:25)
    int pool;          // 2x2/2 max-pool fused after requantisation
    // contract F: t = sh(acc,iofs,idir) + bias_sh[c]; sat16; leaky; o = sat8(sh(t,oofs,odir))
    int iofs, idir, oofs, odir;
    // contract P: num = (acc << la) + bias_sh[c]; o = RNE(num >> (sh + 3*[activ && num<0]))  (sh may be < 0)
    int la, sh;
};

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return min(max(x, lo), hi); }

// Rounding right shift, 1 <= n <= 30, |x| < 2^30.
template <int MODE>
__device__ __forceinline__ int shr_round(int x, int n)
{
    if (MODE == ROUND_FLOOR) return x >> n;
    if (MODE == ROUND_HALF_UP) return (x + (1 << (n - 1))) >> n;
    // RNE: add (half - 1) plus the parity bit of the floor quotient, then floor
    return (x + ((1 << (n - 1)) - 1) + ((x >> n) & 1)) >> n;
}

__device__ __forceinline__ int shr_round_rt(int x, int n, int mode)
{
    if (n <= 0) return x;
    if (n > 30) n = 30;             // |x| < 2^30: every larger shift behaves like 30 (result 0 or -1/0 by mode)
    if (mode == ROUND_FLOOR) return shr_round<ROUND_FLOOR>(x, n);
    if (mode == ROUND_HALF_UP) return shr_round<ROUND_HALF_UP>(x, n);
    return shr_round<ROUND_RNE>(x, n);
}

// Left shift that saturates instead of wrapping (the value is clamped to 16 or 8 bits right after,
// so saturating at +-2^30 is exact).
__device__ __forceinline__ int shl_sat(int x, int n)
{
    if (n <= 0) return x;
    if (n > 30) n = 30;
    int lim = 0x3fffffff >> n;
    return clampi(x, -lim, lim) << n;
}

// Exact integer requantisation of one accumulator (all contracts / rounding modes).
// Contract F returns the final value (already saturated to int8 by the contract).  Contract P returns the rounded value
// UNCLAMPED: the reference never clamps (slim_yolo_v2.py:35); store8() saturates what is finally stored and counts it.
// Both are monotone non-decreasing in acc, so a 2x2 max-pool may be taken before or after.
__device__ __forceinline__ int requant(int acc, int bias_sh, const LayerQ &q)
{
    if (q.contract == CONTRACT_F) {
        int t = (q.idir ? shl_sat(acc, q.iofs) : shr_round_rt(acc, q.iofs, q.round_mode)) + bias_sh;
        t = clampi(t, -32768, 32767);
        if (q.activ && t < 0) t = shr_round_rt(t, 3, q.round_mode);
        int o = q.odir ? shl_sat(t, q.oofs) : shr_round_rt(t, q.oofs, q.round_mode);
        return clampi(o, -128, 127);
    } else {
        int num = shl_sat(acc, q.la) + bias_sh;
        int s = q.sh + ((q.activ && num < 0) ? 3 : 0);
        return s > 0 ? shr_round_rt(num, s, ROUND_RNE) : shl_sat(num, -s);
    }
}

// Saturate a value that is about to be STORED; ovf counts stored elements that had to be clamped (contract P only;
// contract F values are already in range).
__device__ __forceinline__ int store8(int o, unsigned &ovf)
{
    int c = clampi(o, -128, 127);
    ovf += (c != o);
    return c;
}

// ---- fast exact requantisation in fp32 ---------------------------------------------------------------------
// The value MAGIC + k (|k| < 2^22, k integer) is an fp32 number whose unit in the last place is 1, so a fused
// multiply-add that lands in [2^23, 2^24) rounds its exact real result to the nearest integer, ties to even: that IS
// the RNE shift of the contracts.  Every step below is monotone in acc, so values outside the exact range can only
// end at a saturation bound, where they belong.  The host enables these paths only when the exponents keep every
// NON-saturating value exactly representable (see epi_mode_for in conv_umma.cu); otherwise the integer requant() runs.
enum { EPI_GENERIC = 0, EPI_F_RNE = 1, EPI_P = 2, EPI_F_RNE_NOHI = 3 };
#define YB_MAGIC 12582912.0f            /* 1.5 * 2^23 = 0x4B400000: low byte 0, so (bits & 0xff) is the int8 result */

struct EpiConst {
    float s_in;       // F: 2^(-iofs) or 2^(+iofs) ; P: 2^(-sh)
    float in_add;     // F: MAGIC * (1 - s_in): first addend when the accumulators already carry MAGIC (conv_first)
    float s_in2;      // P: 2^(-(sh+3)) (leaky branch)
    float leak_add;   // F: MAGIC * 7/8
    float s_out;      // F: 2^(-oofs) or 2^(+oofs)
    float out_add;    // F: MAGIC * (1 - s_out)
    int la;           // P: left shift of acc
    int one;          // 1, opaque to the compiler: x * one - MAGIC_BITS is an IMAD (FMA pipe) where x - MAGIC_BITS would be an integer add on the
                      // half-rate ALU pipe that the requantisation epilogues saturate (ncu: alu 63 %, fma 15 % on conv2)
};

// Contract F, round-half-even.  fb = (float)sh(bias).  Returns MAGIC + o as float bits.
// The bias is added AFTER the rounding FMA: round-half-even is not invariant under adding an odd integer, so folding
// an odd sh(bias) into the FMA's addend would flip the direction of exact ties.
template <bool ACT>
__device__ __forceinline__ unsigned requant_f_rne(int acc, float fb, const EpiConst &k)
{
    float v = __fmaf_rn(__int2float_rn(acc), k.s_in, YB_MAGIC);           // MAGIC + rne(sh(acc, iofs))   (MAGIC is even)
    v = __fadd_rn(v, fb);                                                 // + sh(b, bofs): exact integer add
    v = fminf(fmaxf(v, YB_MAGIC - 32768.f), YB_MAGIC + 32767.f);          // 16-bit accumulator
    if (ACT) v = fmaxf(v, __fmaf_rn(v, 0.125f, k.leak_add));              // leaky(t) = max(t, rne(t/8))
    v = __fmaf_rn(v, k.s_out, k.out_add);                                 // MAGIC + rne(sh(t, oofs))
    v = fminf(fmaxf(v, YB_MAGIC - 128.f), YB_MAGIC + 127.f);
    return __float_as_uint(v);
}

// Contract P.  bp = bias << (E - sb).  Returns MAGIC + clamp(o) as float bits; counts clamped values in ovf.
template <bool ACT>
__device__ __forceinline__ unsigned requant_p(int acc, int bp, const EpiConst &k, unsigned &ovf, bool count)
{
    float nf = __int2float_rn((acc << k.la) + bp);
    float r = __fmaf_rn(nf, k.s_in, YB_MAGIC);
    if (ACT) r = fmaxf(r, __fmaf_rn(nf, k.s_in2, YB_MAGIC));             // negatives: one RNE shift by sh+3
    float c = fminf(fmaxf(r, YB_MAGIC - 128.f), YB_MAGIC + 127.f);
    ovf += (count && c != r);
    return __float_as_uint(c);
}

// low bytes of four words -> one word
__device__ __forceinline__ unsigned pack_bytes(unsigned a, unsigned b, unsigned c, unsigned d)
{
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__device__ __forceinline__ int dp4a_s8(int a, int b, int c)
{
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// max of 4 packed signed bytes
__device__ __forceinline__ unsigned vmax4(unsigned a, unsigned b) { return __vmaxs4(a, b); }

}  // namespace yb
