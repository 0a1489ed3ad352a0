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

__device__ __forceinline__ int dp4a_s8(int a, int b, int c)
{
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// max of 4 packed signed bytes
__device__ __forceinline__ unsigned vmax4(unsigned a, unsigned b) { return __vmaxs4(a, b); }

}  // namespace yb
