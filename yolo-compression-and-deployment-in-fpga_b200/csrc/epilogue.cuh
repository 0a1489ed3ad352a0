// epilogue.cuh — requantisation epilogue pieces shared by the tcgen05 convolution kernels: P is any parameter struct
// with members `LayerQ q` and `EpiConst k`.
#pragma once
#include <cstring>
#include <math.h>
#include "kernels.h"

namespace yb {

// Requantise 4 accumulators of channels c..c+3 and pack them into one word.  `bias` points at the per-channel
// words in shared memory (fp32 bias for EPI_F_RNE, int otherwise).
template <int EPI, bool ACT, class P>
__device__ __forceinline__ unsigned requant4(const int *acc, const int *bias, int c, const P &p, unsigned &ovf, bool count)
{
    const int4 b = *reinterpret_cast<const int4 *>(bias + c);
    if (EPI == EPI_F_RNE) {
        return pack_bytes(requant_f_rne<ACT>(acc[0], __int_as_float(b.x), p.k), requant_f_rne<ACT>(acc[1], __int_as_float(b.y), p.k),
                          requant_f_rne<ACT>(acc[2], __int_as_float(b.z), p.k), requant_f_rne<ACT>(acc[3], __int_as_float(b.w), p.k));
    } else if (EPI == EPI_P) {
        return pack_bytes(requant_p<ACT>(acc[0], b.x, p.k, ovf, count), requant_p<ACT>(acc[1], b.y, p.k, ovf, count),
                          requant_p<ACT>(acc[2], b.z, p.k, ovf, count), requant_p<ACT>(acc[3], b.w, p.k, ovf, count));
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

// 16 accumulators (channels c0..c0+15) -> 16 output bytes
template <int EPI, bool ACT, class P>
__device__ __forceinline__ uint4 requant16(const int (&v)[16], const int *bias, int c0, const P &p, unsigned &ovf, bool count)
{
    uint4 w;
    w.x = requant4<EPI, ACT, P>(&v[0], bias, c0, p, ovf, count);
    w.y = requant4<EPI, ACT, P>(&v[4], bias, c0 + 4, p, ovf, count);
    w.z = requant4<EPI, ACT, P>(&v[8], bias, c0 + 8, p, ovf, count);
    w.w = requant4<EPI, ACT, P>(&v[12], bias, c0 + 12, p, ovf, count);
    return w;
}

// Which epilogue may run: the fp32 paths need every non-saturating intermediate to be exactly representable.
static inline int epi_mode_for(const ConvArgs &a, EpiConst *k)
{
    memset(k, 0, sizeof *k);
    const LayerQ &q = a.q;
    if (a.force_generic_epilogue) return EPI_GENERIC;
    if (q.contract == CONTRACT_F && q.round_mode == ROUND_RNE) {
        const long long bmax = a.bias_abs_max;
        const bool in_ok = q.idir ? q.iofs <= 20 : (q.iofs <= 20 && ((32769LL + bmax) << q.iofs) <= (1LL << 24));
        if (!in_ok || bmax >= (1 << 21) || q.oofs > 20) return EPI_GENERIC;
        k->s_in = ldexpf(1.0f, q.idir ? q.iofs : -q.iofs);
        k->leak_add = YB_MAGIC * 0.875f;
        k->s_out = ldexpf(1.0f, q.odir ? q.oofs : -q.oofs);
        k->out_add = (float)((double)YB_MAGIC * (1.0 - (double)k->s_out));
        return EPI_F_RNE;
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
