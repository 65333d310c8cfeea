// llz_poly_device.cuh -- device helpers shared by the polyphase kernels (llz_cuda_resample.cu,
// llz_cuda_polybank.cu): stream sample lookup, the reference's finish step, and the reference-order
// recompute used by the exactness guard.
#pragma once

#include <limits.h>

#include "llz_poly_kernels.h"

namespace llz {

// ---- shared device helpers ---------------------------------------------------------------------

__device__ __forceinline__ int poly_sample(const PolyLaunch &a, const int16_t *xc, const int16_t *hc,
                                           long long s)
{
    const long long sp = s - a.in0;
    if (sp >= 0) return (sp < a.n_in && xc) ? (int)xc[sp] : 0;
    if (hc && sp >= -(long long)a.hist_len) return (int)hc[a.hist_len + sp];
    return 0;
}

// saturate, truncate toward zero: llz_resample.c:596-601.  Truncating first (F2I.TRUNC saturates at the int32
// range) and clamping the integer gives the same result as the reference's clamp-then-cast for every finite v,
// and keeps the FP64 pipe free for the multiply-accumulates.
__device__ __forceinline__ int16_t poly_finish(double v)
{
    const int t = __double2int_rz(v);
    return (int16_t)min(max(t, -32768), 32767);
}

// |v - nearest integer| < thr for a NON-ZERO nearest integer: the outputs whose truncation could differ between
// two FP64 evaluation orders (truncation toward zero is continuous at 0)
__device__ __forceinline__ bool poly_near_nonzero_integer(double v, double thr)
{
    const int n = __double2int_rn(v);                          // saturates: |v| >= 2^31 is far from n
    return n != 0 && fabs(v - (double)n) < thr;
}

// The reference's own loop for one output: order[] walks the taps as llz_resample.c does, products
// and sums rounded separately.  Reads global memory; only the guard and the strict mode call it.
static __device__ __noinline__ double poly_reference_order_sum(const PolyLaunch &a, const int16_t *xc,
                                                        const int16_t *hc, long long o)
{
    const int r = (int)(o % a.L);
    const long long base = (o * a.M) / a.L + a.shift;
    long long frame_end = LLONG_MAX;
    if (a.frame_len > 0) frame_end = ((o / a.L) / a.frame_len + 1) * (long long)a.frame_len;
    const double *row = a.cbank + (long long)r * a.ctaps;
    double acc = 0.0;
    for (int t = 0; t < a.order_len; ++t) {
        const int k = a.order[t];
        const long long s = base - k;
        const double xv = (s < frame_end) ? (double)poly_sample(a, xc, hc, s) : 0.0;
        acc = __dadd_rn(acc, __dmul_rn(xv, row[k]));
    }
    return acc;
}

// turn an accumulated sum into the output sample under the selected mode
template <int MODE, typename TA>
__device__ __forceinline__ int16_t poly_emit(const PolyLaunch &a, const int16_t *xc, const int16_t *hc,
                                             long long o, TA acc, double thr, double single_x)
{
    const int r = (int)(o % a.L);
    if constexpr (MODE == LLZ_CUDA_ACC_F32) {
        double v = (double)acc;
        const int st = a.single_tap[r];
        if (st >= 0) v = __dmul_rn(single_x, a.cbank[(long long)r * a.ctaps + st]);
        return poly_finish(__dmul_rn(v, a.gain));
    } else {
        double v = __dmul_rn((double)acc, a.gain);
        if (MODE == LLZ_CUDA_ACC_F64 && poly_near_nonzero_integer(v, thr) && a.single_tap[r] < 0) {
            v = __dmul_rn(poly_reference_order_sum(a, xc, hc, o), a.gain);
            atomicAdd(a.guard_count, 1ULL);
        }
        return poly_finish(v);
    }
}

}  // namespace llz
