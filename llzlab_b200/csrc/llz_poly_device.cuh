// llz_poly_device.cuh -- device helpers shared by the polyphase kernels (llz_cuda_resample.cu,
// llz_cuda_polybank.cu): stream sample lookup, the reference's finish step, and the reference-order
// recompute used by the exactness guard.
#pragma once

#include <limits.h>

#include "llz_poly_kernels.h"

namespace llz {

// ---- shared device helpers ---------------------------------------------------------------------

// one interleaved PCM sample as the int16 the resampler filters: s16 as is, s24 >> 8, f32 trunc(clamp(x * 2^15)) -- the
// conversions of llz_cuda_pcm.cu (llz_cuda_pcm_deinterleave to planar s16), so the fused path equals de-interleave + run
__device__ __forceinline__ int pcm_load_s16(int fmt, const unsigned char *p)
{
    if (fmt == LLZ_CUDA_PCM_S16) return (int)*reinterpret_cast<const int16_t *>(p);
    if (fmt == LLZ_CUDA_PCM_S24) return (int)(int16_t)((uint32_t)p[1] | ((uint32_t)p[2] << 8));
    float v = *reinterpret_cast<const float *>(p) * 32768.0f;
    v = fminf(fmaxf(v, -32768.0f), 32767.0f);
    return __float2int_rz(v);
}

// base pointer of channel ch of this call's input (planar: its row; interleaved: its first sample)
__device__ __forceinline__ const int16_t *poly_channel_base(const PolyLaunch &a, int ch)
{
    if (!a.x) return nullptr;
    if (a.pcm_frame_bytes > 0) return reinterpret_cast<const int16_t *>(reinterpret_cast<const unsigned char *>(a.x) + (size_t)ch * a.pcm_sample_bytes);
    return a.x + (long long)ch * a.x_stride;
}

__device__ __forceinline__ int poly_sample(const PolyLaunch &a, const int16_t *xc, const int16_t *hc,
                                           long long s)
{
    const long long sp = s - a.in0;
    if (sp >= 0) {
        if (sp >= a.n_in || !xc) return 0;
        if (a.pcm_frame_bytes > 0) return pcm_load_s16(a.pcm_fmt, reinterpret_cast<const unsigned char *>(xc) + sp * a.pcm_frame_bytes);
        return (int)xc[sp];
    }
    if (hc && sp >= -(long long)a.hist_len) return (int)hc[a.hist_len + sp];
    return 0;
}

// Stage a tile's input span in shared memory: raw[off + e] = X(S0 + e) for e < need (X = the stream sample of
// llz_poly_kernels.h), off = the return value (0..7).  The run of samples that lies inside this call's input arrives
// by ONE TMA bulk copy (cp.async.bulk + mbarrier) of whole 16-byte granules -- for an interior tile that is the whole
// span, a few samples over-read on either side -- and only the fringes (history before x[0], zeros beyond the input,
// the last granule of an input whose length is not a multiple of 8) are filled by the threads.  `raw` must be 16-byte
// aligned and hold need + 16 elements.  (The first version filled every tile that touched the history or the end of
// the input sample by sample; a drop-in frame, where every tile does, spent 20-30 us in that loop.)  *bulk tells the
// caller to mbar_wait(bar, 0) after its __syncthreads.
struct PolySpanPlan {
    int off;                   // raw[off + e] = X(S0 + e)
    bool tma;                  // part of the span arrives by a bulk copy
    long long rel_al, lo_al;   // bulk copy: raw + (lo_al - rel_al) <- xc + lo_al
    uint32_t bytes;
    int e_lo, e_hi;            // the bulk run covers elements [e_lo, e_hi); the threads fill the rest of [0, need)
};

__device__ __forceinline__ PolySpanPlan poly_span_plan(const PolyLaunch &a, const int16_t *xc, long long S0, int need)
{
    PolySpanPlan p;
    const long long rel = S0 - a.in0;                          // x index of element 0 (negative: history)
    p.rel_al = rel & ~7LL;
    p.off = (int)(rel - p.rel_al);
    const long long lo = rel > 0 ? rel : 0;
    const long long hi = (rel + need < a.n_in) ? rel + need : a.n_in;
    long long lo_al = lo & ~7LL;                               // >= max(rel_al, 0): inside x, at or after raw[0]
    long long hi_al = (hi + 7) & ~7LL;
    if (hi_al > a.n_in) hi_al = hi & ~7LL;
    p.tma = xc != nullptr && a.pcm_frame_bytes == 0 && (reinterpret_cast<uintptr_t>(xc) & 15u) == 0 && hi_al > lo_al;
    if (!p.tma) lo_al = hi_al = lo;                            // no aligned run: the two fringes meet at lo
    p.lo_al = lo_al;
    p.bytes = (uint32_t)(hi_al - lo_al) * 2u;
    // a span that ends before x[0] or starts beyond the input has no bulk run
    p.e_lo = (int)min(max(lo_al - rel, 0LL), (long long)need);
    p.e_hi = (int)min(max(hi_al - rel, (long long)p.e_lo), (long long)need);
    return p;
}

// one thread, barrier initialised (count 1)
__device__ __forceinline__ void poly_span_issue(const PolySpanPlan &p, const int16_t *xc, int16_t *raw, uint64_t *bar)
{
    mbar_expect_tx(bar, p.bytes);
    tma_bulk_g2s(raw + (p.lo_al - p.rel_al), xc + p.lo_al, p.bytes, bar);
}

template <int NT>
__device__ __forceinline__ void poly_span_fringes(const PolyLaunch &a, const int16_t *xc, const int16_t *hc, long long S0,
                                                  int need, const PolySpanPlan &p, int16_t *raw, int tid)
{
    int16_t *dst = raw + p.off;
#pragma unroll 4
    for (int e = tid; e < p.e_lo; e += NT) dst[e] = (int16_t)poly_sample(a, xc, hc, S0 + e);
#pragma unroll 4
    for (int e = p.e_hi + tid; e < need; e += NT) dst[e] = (int16_t)poly_sample(a, xc, hc, S0 + e);
}

template <int NT>
__device__ __forceinline__ int poly_stage_span(const PolyLaunch &a, const int16_t *xc, const int16_t *hc, long long S0,
                                               int need, int16_t *raw, uint64_t *bar, int tid, bool *bulk)
{
    const PolySpanPlan p = poly_span_plan(a, xc, S0, need);
    if (p.tma && tid == 0) {
        mbar_init(bar, 1);
        poly_span_issue(p, xc, raw, bar);
    }
    poly_span_fringes<NT>(a, xc, hc, S0, need, p, raw, tid);
    *bulk = p.tma;
    return p.off;
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
    if (a.frame_len > 0) frame_end = ((base - a.shift) / a.frame_len + 1) * (long long)a.frame_len;   // of the output's own input frame
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
