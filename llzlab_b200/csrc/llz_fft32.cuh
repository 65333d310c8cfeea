// llz_fft32.cuh -- the register-resident 32-point complex DFT that the overlap-save FIR kernel
// (llz_cuda_fir_fft.cu) is built on, plus the host-side tables of the 1024-point transform.
//
// A 1024-point transform is split 32 x 32 (four-step): every lane of a warp owns 32 points in
// registers, runs one 32-point DFT, multiplies by the inter-pass twiddles, the warp transposes
// the 32 x 32 matrix through shared memory, and every lane runs a second 32-point DFT.  All
// indices below are compile-time constants after unrolling, so `re` / `im` live in registers.
//
// dft32 is a radix-2 decimation-in-time network (bit-reversed gather, then 5 butterfly stages).
// A butterfly with twiddle w = c*(1 -+ i*t) (c = cos, t = tan) costs 6 FMA-pipe instructions
// instead of 8:   p = v.re +- t*v.im,  q = v.im -+ t*v.re,  out = u +- c*(p + i q).
// Instruction count per 32-point DFT: 388 (64 + 64 + 80 + 88 + 92 over the five stages).
//
// The file compiles for the host too (tests/cpu harness emulates the warp lane by lane).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define LLZ_HD __host__ __device__ __forceinline__
#else
#define LLZ_HD inline
#endif

namespace llz {

constexpr int kFftN = 1024;     // transform length of the overlap-save kernel
constexpr int kFftR = 32;       // radix per pass = lanes per warp

LLZ_HD constexpr int brev5(int i)
{
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos(2*pi*m/32) and tan(2*pi*m/32), m = 0..15 (m = 8 is never used through the tangent form)
LLZ_HD constexpr double tw32_cos(int m)
{
    switch (m) {
    case 0: return 1.0;
    case 1: return 0.98078528040323043;
    case 2: return 0.92387953251128674;
    case 3: return 0.83146961230254524;
    case 4: return 0.70710678118654757;
    case 5: return 0.55557023301960218;
    case 6: return 0.38268343236508978;
    case 7: return 0.19509032201612825;
    case 8: return 0.0;
    case 9: return -0.19509032201612825;
    case 10: return -0.38268343236508978;
    case 11: return -0.55557023301960218;
    case 12: return -0.70710678118654757;
    case 13: return -0.83146961230254524;
    case 14: return -0.92387953251128674;
    default: return -0.98078528040323043;
    }
}
LLZ_HD constexpr double tw32_tan(int m)
{
    switch (m) {
    case 0: return 0.0;
    case 1: return 0.19891236737965800;
    case 2: return 0.41421356237309503;
    case 3: return 0.66817863791929888;
    case 4: return 1.0;
    case 5: return 1.4966057626654890;
    case 6: return 2.4142135623730949;
    case 7: return 5.0273394921258481;
    case 8: return 0.0;
    case 9: return -5.0273394921258481;
    case 10: return -2.4142135623730949;
    case 11: return -1.4966057626654890;
    case 12: return -1.0;
    case 13: return -0.66817863791929888;
    case 14: return -0.41421356237309503;
    default: return -0.19891236737965800;
    }
}

template <typename T> LLZ_HD T fma_t(T a, T b, T c);
template <> LLZ_HD float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> LLZ_HD double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// (u, v) -> (u + w*v, u - w*v),  w = exp(-+ 2*pi*i*m/32)  (forward: minus; INV: plus)
template <typename T, bool INV>
LLZ_HD void bfly32(T &ur, T &ui, T &vr, T &vi, int m)
{
    T pr, pi;
    if (m == 0) {
        pr = vr; pi = vi;
        const T ar = ur, ai = ui;
        ur = ar + pr; ui = ai + pi;
        vr = ar - pr; vi = ai - pi;
    } else if (m == 8) {
        // forward w = -i: w*v = (v.im, -v.re);  inverse w = +i: w*v = (-v.im, v.re)
        pr = INV ? -vi : vi;
        pi = INV ? vr : -vr;
        const T ar = ur, ai = ui;
        ur = ar + pr; ui = ai + pi;
        vr = ar - pr; vi = ai - pi;
    } else {
        const T c = (T)tw32_cos(m);
        const T t = (T)(INV ? -tw32_tan(m) : tw32_tan(m));
        // forward: (1 - i t)(vr + i vi) = (vr + t vi) + i (vi - t vr)
        pr = fma_t<T>(t, vi, vr);
        pi = fma_t<T>(-t, vr, vi);
        const T ar = ur, ai = ui;
        ur = fma_t<T>(c, pr, ar);  ui = fma_t<T>(c, pi, ai);
        vr = fma_t<T>(-c, pr, ar); vi = fma_t<T>(-c, pi, ai);
    }
}

template <typename T, bool INV, int LEN>
LLZ_HD void dft32_stage(T (&ar)[32], T (&ai)[32])
{
    constexpr int half = LEN / 2, step = 32 / LEN;
#pragma unroll
    for (int b = 0; b < 32; b += LEN) {
#pragma unroll
        for (int k = 0; k < half; ++k)
            bfly32<T, INV>(ar[b + k], ai[b + k], ar[b + k + half], ai[b + k + half], k * step);
    }
}

// X[k] = sum_j x[j] * exp(-+ 2*pi*i*j*k/32), natural order in, natural order out, in place.
template <typename T, bool INV>
LLZ_HD void dft32(T (&re)[32], T (&im)[32])
{
    T ar[32], ai[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { ar[i] = re[brev5(i)]; ai[i] = im[brev5(i)]; }
    dft32_stage<T, INV, 2>(ar, ai);
    dft32_stage<T, INV, 4>(ar, ai);
    dft32_stage<T, INV, 8>(ar, ai);
    dft32_stage<T, INV, 16>(ar, ai);
    dft32_stage<T, INV, 32>(ar, ai);
#pragma unroll
    for (int i = 0; i < 32; ++i) { re[i] = ar[i]; im[i] = ai[i]; }
}

// ---- DFT-32 with the four-step twiddle folded in ------------------------------------------------------
// X[k] = sum_t (x[t] * w^t) * exp(-+ 2*pi*i*t*k/32),   w = exp(-+ 2*pi*i*l/1024)  (l = the lane's index in the
// other dimension of the 32 x 32 split).  In the decimation-in-time network the factor w^t turns the stage-`len`
// twiddle W_len^k into  T_len[k] = w^(32/len) * W_len^k  = exp(-+ i*theta),  theta = 2*pi*(l + 32 k)/(32 len):
// every butterfly is "general" (6 instructions), but there is no separate twiddle pass (31 complex multiplies)
// and only 16 table entries per lane instead of 31, because T_len[k + len/4] = -+i * T_len[k] (a free swap).
// For len >= 4 and k < len/4, theta < pi/2, so the (cos, tan) form is safe (|cos * tan| <= 1: no error growth);
// the single len = 2 twiddle reaches theta = pi/2 (l = 16) and uses the (cos, sin) form (8 instructions).
// tab: 16 entries for this lane, stride `ts` elements of C = {x, y}:
//   [0] (cos, sin) len 2 | [1] len 4 k0 | [2..3] len 8 k0..1 | [4..7] len 16 k0..3 | [8..15] len 32 k0..7, (cos, tan)
constexpr int kTwistEntries = 16;

template <typename T, bool INV, bool PARTNER>
LLZ_HD void bfly_tan(T &ur, T &ui, T &vr, T &vi, T c, T t)
{
    // forward: T v = c (1 - i t)(vr + i vi) = c (p + i q),  p = vr + t vi,  q = vi - t vr;   inverse: t -> -t
    const T p = fma_t<T>(INV ? -t : t, vi, vr);
    const T q = fma_t<T>(INV ? t : -t, vr, vi);
    const T ar = ur, ai = ui;
    if (!PARTNER) {
        ur = fma_t<T>(c, p, ar);  ui = fma_t<T>(c, q, ai);
        vr = fma_t<T>(-c, p, ar); vi = fma_t<T>(-c, q, ai);
    } else if (!INV) {
        // twiddle -i*T:  c (q - i p)
        ur = fma_t<T>(c, q, ar);  ui = fma_t<T>(-c, p, ai);
        vr = fma_t<T>(-c, q, ar); vi = fma_t<T>(c, p, ai);
    } else {
        // twiddle +i*T:  c (-q + i p)
        ur = fma_t<T>(-c, q, ar); ui = fma_t<T>(c, p, ai);
        vr = fma_t<T>(c, q, ar);  vi = fma_t<T>(-c, p, ai);
    }
}

template <typename T, bool INV, int LEN, int TAB0, typename C>
LLZ_HD void dft32_twisted_stage(T (&ar)[32], T (&ai)[32], const C *tab, int ts)
{
    constexpr int quarter = LEN / 4, half = LEN / 2;
#pragma unroll
    for (int k = 0; k < quarter; ++k) {
        const C w = tab[(TAB0 + k) * ts];
#pragma unroll
        for (int b = 0; b < 32; b += LEN) {
            bfly_tan<T, INV, false>(ar[b + k], ai[b + k], ar[b + k + half], ai[b + k + half], w.x, w.y);
            bfly_tan<T, INV, true>(ar[b + k + quarter], ai[b + k + quarter], ar[b + k + quarter + half],
                                   ai[b + k + quarter + half], w.x, w.y);
        }
    }
}

template <typename T, bool INV, typename C>
LLZ_HD void dft32_twisted(T (&re)[32], T (&im)[32], const C *tab, int ts)
{
    T ar[32], ai[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { ar[i] = re[brev5(i)]; ai[i] = im[brev5(i)]; }
    {
        // len 2: twiddle (c -+ i s) on the odd element
        const C w = tab[0];
        const T c = w.x, s = INV ? -w.y : w.y;
#pragma unroll
        for (int b = 0; b < 32; b += 2) {
            const T ur = ar[b], ui = ai[b], vr = ar[b + 1], vi = ai[b + 1];
            // (c - i s)(vr + i vi) = (c vr + s vi) + i (c vi - s vr)
            ar[b]     = fma_t<T>(s, vi, fma_t<T>(c, vr, ur));
            ai[b]     = fma_t<T>(-s, vr, fma_t<T>(c, vi, ui));
            ar[b + 1] = fma_t<T>(-s, vi, fma_t<T>(-c, vr, ur));
            ai[b + 1] = fma_t<T>(s, vr, fma_t<T>(-c, vi, ui));
        }
    }
    dft32_twisted_stage<T, INV, 4, 1>(ar, ai, tab, ts);
    dft32_twisted_stage<T, INV, 8, 2>(ar, ai, tab, ts);
    dft32_twisted_stage<T, INV, 16, 4>(ar, ai, tab, ts);
    dft32_twisted_stage<T, INV, 32, 8>(ar, ai, tab, ts);
#pragma unroll
    for (int i = 0; i < 32; ++i) { re[i] = ar[i]; im[i] = ai[i]; }
}

// (r + i*s) *= (c + i*d)        [CONJ: *= (c - i*d)]
template <typename T, bool CONJ>
LLZ_HD void cmul_inplace(T &r, T &s, T c, T d)
{
    if (CONJ) d = -d;
    const T nr = fma_t<T>(-s, d, r * c);
    const T ns = fma_t<T>(s, c, r * d);
    r = nr; s = ns;
}

// ---- host-side tables ---------------------------------------------------------------------------
// tab[e][l], e < 16, l < 32: the folded twiddles of dft32_twisted for lane l, interleaved pairs (see above)
inline void fft1024_make_twist_table(double *tab /* 16*32*2 */)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    int e = 0;
    for (int len = 2; len <= 32; len <<= 1) {
        const int nk = len == 2 ? 1 : len / 4;
        for (int k = 0; k < nk; ++k, ++e)
            for (int l = 0; l < kFftR; ++l) {
                const long double th = two_pi * (long double)(l + 32 * k) / (long double)(32 * len);
                tab[2 * (e * kFftR + l)] = (double)cosl(th);
                tab[2 * (e * kFftR + l) + 1] = (double)(len == 2 ? sinl(th) : tanl(th));
            }
    }
}

// H[k1][k2] = (1/1024) * sum_n h[n] * exp(-2*pi*i*n*k/1024),  k = k2 + 32*k1, interleaved (re, im).
// The 1/N of the inverse transform is folded in here.
inline void fft1024_make_spectrum(const double *h, int ntaps, double *H /* 1024*2 */)
{
    long double ct[kFftN], st[kFftN];
    const long double w = -2.0L * 3.14159265358979323846264338327950288L / (long double)kFftN;
    for (int i = 0; i < kFftN; ++i) { ct[i] = cosl(w * i); st[i] = sinl(w * i); }
    for (int k = 0; k < kFftN; ++k) {
        long double sr = 0.0L, si = 0.0L;
        for (int n = 0; n < ntaps; ++n) {
            const int idx = (int)(((long long)n * k) % kFftN);
            sr += (long double)h[n] * ct[idx];
            si += (long double)h[n] * st[idx];
        }
        const int k2 = k % kFftR, k1 = k / kFftR;
        H[2 * (k1 * kFftR + k2)] = (double)(sr / kFftN);
        H[2 * (k1 * kFftR + k2) + 1] = (double)(si / kFftN);
    }
}

}  // namespace llz
