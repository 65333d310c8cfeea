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
    if (m == 0) {
        const T pr = vr, pi = vi;
        const T ar = ur, ai = ui;
        ur = ar + pr; ui = ai + pi;
        vr = ar - pr; vi = ai - pi;
    } else if (m == 8) {
        // forward w = -i: w*v = (v.im, -v.re);  inverse w = +i: w*v = (-v.im, v.re)   (no negation of data: add / sub)
        const T ar = ur, ai = ui, br = vr, bi = vi;
        if (!INV) { ur = ar + bi; ui = ai - br; vr = ar - bi; vi = ai + br; }
        else      { ur = ar - bi; ui = ai + br; vr = ar + bi; vi = ai - br; }
    } else {
        const T c = (T)tw32_cos(m);
        const T t = (T)(INV ? -tw32_tan(m) : tw32_tan(m));
        // forward: (1 - i t)(vr + i vi) = (vr + t vi) + i (vi - t vr)
        const T pr = fma_t<T>(t, vi, vr);
        const T pi = fma_t<T>(-t, vr, vi);
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
    // only the table value d is negated (once), never the data
    const T dp = CONJ ? -d : d, dn = CONJ ? d : -d;
    const T nr = fma_t<T>(s, dn, r * c);
    const T ns = fma_t<T>(s, c, r * dp);
    r = nr; s = ns;
}

// ---- 8-point transforms (the outer radix of the 8192-point transform, llz_cuda_fir_fft8k.cu) -------------
LLZ_HD constexpr int brev3(int i) { return ((i & 1) << 2) | (i & 2) | ((i & 4) >> 2); }

// X[k] = sum_a x[a] * exp(-+ 2*pi*i*a*k/8) on the strided slice v[O + a], a < 8; natural order in and out
template <typename T, bool INV, int O>
LLZ_HD void dft8(T (&re)[32], T (&im)[32])
{
    T ar[8], ai[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ar[i] = re[O + brev3(i)]; ai[i] = im[O + brev3(i)]; }
#pragma unroll
    for (int b = 0; b < 8; b += 2) bfly32<T, INV>(ar[b], ai[b], ar[b + 1], ai[b + 1], 0);
#pragma unroll
    for (int b = 0; b < 8; b += 4) {
        bfly32<T, INV>(ar[b], ai[b], ar[b + 2], ai[b + 2], 0);
        bfly32<T, INV>(ar[b + 1], ai[b + 1], ar[b + 3], ai[b + 3], 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) bfly32<T, INV>(ar[k], ai[k], ar[k + 4], ai[k + 4], 4 * k);
#pragma unroll
    for (int i = 0; i < 8; ++i) { re[O + i] = ar[i]; im[O + i] = ai[i]; }
}

// X[k] = sum_a (x[a] * w^a) * exp(-+ 2*pi*i*a*k/8) with the twiddle folded in as in dft32_twisted:
// e0 = (cos, sin) of the len-2 twiddle w^4, e1 = (cos, tan) of w^2, e2 / e3 = (cos, tan) of w and w*W_8
template <typename T, bool INV, int O, typename C>
LLZ_HD void dft8_twisted(T (&re)[32], T (&im)[32], C e0, C e1, C e2, C e3)
{
    T ar[8], ai[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ar[i] = re[O + brev3(i)]; ai[i] = im[O + brev3(i)]; }
    {
        const T c = e0.x, s = INV ? -e0.y : e0.y;
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
            const T ur = ar[b], ui = ai[b], vr = ar[b + 1], vi = ai[b + 1];
            ar[b]     = fma_t<T>(s, vi, fma_t<T>(c, vr, ur));
            ai[b]     = fma_t<T>(-s, vr, fma_t<T>(c, vi, ui));
            ar[b + 1] = fma_t<T>(-s, vi, fma_t<T>(-c, vr, ur));
            ai[b + 1] = fma_t<T>(s, vr, fma_t<T>(-c, vi, ui));
        }
    }
#pragma unroll
    for (int b = 0; b < 8; b += 4) {
        bfly_tan<T, INV, false>(ar[b], ai[b], ar[b + 2], ai[b + 2], e1.x, e1.y);
        bfly_tan<T, INV, true>(ar[b + 1], ai[b + 1], ar[b + 3], ai[b + 3], e1.x, e1.y);
    }
    bfly_tan<T, INV, false>(ar[0], ai[0], ar[4], ai[4], e2.x, e2.y);
    bfly_tan<T, INV, false>(ar[1], ai[1], ar[5], ai[5], e3.x, e3.y);
    bfly_tan<T, INV, true>(ar[2], ai[2], ar[6], ai[6], e2.x, e2.y);
    bfly_tan<T, INV, true>(ar[3], ai[3], ar[7], ai[7], e3.x, e3.y);
#pragma unroll
    for (int i = 0; i < 8; ++i) { re[O + i] = ar[i]; im[O + i] = ai[i]; }
}

// ---- 16-point transforms (the outer radix of the 16384-point transform, llz_cuda_fir_fft16k.cu) ----------
LLZ_HD constexpr int brev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

// X[k] = sum_a x[a] * exp(-+ 2*pi*i*a*k/16) on the slice v[O + a], a < 16; natural order in and out
template <typename T, bool INV, int O>
LLZ_HD void dft16(T (&re)[32], T (&im)[32])
{
    T ar[16], ai[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { ar[i] = re[O + brev4(i)]; ai[i] = im[O + brev4(i)]; }
#pragma unroll
    for (int b = 0; b < 16; b += 2) bfly32<T, INV>(ar[b], ai[b], ar[b + 1], ai[b + 1], 0);
#pragma unroll
    for (int b = 0; b < 16; b += 4) {
        bfly32<T, INV>(ar[b], ai[b], ar[b + 2], ai[b + 2], 0);
        bfly32<T, INV>(ar[b + 1], ai[b + 1], ar[b + 3], ai[b + 3], 8);
    }
#pragma unroll
    for (int b = 0; b < 16; b += 8)
#pragma unroll
        for (int k = 0; k < 4; ++k) bfly32<T, INV>(ar[b + k], ai[b + k], ar[b + k + 4], ai[b + k + 4], 4 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) bfly32<T, INV>(ar[k], ai[k], ar[k + 8], ai[k + 8], 2 * k);
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[O + i] = ar[i]; im[O + i] = ai[i]; }
}

// X[k] = sum_a (x[a] * w^a) * exp(-+ 2*pi*i*a*k/16), twiddle folded in as in dft32_twisted.  e[0] = (cos, sin) of the
// len-2 twiddle w^8; e[1] = (cos, tan) of w^4; e[2..3] of w^2 * W_8^{0,1}; e[4..7] of w * W_16^{0..3}
template <typename T, bool INV, int O, typename C>
LLZ_HD void dft16_twisted(T (&re)[32], T (&im)[32], const C (&e)[8])
{
    T ar[16], ai[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { ar[i] = re[O + brev4(i)]; ai[i] = im[O + brev4(i)]; }
    {
        const T c = e[0].x, s = INV ? -e[0].y : e[0].y;
#pragma unroll
        for (int b = 0; b < 16; b += 2) {
            const T ur = ar[b], ui = ai[b], vr = ar[b + 1], vi = ai[b + 1];
            ar[b]     = fma_t<T>(s, vi, fma_t<T>(c, vr, ur));
            ai[b]     = fma_t<T>(-s, vr, fma_t<T>(c, vi, ui));
            ar[b + 1] = fma_t<T>(-s, vi, fma_t<T>(-c, vr, ur));
            ai[b + 1] = fma_t<T>(s, vr, fma_t<T>(-c, vi, ui));
        }
    }
#pragma unroll
    for (int b = 0; b < 16; b += 4) {
        bfly_tan<T, INV, false>(ar[b], ai[b], ar[b + 2], ai[b + 2], e[1].x, e[1].y);
        bfly_tan<T, INV, true>(ar[b + 1], ai[b + 1], ar[b + 3], ai[b + 3], e[1].x, e[1].y);
    }
#pragma unroll
    for (int b = 0; b < 16; b += 8)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            bfly_tan<T, INV, false>(ar[b + k], ai[b + k], ar[b + k + 4], ai[b + k + 4], e[2 + k].x, e[2 + k].y);
            bfly_tan<T, INV, true>(ar[b + k + 2], ai[b + k + 2], ar[b + k + 6], ai[b + k + 6], e[2 + k].x, e[2 + k].y);
        }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bfly_tan<T, INV, false>(ar[k], ai[k], ar[k + 8], ai[k + 8], e[4 + k].x, e[4 + k].y);
        bfly_tan<T, INV, true>(ar[k + 4], ai[k + 4], ar[k + 12], ai[k + 12], e[4 + k].x, e[4 + k].y);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { re[O + i] = ar[i]; im[O + i] = ai[i]; }
}

// X[k] = sum_a (x[a] * w^a) * exp(-+ 2*pi*i*a*k/16) with the powers of w = wr + i*wi (|w| = 1) built on the fly
// instead of read from a folded table: 14 complex products for w^2 .. w^15 (every power a product of two powers of
// depth <= 3: no long error chain), 15 to apply them, then the plain transform -- 264 instead of 208 FMA-pipe
// instructions, and 16 bytes of table per transform (w) instead of 128.  Used where the table traffic costs more than
// the arithmetic: the last pass of the FP64 16384-point kernel is bound by its L2 reads (llz_cuda_fir_fft16k.cu).
template <typename T, bool INV, int O>
LLZ_HD void dft16_powers(T (&re)[32], T (&im)[32], T wr, T wi)
{
    T pr[9], pi[9];                                      // w^1 .. w^8
    pr[1] = wr; pi[1] = wi;
    auto mul = [](T ar, T ai, T br, T bi, T &cr, T &ci) {
        cr = fma_t<T>(-ai, bi, ar * br);
        ci = fma_t<T>(ai, br, ar * bi);
    };
    mul(pr[1], pi[1], pr[1], pi[1], pr[2], pi[2]);
    mul(pr[2], pi[2], pr[1], pi[1], pr[3], pi[3]);
    mul(pr[2], pi[2], pr[2], pi[2], pr[4], pi[4]);
    mul(pr[4], pi[4], pr[1], pi[1], pr[5], pi[5]);
    mul(pr[4], pi[4], pr[2], pi[2], pr[6], pi[6]);
    mul(pr[4], pi[4], pr[3], pi[3], pr[7], pi[7]);
    mul(pr[4], pi[4], pr[4], pi[4], pr[8], pi[8]);
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        T qr, qi;
        if (k <= 8) { qr = pr[k]; qi = pi[k]; }
        else mul(pr[8], pi[8], pr[k - 8], pi[k - 8], qr, qi);
        const T xr = re[O + k], xi = im[O + k];
        re[O + k] = fma_t<T>(-xi, qi, xr * qr);
        im[O + k] = fma_t<T>(xi, qr, xr * qi);
    }
    dft16<T, INV, O>(re, im);
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

// ---- tables of the 8192-point transform (8 x 1024, llz_cuda_fir_fft8k.cu) -----------------------------------
constexpr int kFft8kN = 8192;

// tab2[b][e][k2]: folded twiddles of the second DFT-32 of warp b (base exp(-2*pi*i*(8*k2 + b)/8192))
inline void fft8k_make_twist2(double *tab /* 8*16*32*2 */)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int b = 0; b < 8; ++b) {
        int e = 0;
        for (int len = 2; len <= 32; len <<= 1) {
            const int nk = len == 2 ? 1 : len / 4;
            for (int k = 0; k < nk; ++k, ++e)
                for (int l = 0; l < kFftR; ++l) {
                    const long double th = two_pi * (long double)(8 * l + b + 256 * k) / (long double)(256 * len);
                    double *o = tab + 2 * ((b * kTwistEntries + e) * kFftR + l);
                    o[0] = (double)cosl(th);
                    o[1] = (double)(len == 2 ? sinl(th) : tanl(th));
                }
        }
    }
}

// tab3[q][e][tid], q < 4, e < 4, tid < 256: folded twiddles of the last 8-point transform for n_lo = tid + 256 q
// (base exp(-2*pi*i*n_lo/8192)); e as in dft8_twisted
inline void fft8k_make_twist3(double *tab /* 4*4*256*2 */)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int q = 0; q < 4; ++q)
        for (int tid = 0; tid < 256; ++tid) {
            const long double base = (long double)(tid + 256 * q) / (long double)kFft8kN;     // turns
            const long double th[4] = {two_pi * base * 4, two_pi * base * 2, two_pi * base, two_pi * (base + 0.125L)};
            for (int e = 0; e < 4; ++e) {
                double *o = tab + 2 * ((q * 4 + e) * 256 + tid);
                o[0] = (double)cosl(th[e]);
                o[1] = (double)(e == 0 ? sinl(th[e]) : tanl(th[e]));
            }
        }
}

// H[b][k1][k2] = (1/8192) * sum_n h[n] * exp(-2*pi*i*n*k/8192),  k = b + 8*(k2 + 32*k1)
inline void fft8k_make_spectrum(const double *h, int ntaps, double *H /* 8192*2 */)
{
    long double *ct = new long double[kFft8kN], *st = new long double[kFft8kN];
    const long double w = -2.0L * 3.14159265358979323846264338327950288L / (long double)kFft8kN;
    for (int i = 0; i < kFft8kN; ++i) { ct[i] = cosl(w * i); st[i] = sinl(w * i); }
    for (int k = 0; k < kFft8kN; ++k) {
        long double sr = 0.0L, si = 0.0L;
        for (int n = 0; n < ntaps; ++n) {
            const int idx = (int)(((long long)n * k) % kFft8kN);
            sr += (long double)h[n] * ct[idx];
            si += (long double)h[n] * st[idx];
        }
        const int b = k % 8, khi = k / 8, k2 = khi % kFftR, k1 = khi / kFftR;
        H[2 * ((b * kFftR + k1) * kFftR + k2)] = (double)(sr / kFft8kN);
        H[2 * ((b * kFftR + k1) * kFftR + k2) + 1] = (double)(si / kFft8kN);
    }
    delete[] ct;
    delete[] st;
}

// ---- tables of the 16384-point transform (16 x 1024 across a cluster of two CTAs, llz_cuda_fir_fft16k.cu) ------
constexpr int kFft16kN = 16384;

// tab2[b][e][k2], b < 16: folded twiddles of the second DFT-32 of residue b (base exp(-2*pi*i*(16*k2 + b)/16384))
inline void fft16k_make_twist2(double *tab /* 16*16*32*2 */)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int b = 0; b < 16; ++b) {
        int e = 0;
        for (int len = 2; len <= 32; len <<= 1) {
            const int nk = len == 2 ? 1 : len / 4;
            for (int k = 0; k < nk; ++k, ++e)
                for (int l = 0; l < kFftR; ++l) {
                    const long double th = two_pi * (long double)(16 * l + b + 512 * k) / (long double)(512 * len);
                    double *o = tab + 2 * ((b * kTwistEntries + e) * kFftR + l);
                    o[0] = (double)cosl(th);
                    o[1] = (double)(len == 2 ? sinl(th) : tanl(th));
                }
        }
    }
}

// tab3[q][e][t], q < 2, e < 8, t < 512: folded twiddles of the last 16-point transform for n_lo = t + 512 q
// (base exp(-2*pi*i*n_lo/16384)); e as in dft16_twisted
inline void fft16k_make_twist3(double *tab /* 2*8*512*2 */)
{
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int q = 0; q < 2; ++q)
        for (int t = 0; t < 512; ++t) {
            const long double base = (long double)(t + 512 * q) / (long double)kFft16kN;        // turns
            const long double th[8] = {base * 8, base * 4, base * 2, base * 2 + 0.125L,
                                       base, base + 0.0625L, base + 0.125L, base + 0.1875L};
            for (int e = 0; e < 8; ++e) {
                double *o = tab + 2 * ((q * 8 + e) * 512 + t);
                o[0] = (double)cosl(two_pi * th[e]);
                o[1] = (double)(e == 0 ? sinl(two_pi * th[e]) : tanl(two_pi * th[e]));
            }
        }
}

// H[b][k1][k2] = (1/16384) * sum_n h[n] * exp(-2*pi*i*n*k/16384),  k = b + 16*(k2 + 32*k1)
inline void fft16k_make_spectrum(const double *h, int ntaps, double *H /* 16384*2 */)
{
    long double *ct = new long double[kFft16kN], *st = new long double[kFft16kN];
    const long double w = -2.0L * 3.14159265358979323846264338327950288L / (long double)kFft16kN;
    for (int i = 0; i < kFft16kN; ++i) { ct[i] = cosl(w * i); st[i] = sinl(w * i); }
    for (int k = 0; k < kFft16kN; ++k) {
        long double sr = 0.0L, si = 0.0L;
        for (int n = 0; n < ntaps; ++n) {
            const int idx = (int)(((long long)n * k) % kFft16kN);
            sr += (long double)h[n] * ct[idx];
            si += (long double)h[n] * st[idx];
        }
        const int b = k % 16, khi = k / 16, k2 = khi % kFftR, k1 = khi / kFftR;
        H[2 * ((b * kFftR + k1) * kFftR + k2)] = (double)(sr / kFft16kN);
        H[2 * ((b * kFftR + k1) * kFftR + k2) + 1] = (double)(si / kFft16kN);
    }
    delete[] ct;
    delete[] st;
}

}  // namespace llz
