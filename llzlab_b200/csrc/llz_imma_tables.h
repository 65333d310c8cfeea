// llz_imma_tables.h -- tile geometry and host-side tap tables of the integer tensor-core phase-bank kernel
// (llz_cuda_polybank_imma.cu).  Plain C++ (no CUDA): the kernel, the shim and the host emulation test
// (tests/cpu/imma_emulate.cpp) all read the layout from here.
#pragma once

#include <math.h>
#include <stddef.h>

#include <vector>

#if defined(__CUDACC__)
#define LLZ_IMMA_HD __host__ __device__
#else
#define LLZ_IMMA_HD
#endif

namespace llz {

constexpr int kIPB = 64, kIJB = 64;                   // CTA tile: phases x cycles
constexpr int kIKC = 64;                              // k'' per chunk = two IMMA.16832 steps per pipeline round
constexpr int kIPlanesExact = 5, kIPlanesFast = 3;    // signed base-256 digits of a tap: exact mode / fast mode
constexpr int kIPitch = kIKC + 16;                    // bytes per operand row in shared memory (80: conflict-free ldmatrix)
constexpr int kIStages = 4;                           // the producers run up to four chunks ahead of the consumers
constexpr int kIXStage = 2 * kIJB * kIPitch;          // 10,240 bytes: one chunk of X'', low and high byte planes
LLZ_IMMA_HD constexpr int imma_gstage(int planes) { return planes * kIPB * kIPitch; }   // one chunk of G'', all planes (5: 25,600 bytes)
LLZ_IMMA_HD constexpr int imma_stage(int planes) { return imma_gstage(planes) + kIXStage; }
constexpr int imma_tap_bits(int planes) { return 8 * planes - 2; }          // |g * 2^s| < 2^(8P-2): P signed digits hold +-2^(8P-1)

// Host: the bank [L][Q] as int8 digit planes in the kernel's tile layout [phase tile][chunk][plane][64 phases][80 bytes].
// Returns 0 when the bank cannot be split (all taps zero), else the number of chunks per tile; *shift = s, *eps = the
// bound on |sum_k (g - q 2^-s) x| for |x| <= 32768.
inline int poly_imma_build_tables(const double *cb, int L, int M, int Q, int planes, std::vector<signed char> *out, int *shift, double *eps)
{
    if (planes != kIPlanesExact && planes != kIPlanesFast) return 0;
    const int kIGStage = imma_gstage(planes);
    double gmax = 0.0;
    for (size_t i = 0; i < (size_t)L * Q; ++i) gmax = fmax(gmax, fabs(cb[i]));
    if (!(gmax > 0.0) || !isfinite(gmax)) return 0;
    int e2 = 0;
    frexp(gmax, &e2);                                          // gmax < 2^e2
    const int s = imma_tap_bits(planes) - e2;
    const int n_tiles = (L + kIPB - 1) / kIPB;
    int nchunks = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int l0 = t * kIPB, pbv = (L - l0 < kIPB) ? L - l0 : kIPB;
        const int c_lo = (int)(((long long)l0 * M) / L), c_hi = (int)(((long long)(l0 + pbv - 1) * M) / L);
        const int kp = Q + (c_hi - c_lo);
        if ((kp + kIKC - 1) / kIKC > nchunks) nchunks = (kp + kIKC - 1) / kIKC;
    }
    out->assign((size_t)n_tiles * nchunks * kIGStage, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int l0 = t * kIPB, pbv = (L - l0 < kIPB) ? L - l0 : kIPB;
        const int c_lo = (int)(((long long)l0 * M) / L);
        for (int l = 0; l < pbv; ++l) {
            const int d = (int)(((long long)(l0 + l) * M) / L) - c_lo;
            for (int k = 0; k < Q; ++k) {
                long long q = llrint(ldexp(cb[(size_t)(l0 + l) * Q + k], s));
                const int kk = Q - 1 + d - k;                  // reversed, shifted tap index k''
                signed char *dst = out->data() + ((size_t)t * nchunks + kk / kIKC) * kIGStage + (size_t)l * kIPitch + kk % kIKC;
                for (int p = 0; p < planes; ++p) {
                    const int dg = (int)((((q % 256) + 256 + 128) % 256) - 128);   // signed digit in [-128, 127]
                    dst[(size_t)p * kIPB * kIPitch] = (signed char)dg;
                    q = (q - dg) / 256;
                }
                if (q != 0) return 0;                          // cannot happen for |g 2^s| < 2^(8 planes - 2)
            }
        }
    }
    *shift = s;
    // bound on |sum_k (g_k - q_k 2^-s) x_k| for |x| <= 32768: the taps' ACTUAL rounding errors, worst row (about half of
    // the a-priori Q * 2^-(s+1), which halves the first-level guard hits)
    double worst = 0.0;
    for (int l = 0; l < L; ++l) {
        double row = 0.0;
        for (int k = 0; k < Q; ++k) {
            const double g = cb[(size_t)l * Q + k];
            row += fabs(g - ldexp((double)llrint(ldexp(g, s)), -s));
        }
        worst = fmax(worst, row);
    }
    *eps = 32768.0 * worst * (1.0 + 1e-9) + (double)Q * 32768.0 * ldexp(1.0, -(s + 40));
    return nchunks;
}

}  // namespace llz
