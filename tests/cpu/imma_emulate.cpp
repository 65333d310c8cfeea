// Host emulation of the integer tensor-core phase-bank kernel (llz_cuda_polybank_imma.cu): the tap tables built by
// poly_imma_build_tables are read back with the kernel's own index algebra (phase tile, chunk, plane, row, byte), the
// digit products are accumulated in 32-bit integers per weight class exactly as the IMMAs do, combined as the
// epilogue does, and compared with the long-double dot product  sum_k g[l][k] x[j*M + c_l - k]
// (libllzfilter/llz_resample.c:586-592).  Checks: every 32-bit accumulator stays in range, the error never exceeds the
// bound `eps` the guard band is built from, and an adversarial input comes close to it (the bound is not slack).
//   imma_emulate L M Q planes   ->  prints  max_err/eps  adversarial_err/eps  max|acc|
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "llz_imma_tables.h"

using namespace llz;

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const int L = atoi(argv[1]), M = atoi(argv[2]), Q = atoi(argv[3]), planes = atoi(argv[4]);
    // a bank shaped like the reference's: windowed-sinc rows, row 0 a single tap of 1 - 2^-53
    std::vector<double> cb((size_t)L * Q, 0.0);
    uint32_t rng = 12345u;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng; };
    for (int l = 0; l < L; ++l)
        for (int k = 0; k < Q; ++k) {
            if (l == 0) { cb[k] = (k == Q / 2) ? 1.0 - ldexp(1.0, -53) : 0.0; continue; }
            const double t = (k - Q / 2) + (double)l / L;
            const double w = 0.42 + 0.5 * cos(M_PI * t / (Q / 2 + 1)) + 0.08 * cos(2 * M_PI * t / (Q / 2 + 1));
            cb[(size_t)l * Q + k] = (t == 0.0 ? 1.0 : sin(M_PI * t) / (M_PI * t)) * w * (0.9 + 0.1 * ((next() >> 8) / 16777216.0));
        }
    std::vector<signed char> tab;
    int s = 0;
    double eps = 0.0;
    const int nchunks = poly_imma_build_tables(cb.data(), L, M, Q, planes, &tab, &s, &eps);
    if (nchunks <= 0) { printf("builder refused\n"); return 1; }
    const int gstage = imma_gstage(planes);
    const int n_tiles = (L + kIPB - 1) / kIPB;
    double worst = 0.0, adversarial = 0.0;
    long long acc_max = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const int l0 = t * kIPB, pbv = (L - l0 < kIPB) ? L - l0 : kIPB;
        const int c_lo = (int)(((long long)l0 * M) / L), c_hi = (int)(((long long)(l0 + pbv - 1) * M) / L);
        const int cspan = c_hi - c_lo, KP = Q + cspan;
        const int my_chunks = (KP + kIKC - 1) / kIKC;
        const int rawn = (kIJB - 1) * M + cspan + Q;
        std::vector<int16_t> span((size_t)rawn + kIKC + 64);
        for (auto &v : span) v = (int16_t)(next() >> 16);          // beyond rawn: garbage that must meet zero taps only
        for (int pass = 0; pass < 2; ++pass) {
            for (int l = 0; l < pbv; ++l) {
                const int d = (int)(((long long)(l0 + l) * M) / L) - c_lo;
                const int j = (l * 13 + t) % kIJB;
                if (pass) {                                        // adversarial: every sample pushes its tap's rounding error the same way
                    for (int k = 0; k < Q; ++k) {
                        const double g = cb[(size_t)(l0 + l) * Q + k];
                        const double err = g - ldexp((double)llrint(ldexp(g, s)), -s);
                        span[(size_t)j * M + Q - 1 + d - k] = err >= 0 ? 32767 : -32768;
                    }
                }
                long long acc[8] = {0};
                for (int c = 0; c < my_chunks; ++c)
                    for (int b = 0; b < kIKC; ++b) {
                        const int kk = c * kIKC + b;
                        const int16_t x = span[(size_t)j * M + kk];
                        const int xl = x & 255, xh = x >> 8;       // x = 256 xh + xl
                        for (int p = 0; p < planes; ++p) {
                            const int dg = tab[((size_t)t * nchunks + c) * gstage + (size_t)p * kIPB * kIPitch + (size_t)l * kIPitch + b];
                            acc[p] += (long long)dg * xl;
                            acc[p + 1] += (long long)dg * xh;
                        }
                    }
                __int128 S = 0;
                for (int dd = planes; dd >= 0; --dd) {
                    if (llabs(acc[dd]) > acc_max) acc_max = llabs(acc[dd]);
                    S = S * 256 + acc[dd];
                }
                const long double got = (long double)S * ldexpl(1.0L, -s);
                long double want = 0.0L;
                for (int k = 0; k < Q; ++k) want += (long double)cb[(size_t)(l0 + l) * Q + k] * span[(size_t)j * M + Q - 1 + d - k];
                const double e = (double)fabsl(got - want) / eps;
                if (pass) { if (e > adversarial) adversarial = e; }
                if (e > worst) worst = e;
            }
        }
    }
    printf("%.6f %.6f %lld\n", worst, adversarial, acc_max);
    return (worst <= 1.0 && acc_max < 2147483647LL) ? 0 : 1;
}
