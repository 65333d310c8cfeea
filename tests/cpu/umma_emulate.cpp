// Host emulation of the tcgen05 phase-bank kernel (llz_cuda_polybank_umma.cu).  What is exercised without a GPU:
//   * umma_replication: the replicated bank (L' = r L, M' = r M) has M' a multiple of 16, L' >= 64, and addresses the same
//     bank row and input sample for every output as the original L / M (llz_resample.c:586-588);
//   * poly_umma_build_tables: the digit planes are read back with the kernel's own index algebra -- phase tile, chunk,
//     plane, K-major SWIZZLE_128B offset -- against a sample operand laid out as the kernel sees it: byte planes of the
//     input, row j of a tile = the 128-byte windows starting at byte j*M' + w0 + 128*chunk (overlapping-row view);
//   * the digit products are accumulated per weight class in 32-bit integers as the UTCIMMAs do (range checked), combined
//     into the 64-bit sum T, and compared with the long-double dot product: |T 2^-s - sum| <= eps, the bound the guard
//     band is built from; an adversarial input comes close to it;
//   * umma_finish's integer arithmetic (32.32 fixed point: floor, fraction, toward-zero truncation, saturation, near-integer
//     flag) against the reference's finish step in long double (llz_resample.c:594-601) for the same T.
//   umma_emulate L M Q planes gain  ->  prints  max_err/eps  adversarial_err/eps  max|acc|  finish_mismatches  replication
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "llz_umma_tables.h"

using namespace llz;

// the kernel's finish step, host copy of umma_finish (llz_cuda_polybank_umma.cu)
static int finish_int(long long T, int ush, uint32_t thr32, bool *near)
{
    const long long U = ush >= 0 ? (T >> ush) : (T << -ush);
    const int nf = (int)(U >> 32);
    const uint32_t f = (uint32_t)U;
    const bool up = (int)f < 0;
    const int n = nf + (int)up;
    const uint32_t dist = up ? 0u - f : f;
    *near = n != 0 && dist < thr32;
    int t = nf + (int)(nf < 0 && f != 0u);
    return t < -32768 ? -32768 : t > 32767 ? 32767 : t;
}

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    const int L0 = atoi(argv[1]), M0 = atoi(argv[2]), Q = atoi(argv[3]), planes = atoi(argv[4]);
    const double gain = atof(argv[5]);
    const int r = umma_replication(L0, M0);
    const int L = L0 * r;
    const long long M = (long long)M0 * r;
    if (M % 16 != 0 || L < kUPB) { printf("bad replication %d\n", r); return 1; }
    for (long long o = 0; o < 5LL * L + 7; o += 3)                 // same row, same sample
        if ((o % L) % L0 != o % L0 || (o * M) / L != (o * M0) / L0) { printf("replication changes output %lld\n", o); return 1; }

    uint32_t rng = 2468u;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng; };
    std::vector<double> cb0((size_t)L0 * Q, 0.0), cbg((size_t)L * Q);
    for (int l = 0; l < L0; ++l)
        for (int k = 0; k < Q; ++k) {
            if (l == 0 && L0 > 1) { cb0[k] = (k == Q / 2) ? 1.0 - ldexp(1.0, -53) : 0.0; continue; }
            const double t = (k - Q / 2) + (double)l / L0 + (L0 == 1 ? 0.37 : 0.0);   // a decimator's only row is not a delta
            const double w = 0.42 + 0.5 * cos(M_PI * t / (Q / 2 + 1)) + 0.08 * cos(2 * M_PI * t / (Q / 2 + 1));
            cb0[(size_t)l * Q + k] = (t == 0.0 ? 1.0 : sin(M_PI * t) / (M_PI * t)) * w * (0.9 + 0.1 * ((next() >> 8) / 16777216.0));
        }
    for (int l = 0; l < L; ++l)
        for (int k = 0; k < Q; ++k) cbg[(size_t)l * Q + k] = cb0[(size_t)(l % L0) * Q + k] * gain;   // gain folded in, as the shim does

    std::vector<signed char> tab;
    int s = 0;
    double eps = 0.0, qsum = 0.0;
    const int nchunks = poly_umma_build_tables(cbg.data(), L, (int)M, Q, planes, &tab, &s, &eps, &qsum);
    if (nchunks <= 0) { printf("builder refused\n"); return 1; }
    const int b_stage = umma_b_stage(planes);
    const int ext = umma_row_extent(L, (int)M, Q);
    const int ush = s - 32;
    const uint32_t thr32 = (uint32_t)ceil(ldexp(1.001 * eps + ldexp(1.0, -31), 32)) + 2u;
    const int n_tiles = (L + kUPB - 1) / kUPB;
    const int rows = 6;                                            // cycles emulated per tile (spread over the 128)
    double worst = 0.0, adversarial = 0.0;
    long long acc_max = 0, finish_bad = 0;
    // byte planes of an input stream: plane byte e of cycle-row j is at j*M + e; sample index = j*M - (Q-1) + e
    const long long plane_len = 127LL * M + ext + 16;
    std::vector<int16_t> x((size_t)plane_len);
    for (int t = 0; t < n_tiles; t += (n_tiles > 12 ? n_tiles / 12 : 1)) {
        const UmmaPhaseTile pt = umma_phase_tile(L, (int)M, Q, t);
        if (pt.w0 + kUKC * pt.nchunks > ext) { printf("row extent too small\n"); return 1; }
        for (int pass = 0; pass < 2; ++pass) {
            for (auto &v : x) v = (int16_t)(next() >> 16);
            for (int l = 0; l < pt.pbv; l += (pass ? 7 : 1)) {
                const int d = (int)(((long long)(pt.l0 + l) * M) / L) - pt.c_lo;
                for (int jr = 0; jr < rows; ++jr) {
                    const int j = (jr * 23 + l + t) % kUJB;
                    if (pass) {                                    // adversarial: every sample pushes its tap's rounding error the same way
                        for (int k = 0; k < Q; ++k) {
                            const double g = cbg[(size_t)(pt.l0 + l) * Q + k];
                            const double err = g - ldexp((double)llrint(ldexp(g, s)), -s);
                            x[(size_t)(j * M + pt.c_lo + d + (Q - 1) - k)] = err >= 0 ? 32767 : -32768;
                        }
                    }
                    // the MMAs: K steps of 32 bytes over the chunks, A = byte planes of the row's window, B = digit planes
                    long long acc[8] = {0};
                    for (int c = 0; c < pt.nchunks; ++c) {
                        const int ks_n = (pt.ksteps - 4 * c < 4) ? pt.ksteps - 4 * c : 4;
                        for (int kk = 0; kk < 32 * ks_n; ++kk) {
                            const int16_t xv = x[(size_t)(j * M + pt.w0 + kUKC * c + kk)];
                            const int lo = (uint8_t)(xv & 255), hi = (int8_t)(xv >> 8);
                            for (int p = 0; p < planes; ++p) {
                                const int dg = tab[((size_t)t * nchunks + c) * b_stage + (size_t)p * kUBPlane + umma_b_offset(l, kk)];
                                acc[p] += (long long)lo * dg;
                                acc[p + 1] += (long long)hi * dg;
                            }
                        }
                    }
                    long long T = 0;
                    for (int p = 0; p <= planes; ++p) {
                        if (llabs(acc[p]) > acc_max) acc_max = llabs(acc[p]);
                        T += acc[p] << (8 * p);
                    }
                    // reference: sum_k g[l][k] * x[j*M + c_l - k] in long double (taps already carry the gain)
                    long double ref = 0.0L;
                    for (int k = 0; k < Q; ++k)
                        ref += (long double)cbg[(size_t)(pt.l0 + l) * Q + k] * (long double)x[(size_t)(j * M + pt.c_lo + d + (Q - 1) - k)];
                    const double err = (double)fabsl((long double)T * ldexpl(1.0L, -s) - ref);
                    if (pass) { if (err / eps > adversarial) adversarial = err / eps; }
                    else if (err / eps > worst) worst = err / eps;
                    // finish: integer version against the definition, on the value the kernel has (T 2^-s)
                    bool near;
                    const int y = finish_int(T, ush, thr32, &near);
                    const long double v = (long double)T * ldexpl(1.0L, -s);
                    long double tv = truncl(v);
                    const int want = tv < -32768 ? -32768 : tv > 32767 ? 32767 : (int)tv;
                    const long double dist = fabsl(v - roundl(v));
                    const bool near_want = roundl(v) != 0 && dist < (long double)(1.001 * eps);
                    if (y != want && !near) ++finish_bad;           // a flagged output is recomputed by the guard: only unflagged ones must be right
                    if (near_want && !near) ++finish_bad;           // everything inside the band must be flagged
                }
            }
        }
    }
    // the walk's shares: for random weights, band widths and grid sizes the cuts cover every tile of a band exactly once
    for (int trial = 0; trial < 200; ++trial) {
        const int np = 1 + (int)(next() % 90), G = 1 + (int)(next() % 160);
        const long long bw = 1 + (long long)(next() % 70);
        std::vector<int> w(np);
        long long wsum = 0;
        for (auto &v : w) { v = 3000 + (int)(next() % 4000); wsum += v; }
        long long covered = 0;
        UmmaCut prev = umma_locate(w.data(), np, bw, 0);
        if (prev.p != 0 || prev.rb != 0) { printf("walk: first share does not start at tile 0\n"); return 1; }
        for (int c = 0; c < G; ++c) {
            const UmmaCut hi = c + 1 == G ? UmmaCut{np, 0} : umma_locate(w.data(), np, bw, wsum * bw * (c + 1) / G);
            const long long a0 = prev.p * bw + prev.rb, a1 = hi.p * bw + hi.rb;
            if (a1 < a0) { printf("walk: share %d runs backwards\n", c); return 1; }
            covered += a1 - a0;
            prev = hi;
        }
        if (covered != (long long)np * bw) { printf("walk: %lld of %lld tiles covered\n", covered, (long long)np * bw); return 1; }
    }
    printf("%.4f %.4f %lld %lld %d\n", worst, adversarial, acc_max, finish_bad, r);
    return 0;
}
