// Host emulation of the CTA-level 8192-point overlap-save algorithm of llzlab_b200/csrc/llz_cuda_fir_fft8k.cu:
// the same dft8 / dft32 / folded-twiddle code and tables (llz_fft32.cuh compiles for the host); the 256 threads run one
// after another and the shared-memory exchanges become array permutations.  Checks one work item (two blocks of
// 8192 - halo outputs) against the direct sum.
// Usage: fft8k_emulate <ntaps> <f32:0|1>   -> prints max |err| relative to sum|h|, exit 0 if within bound.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "llz_fft32.cuh"

using namespace llz;

template <typename T>
static double run(int ntaps)
{
    const int halo = (ntaps - 1 + 255) / 256 * 256, B = kFft8kN - halo;
    std::vector<double> h(ntaps), x(B + kFft8kN + 8);
    unsigned s = 777u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)(s >> 8) - 8388608) / 8388608.0; };
    double hsum = 0;
    for (auto &v : h) { v = rnd() / ntaps * 4; hsum += fabs(v); }
    for (auto &v : x) v = rnd();
    std::vector<double> Hd(2 * kFft8kN), t1(2 * kTwistEntries * kFftR), t2(2 * 8 * kTwistEntries * kFftR), t3(2 * 16 * 256);
    fft8k_make_spectrum(h.data(), ntaps, Hd.data());
    fft1024_make_twist_table(t1.data());
    fft8k_make_twist2(t2.data());
    fft8k_make_twist3(t3.data());
    struct C2 { T x, y; };
    std::vector<T> Hv(Hd.begin(), Hd.end()), t1v(t1.begin(), t1.end()), t2v(t2.begin(), t2.end()), t3v(t3.begin(), t3.end());
    const C2 *H = reinterpret_cast<const C2 *>(Hv.data());
    const C2 *tabW = reinterpret_cast<const C2 *>(t1v.data());
    const C2 *tab2 = reinterpret_cast<const C2 *>(t2v.data());
    const C2 *tab3 = reinterpret_cast<const C2 *>(t3v.data());

    static T re[256][32], im[256][32], xr[8][32][32], xi[8][32][32];
    // gather: thread tid holds z[n], n = tid + 256 q + 1024 a at register q*8 + a
    for (int tid = 0; tid < 256; ++tid)
        for (int q = 0; q < 4; ++q)
            for (int a = 0; a < 8; ++a) {
                const int n = tid + 256 * q + 1024 * a;
                re[tid][q * 8 + a] = (T)x[n];
                im[tid][q * 8 + a] = (T)x[B + n];
            }
    // S1: DFT-8 over a, CTA exchange to [b][j][t]
    for (int tid = 0; tid < 256; ++tid) {
        dft8<T, false, 0>(re[tid], im[tid]); dft8<T, false, 8>(re[tid], im[tid]);
        dft8<T, false, 16>(re[tid], im[tid]); dft8<T, false, 24>(re[tid], im[tid]);
        const int w = tid >> 5, t = tid & 31;
        for (int q = 0; q < 4; ++q)
            for (int b = 0; b < 8; ++b) { xr[b][w + 8 * q][t] = re[tid][q * 8 + b]; xi[b][w + 8 * q][t] = im[tid][q * 8 + b]; }
    }
    for (int tid = 0; tid < 256; ++tid) {
        const int b = tid >> 5, t = tid & 31;
        for (int j = 0; j < 32; ++j) { re[tid][j] = xr[b][j][t]; im[tid][j] = xi[b][j][t]; }
        dft32_twisted<T, false>(re[tid], im[tid], tabW + 4 * b, 32);
    }
    // warp exchange (transpose within each warp)
    auto warp_transpose = [&]() {
        for (int b = 0; b < 8; ++b) {
            for (int t = 0; t < 32; ++t) for (int k = 0; k < 32; ++k) { xr[b][t][k] = re[b * 32 + t][k]; xi[b][t][k] = im[b * 32 + t][k]; }
            for (int c = 0; c < 32; ++c) for (int k = 0; k < 32; ++k) { re[b * 32 + c][k] = xr[b][k][c]; im[b * 32 + c][k] = xi[b][k][c]; }
        }
    };
    warp_transpose();
    for (int tid = 0; tid < 256; ++tid) {
        const int b = tid >> 5, k2 = tid & 31;
        dft32_twisted<T, false>(re[tid], im[tid], tab2 + b * kTwistEntries * 32 + k2, 32);
        for (int k1 = 0; k1 < 32; ++k1) {
            const C2 hh = H[(b * 32 + k1) * 32 + k2];
            cmul_inplace<T, false>(re[tid][k1], im[tid][k1], hh.x, hh.y);
        }
        dft32<T, true>(re[tid], im[tid]);
    }
    warp_transpose();
    for (int tid = 0; tid < 256; ++tid) {
        const int b = tid >> 5, t = tid & 31;
        dft32_twisted<T, true>(re[tid], im[tid], tabW + t, 32);
        for (int j = 0; j < 32; ++j) { xr[b][j][t] = re[tid][j]; xi[b][j][t] = im[tid][j]; }
    }
    double worst = 0;
    for (int tid = 0; tid < 256; ++tid) {
        const int w = tid >> 5, t = tid & 31;
        for (int q = 0; q < 4; ++q)
            for (int b = 0; b < 8; ++b) { re[tid][q * 8 + b] = xr[b][w + 8 * q][t]; im[tid][q * 8 + b] = xi[b][w + 8 * q][t]; }
        auto E = [&](int q, int e) { return tab3[(q * 4 + e) * 256 + tid]; };
        dft8_twisted<T, true, 0>(re[tid], im[tid], E(0, 0), E(0, 1), E(0, 2), E(0, 3));
        dft8_twisted<T, true, 8>(re[tid], im[tid], E(1, 0), E(1, 1), E(1, 2), E(1, 3));
        dft8_twisted<T, true, 16>(re[tid], im[tid], E(2, 0), E(2, 1), E(2, 2), E(2, 3));
        dft8_twisted<T, true, 24>(re[tid], im[tid], E(3, 0), E(3, 1), E(3, 2), E(3, 3));
        for (int q = 0; q < 4; ++q)
            for (int a = 0; a < 8; ++a) {
                const int m = tid + 256 * q + 1024 * a;
                if (m < halo || (m & 63)) continue;              // spot check every 64th output
                double ya = 0, yb = 0;
                for (int i = 0; i < ntaps; ++i) { ya += h[i] * (double)(T)x[m - i]; yb += h[i] * (double)(T)x[m + B - i]; }
                worst = fmax(worst, fabs((double)re[tid][q * 8 + a] - ya));
                worst = fmax(worst, fabs((double)im[tid][q * 8 + a] - yb));
            }
    }
    return worst / hsum;
}

int main(int argc, char **argv)
{
    const int ntaps = argc > 1 ? atoi(argv[1]) : 4095;
    const int f32 = argc > 2 ? atoi(argv[2]) : 0;
    const double e = f32 ? run<float>(ntaps) : run<double>(ntaps);
    printf("%.3e\n", e);
    return e < (f32 ? 2e-6 : 1e-14) ? 0 : 1;
}
