// Host emulation of the warp-level overlap-save algorithm of llzlab_b200/csrc/llz_cuda_fir_fft.cu: the same
// dft32 / dft32_twisted / spectrum code (llz_fft32.cuh compiles for the host), lanes run one after another and the
// shared-memory transposes become array transposes.  Checks one work item (two blocks) against the direct sum.
// Usage: fft_emulate <ntaps> <f32:0|1>   -> prints max |err| relative to full scale, exit 0 if within bound.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "llz_fft32.cuh"

using namespace llz;

template <typename T>
static double run(int ntaps)
{
    const int hl = ntaps - 1, B = kFftN - hl;
    std::vector<double> h(ntaps), x(hl + 2 * B + 8);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)(s >> 8) - 8388608) / 8388608.0; };
    double hsum = 0;
    for (auto &v : h) { v = rnd() / ntaps * 4; hsum += fabs(v); }
    for (auto &v : x) v = rnd();
    std::vector<double> Hd(2 * kFftN), twd(2 * kTwistEntries * kFftR);
    fft1024_make_spectrum(h.data(), ntaps, Hd.data());
    fft1024_make_twist_table(twd.data());
    std::vector<T> H(Hd.begin(), Hd.end()), tw(twd.begin(), twd.end());
    struct C2 { T x, y; };
    const C2 *tab = reinterpret_cast<const C2 *>(tw.data());

    static T re[32][32], im[32][32], tr[32][32], ti[32][32];
    // gather: lane t holds z[t + 32 j];  block A starts at input index 0, block B at B
    for (int t = 0; t < 32; ++t)
        for (int j = 0; j < 32; ++j) { re[t][j] = (T)x[t + 32 * j]; im[t][j] = (T)x[B + t + 32 * j]; }
    for (int t = 0; t < 32; ++t) dft32<T, false>(re[t], im[t]);
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) { tr[a][b] = re[b][a]; ti[a][b] = im[b][a]; }
    for (int k2 = 0; k2 < 32; ++k2) {
        dft32_twisted<T, false>(tr[k2], ti[k2], tab + k2, 32);
        for (int k = 0; k < 32; ++k) cmul_inplace<T, false>(tr[k2][k], ti[k2][k], H[2 * (k * 32 + k2)], H[2 * (k * 32 + k2) + 1]);
        dft32<T, true>(tr[k2], ti[k2]);
    }
    for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) { re[a][b] = tr[b][a]; im[a][b] = ti[b][a]; }
    double worst = 0;
    for (int t = 0; t < 32; ++t) {
        dft32_twisted<T, true>(re[t], im[t], tab + t, 32);
        for (int j = 0; j < 32; ++j) {
            const int m = t + 32 * j;
            if (m < hl) continue;
            // direct sums: output time m (block A), m + B (block B), input index == time here
            double ya = 0, yb = 0;
            for (int i = 0; i < ntaps; ++i) { ya += h[i] * (double)(T)x[m - i]; yb += h[i] * (double)(T)x[m + B - i]; }
            worst = fmax(worst, fabs((double)re[t][j] - ya));
            worst = fmax(worst, fabs((double)im[t][j] - yb));
        }
    }
    return worst / hsum;
}

int main(int argc, char **argv)
{
    const int ntaps = argc > 1 ? atoi(argv[1]) : 127;
    const int f32 = argc > 2 ? atoi(argv[2]) : 0;
    const double e = f32 ? run<float>(ntaps) : run<double>(ntaps);
    printf("%.3e\n", e);
    return e < (f32 ? 2e-6 : 1e-14) ? 0 : 1;
}
