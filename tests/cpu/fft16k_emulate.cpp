// Host emulation of the 16384-point (16 x 1024) overlap-save algorithm of llzlab_b200/csrc/llz_cuda_fir_fft16k.cu:
// the same dft16 / dft32 / folded-twiddle code and tables (llz_fft32.cuh compiles for the host); the 512 (thread,
// half-pass pair) slots of the item run one after another and the exchanges through shared memory and the L2 scratch
// become array permutations.
// Checks one work item (two blocks of 16384 - halo outputs) against the direct sum.
// Usage: fft16k_emulate <ntaps> <f32:0|1>   -> prints max |err| relative to sum|h|, exit 0 if within bound.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "llz_fft32.cuh"

using namespace llz;

template <typename T>
static double run(int ntaps)
{
    const int halo = (ntaps - 1 + 511) / 512 * 512, B = kFft16kN - halo;
    std::vector<double> h(ntaps), x(B + kFft16kN + 8);
    unsigned s = 4242u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)(s >> 8) - 8388608) / 8388608.0; };
    double hsum = 0;
    for (auto &v : h) { v = rnd() / ntaps * 4; hsum += fabs(v); }
    for (auto &v : x) v = rnd();
    std::vector<double> Hd(2 * kFft16kN), t1(2 * kTwistEntries * kFftR), t2(2 * 16 * kTwistEntries * kFftR), t3(2 * 16 * 512);
    fft16k_make_spectrum(h.data(), ntaps, Hd.data());
    fft1024_make_twist_table(t1.data());
    fft16k_make_twist2(t2.data());
    fft16k_make_twist3(t3.data());
    struct C2 { T x, y; };
    std::vector<T> Hv(Hd.begin(), Hd.end()), t1v(t1.begin(), t1.end()), t2v(t2.begin(), t2.end()), t3v(t3.begin(), t3.end());
    const C2 *H = reinterpret_cast<const C2 *>(Hv.data());
    const C2 *tabW = reinterpret_cast<const C2 *>(t1v.data());
    const C2 *tab2 = reinterpret_cast<const C2 *>(t2v.data());
    const C2 *tab3 = reinterpret_cast<const C2 *>(t3v.data());

    static T re[512][32], im[512][32], xr[16][32][32], xi[16][32][32];
    // gather: thread t (cluster-wide) holds z[n], n = t + 512 q + 1024 a at register q*16 + a
    for (int t = 0; t < 512; ++t)
        for (int q = 0; q < 2; ++q)
            for (int a = 0; a < 16; ++a) {
                const int n = t + 512 * q + 1024 * a;
                re[t][q * 16 + a] = (T)x[n];
                im[t][q * 16 + a] = (T)x[B + n];
            }
    // DFT-16 over a, cluster exchange to [b][j][lane]: n_lo = t + 512 q = lane + 32 j with j = warp + 16 q
    for (int t = 0; t < 512; ++t) {
        dft16<T, false, 0>(re[t], im[t]); dft16<T, false, 16>(re[t], im[t]);
        const int w = t >> 5, lane = t & 31;
        for (int q = 0; q < 2; ++q)
            for (int b = 0; b < 16; ++b) { xr[b][w + 16 * q][lane] = re[t][q * 16 + b]; xi[b][w + 16 * q][lane] = im[t][q * 16 + b]; }
    }
    for (int t = 0; t < 512; ++t) {
        const int b = t >> 5, lane = t & 31;
        for (int j = 0; j < 32; ++j) { re[t][j] = xr[b][j][lane]; im[t][j] = xi[b][j][lane]; }
        dft32_twisted<T, false>(re[t], im[t], tabW + 2 * b, 32);
    }
    auto warp_transpose = [&]() {
        for (int b = 0; b < 16; ++b) {
            for (int l = 0; l < 32; ++l) for (int k = 0; k < 32; ++k) { xr[b][l][k] = re[b * 32 + l][k]; xi[b][l][k] = im[b * 32 + l][k]; }
            for (int c = 0; c < 32; ++c) for (int k = 0; k < 32; ++k) { re[b * 32 + c][k] = xr[b][k][c]; im[b * 32 + c][k] = xi[b][k][c]; }
        }
    };
    warp_transpose();
    for (int t = 0; t < 512; ++t) {
        const int b = t >> 5, k2 = t & 31;
        dft32_twisted<T, false>(re[t], im[t], tab2 + b * kTwistEntries * 32 + k2, 32);
        for (int k1 = 0; k1 < 32; ++k1) {
            const C2 hh = H[(b * 32 + k1) * 32 + k2];
            cmul_inplace<T, false>(re[t][k1], im[t][k1], hh.x, hh.y);
        }
        dft32<T, true>(re[t], im[t]);
    }
    warp_transpose();
    for (int t = 0; t < 512; ++t) {
        const int b = t >> 5, lane = t & 31;
        dft32_twisted<T, true>(re[t], im[t], tabW + lane, 32);
        for (int j = 0; j < 32; ++j) { xr[b][j][lane] = re[t][j]; xi[b][j][lane] = im[t][j]; }
    }
    double worst = 0;
    for (int t = 0; t < 512; ++t) {
        const int w = t >> 5, lane = t & 31;
        for (int q = 0; q < 2; ++q)
            for (int b = 0; b < 16; ++b) { re[t][q * 16 + b] = xr[b][w + 16 * q][lane]; im[t][q * 16 + b] = xi[b][w + 16 * q][lane]; }
        C2 e0[8], e1[8];
        for (int e = 0; e < 8; ++e) { e0[e] = tab3[(0 * 8 + e) * 512 + t]; e1[e] = tab3[(1 * 8 + e) * 512 + t]; }
        if (sizeof(T) == 4) {
            dft16_twisted<T, true, 0>(re[t], im[t], e0);
            dft16_twisted<T, true, 16>(re[t], im[t], e1);
        } else {
            // FP64 kernel: the conjugate outer twiddle exp(+2 pi i n_lo / 16384) from the thread's own w0 (the table's
            // (cos, tan) entry of the angle of n_lo mod 256) times a constant rotation, its powers built on the fly
            const int t0 = t & 255;
            const C2 cw = tab3[4 * 512 + t0];
            const T w0r = cw.x, w0i = cw.x * cw.y;
            for (int q = 0; q < 2; ++q) {
                const int rot = t - t0 + 512 * q;                     // n_lo - t0: a multiple of 256
                const T cr = (T)cos(2 * M_PI * rot / kFft16kN), ci = (T)sin(2 * M_PI * rot / kFft16kN);
                const T wr = fma_t<T>(-w0i, ci, w0r * cr), wi = fma_t<T>(w0i, cr, w0r * ci);
                if (q == 0) dft16_powers<T, true, 0>(re[t], im[t], wr, wi); else dft16_powers<T, true, 16>(re[t], im[t], wr, wi);
            }
        }
        for (int q = 0; q < 2; ++q)
            for (int a = 0; a < 16; ++a) {
                const int m = t + 512 * q + 1024 * a;
                if (m < halo || (m & 127)) continue;             // spot check every 128th output
                double ya = 0, yb = 0;
                for (int i = 0; i < ntaps; ++i) { ya += h[i] * (double)(T)x[m - i]; yb += h[i] * (double)(T)x[m + B - i]; }
                worst = fmax(worst, fabs((double)re[t][q * 16 + a] - ya));
                worst = fmax(worst, fabs((double)im[t][q * 16 + a] - yb));
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
