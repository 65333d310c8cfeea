// Host emulation of the 16384-point (16 x 1024) overlap-save algorithm of llzlab_b200/csrc/llz_cuda_fir_fft16k.cu with the
// kernel's own index maps: one CTA of 256 threads, four half-passes of 16 points per thread (n_lo = tid + 256 qq, row
// j = warp + 8 qq of the residue's 32 x 32 slice), residues 0..7 pushed into the eight shared-memory slices and residues
// 8..15 into blocks 0..7 of the scratch, two rounds (warp w transforms residue w + 8 r; round 0's result goes to scratch
// block 8 + w, round 1's stays in the slice), pull with the kernel's block map, last DFT-16 from the folded table (f32)
// or from computed powers of the thread's own root (f64, dft16_powers).  The same dft16 / dft32 / folded-twiddle code and
// tables as the device (llz_fft32.cuh compiles for the host); the threads run one after another and the barriers become
// loop boundaries.  Checks one work item (two blocks of 16384 - halo outputs) against the direct sum.
// Usage: fft16k_emulate <ntaps> <f32:0|1>   -> prints max |err| relative to sum|h|, exit 0 if within bound.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "llz_fft32.cuh"

using namespace llz;

template <typename T>
static double run(int ntaps)
{
    constexpr int WG = 8, KEPT = 16 - WG, NQ = 4, GT = 256;
    const int halo = (ntaps - 1 + 511) / 512 * 512, B = kFft16kN - halo;
    std::vector<double> h(ntaps), x(B + kFft16kN + 8);
    unsigned s = 4242u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)(s >> 8) - 8388608) / 8388608.0; };
    double hsum = 0;
    for (auto &v : h) { v = rnd() / ntaps * 4; hsum += fabs(v); }
    for (auto &v : x) v = rnd();
    // the kernel zeroes the first halo - (N-1) samples of a block (they reach only discarded outputs)
    const int zero_below = halo - (ntaps - 1);
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

    static C2 xbuf[8][32][32], scr[16][32][32];          // shared-memory slices [slice][row j][column]; scratch blocks
    static T re[256][32], im[256][32];

    // ---- gather, DFT-16 over a, push ----------------------------------------------------------------------------
    for (int tid = 0; tid < 256; ++tid) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int qq = 0; qq < NQ; ++qq) {
            T r16[32], i16[32];                          // the half-pass in registers 0..15
            for (int a = 0; a < 16; ++a) {
                const int n = tid + GT * qq + 1024 * a;
                r16[a] = (T)x[n];
                i16[a] = (T)x[B + n];
            }
            if (tid + GT * qq < zero_below) { r16[0] = 0; i16[0] = 0; }
            dft16<T, false, 0>(r16, i16);
            const int row = warp + WG * qq;
            for (int b = 0; b < 16; ++b) {
                const C2 v{r16[b], i16[b]};
                if (b < WG) xbuf[b][row][lane] = v;
                else        scr[b - WG][row][lane] = v;
            }
        }
    }
    // ---- two rounds: warp w transforms residue b = w + 8 r ---------------------------------------------------------
    for (int r = 0; r < 2; ++r) {
        for (int tid = 0; tid < 256; ++tid) {
            const int warp = tid >> 5, lane = tid & 31, b = warp + WG * r;
            for (int j = 0; j < 32; ++j) {
                const C2 v = r == 0 ? xbuf[warp][j][lane] : scr[b - WG][j][lane];
                re[tid][j] = v.x; im[tid][j] = v.y;
            }
            dft32_twisted<T, false>(re[tid], im[tid], tabW + 2 * b, 32);
        }
        auto warp_transpose = [&]() {
            static T tr[32][32], ti[32][32];
            for (int w = 0; w < 8; ++w) {
                for (int l = 0; l < 32; ++l) for (int k = 0; k < 32; ++k) { tr[l][k] = re[w * 32 + l][k]; ti[l][k] = im[w * 32 + l][k]; }
                for (int c = 0; c < 32; ++c) for (int k = 0; k < 32; ++k) { re[w * 32 + c][k] = tr[k][c]; im[w * 32 + c][k] = ti[k][c]; }
            }
        };
        warp_transpose();
        for (int tid = 0; tid < 256; ++tid) {
            const int warp = tid >> 5, k2 = tid & 31, b = warp + WG * r;
            dft32_twisted<T, false>(re[tid], im[tid], tab2 + b * kTwistEntries * 32 + k2, 32);
            for (int k1 = 0; k1 < 32; ++k1) {
                const C2 hh = H[(b * 32 + k1) * 32 + k2];
                cmul_inplace<T, false>(re[tid][k1], im[tid][k1], hh.x, hh.y);
            }
            dft32<T, true>(re[tid], im[tid]);
        }
        warp_transpose();
        for (int tid = 0; tid < 256; ++tid) {
            const int warp = tid >> 5, lane = tid & 31;
            dft32_twisted<T, true>(re[tid], im[tid], tabW + lane, 32);
            for (int j = 0; j < 32; ++j) {
                const C2 v{re[tid][j], im[tid][j]};
                if (r == 0) scr[KEPT + warp][j][lane] = v;          // round 0: to the scratch
                else        xbuf[warp][j][lane] = v;                // last round: stays in the slice
            }
        }
    }
    // ---- pull, DFT-16 over b with the conjugate outer twiddle, compare -------------------------------------------
    double worst = 0;
    for (int tid = 0; tid < 256; ++tid) {
        const int warp = tid >> 5, lane = tid & 31;
        const C2 cw = tab3[4 * 512 + tid];               // (cos, tan) of 2 pi tid / 16384: the thread's own root
        const T w0r = cw.x, w0i = cw.x * cw.y;
        for (int qq = 0; qq < NQ; ++qq) {
            T r16[32], i16[32];
            const int row = warp + WG * qq;
            for (int b = 0; b < 16; ++b) {
                C2 v;
                if (b < WG)        v = scr[KEPT + b][row][lane];
                else if (b < KEPT) v = scr[b - WG][row][lane];
                else               v = xbuf[b - KEPT][row][lane];
                r16[b] = v.x; i16[b] = v.y;
            }
            if (sizeof(T) == 4) {
                const int nl = GT * qq;                  // n_lo = tid + GT qq = t + 512 q
                C2 e[8];
                for (int i = 0; i < 8; ++i) e[i] = tab3[((nl >> 9) * 8 + i) * 512 + (nl & 511) + tid];
                dft16_twisted<T, true, 0>(r16, i16, e);
            } else {
                const T cr = (T)cos(2 * M_PI * (GT * qq) / kFft16kN), ci = (T)sin(2 * M_PI * (GT * qq) / kFft16kN);
                const T wr = fma_t<T>(-w0i, ci, w0r * cr), wi = fma_t<T>(w0i, cr, w0r * ci);
                dft16_powers<T, true, 0>(r16, i16, wr, wi);
            }
            for (int a = 0; a < 16; ++a) {
                const int m = tid + GT * qq + 1024 * a;
                if (m < halo || (m & 127)) continue;     // valid rows only; spot check every 128th output
                double ya = 0, yb = 0;
                for (int i = 0; i < ntaps; ++i) { ya += h[i] * (double)(T)x[m - i]; yb += h[i] * (double)(T)x[m + B - i]; }
                worst = fmax(worst, fabs((double)r16[a] - ya));
                worst = fmax(worst, fabs((double)i16[a] - yb));
            }
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
