/*
 * llz_oracle.c -- CPU restatement of llzlab's FIR / polyphase-resampling hot path.
 *
 * TEST INFRASTRUCTURE (see llz_oracle.h).  Parity PINNED against oracle/_ref (the unmodified
 * reference compiled here) and tests/golden/kat.json.  Citations are /root/reference paths.
 *
 * All arithmetic is IEEE double with separate multiply and add in the reference's order; build
 * with -ffp-contract=off so the compiler cannot fuse them.
 */
#include "llz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ windows --------------- */

/* modified Bessel I0 by power series, stop when the term drops under 1e-16 * sum
 * (llz_fir.c:85-103; EPS at :20-21) */
static double i0_series(double x)
{
    const double half = 0.5 * x;
    double term = 1.0, total = 1.0, sq = 1.0;
    int k = 0;
    while (sq > total * 1E-16) {
        k += 1;
        term = term * (half / k);
        sq = term * term;
        total = total + sq;
    }
    return total;
}

/* llz_fir.c:141-158 (and :121-139 with beta fixed at 8.96): every index computed, no mirroring */
int orc_kaiser_beta(double *w, int N, double beta)
{
    for (int i = 0; i < N; i++) {
        double denom = i0_series(beta);
        double u = (2. * i / (N - 1)) - 1;
        double numer = i0_series(beta * sqrt(1. - u * u));
        w[i] = numer / denom;
    }
    return N;
}

/* llz_fir.c:61-71 (Hamming), :73-83 (Blackman): first half computed, second half mirrored */
int orc_window(double *w, int N, int win)
{
    if (win == ORC_KAISER)
        return orc_kaiser_beta(w, N, 8.96);
    for (int lo = 0, hi = N - 1; lo <= hi; lo++, hi--) {
        double v;
        if (win == ORC_HAMMING)
            v = 0.54 - 0.46 * cos(2 * M_PI * lo / (N - 1));
        else
            v = 0.42 - 0.5 * cos(2 * M_PI * lo / (N - 1)) + 0.08 * cos(4 * M_PI * lo / (N - 1));
        w[lo] = v;
        w[hi] = v;
    }
    return N;
}

/* llz_fir.c:105-118 */
double orc_kaiser_atten2beta(double atten)
{
    if (atten <= 21.)
        return 0.;
    if (atten < 50.)
        return 0.5842 * pow(atten - 21., 0.4) + 0.07886 * (atten - 21.);
    return 0.1102 * (atten - 8.7);
}

/* llz_fir.c:173-193: C truncation of the double quotient */
int orc_cof_num(int win, double ftrans, double atten)
{
    switch (win) {
    case ORC_HAMMING:  return (int)(6.2 / ftrans);
    case ORC_BLACKMAN: return (int)(6.6 / ftrans);
    default:
        if (atten <= 21.)
            return (int)((0.9222 * 2.) / ftrans);
        return (int)(((atten - 7.95) * 2.) / (14.36 * ftrans));
    }
}

/* ------------------------------------------------------------------ designs --------------- */

/* sin(pi x)/(pi x) with exact 0.0 at non-zero integers (llz_fir.c:39-59; :48 is dead code) */
static double sinc_pi(double x)
{
    if (x == 0.0)
        return 1.0;
    if (x == floor(x))
        return 0.0;
    return sin(M_PI * fmod(x, 2.0)) / (M_PI * x);
}

/* llz_fir.c:201-269 (kernels) and :271-393 (allocation wrappers, even-N bump for HPF/BPF/BSF) */
int orc_fir_design(double **h_out, int kind, int N, double fc1, double fc2, int win)
{
    if (kind != ORC_LPF && !(N & 1))
        N += 1;                                            /* :305-307, 337-339, 369-371 */
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    double *h = (double *)malloc(sizeof(double) * (size_t)N);
    orc_window(w, N, win);

    if (kind == ORC_LPF) {
        /* delay is a double here: even N is legal (:206) */
        const double centre = (double)(N - 1) / 2;
        for (int lo = 0, hi = N - 1; lo <= centre; lo++, hi--) {
            h[lo] = fc1 * sinc_pi(fc1 * (lo - centre)) * w[lo];
            h[hi] = h[lo];
        }
    } else {
        const int centre = (N - 1) / 2;                    /* int delay, N odd (:222-223) */
        for (int lo = 0, hi = N - 1; lo <= centre; lo++, hi--) {
            const int d = lo - centre;
            double v;
            if (kind == ORC_HPF)
                v = -fc1 * sinc_pi(fc1 * d) * w[lo];
            else if (kind == ORC_BPF)
                v = (fc2 * sinc_pi(fc2 * d) - fc1 * sinc_pi(fc1 * d)) * w[lo];
            else
                v = -(fc2 * sinc_pi(fc2 * d) - fc1 * sinc_pi(fc1 * d)) * w[lo];
            h[lo] = v;
            h[hi] = v;
        }
        if (kind == ORC_HPF)      h[centre] = 1 - fc1;              /* :229 */
        else if (kind == ORC_BPF) h[centre] = fc2 - fc1;            /* :247 */
        else                      h[centre] = 1 - (fc2 - fc1);      /* :265 */
    }
    free(w);
    *h_out = h;
    return N;
}

/* ------------------------------------------------------------------ FIR ------------------- */

void orc_fir_run(const double *h, int N, const double *hist,
                 const double *x, long long n_in, double *y, long long n_out)
{
    for (long long t = 0; t < n_out; t++) {
        double acc = 0.0;
        for (int i = 0; i < N; i++) {                      /* newest sample first (:419-423) */
            long long s = t - i;
            double xv;
            if (s >= n_in)      xv = 0.0;                  /* flush region (:608-609) */
            else if (s >= 0)    xv = x[s];
            else if (hist)      xv = hist[(N - 1) + s];    /* s in [-(N-1), -1] */
            else                xv = 0.0;
            double prod = h[i] * xv;
            acc = acc + prod;
        }
        y[t] = acc;
    }
}

/* ------------------------------------------------------------------ IIR ------------------- */

/* libllzfilter/llz_iir.c:103-132 as a whole-signal loop: y[n] = sum_{k<=N} b[k] x[n-k] - sum_{1<=k<=M} a[k] y[n-k], the b
 * terms first, k ascending, every product and sum rounded separately (no contraction); xs / ys carry the N / M samples
 * before x[0] / y[0] (oldest first, like the reference's shift buffers) and are updated for the next call */
void orc_iir_run(int M, const double *a, int N, const double *b, double *xs, double *ys,
                 const double *x, long long n, double *y)
{
    for (long long t = 0; t < n; t++) {
        for (int i = 0; i < N; i++) xs[i] = xs[i + 1];               /* :115-117 */
        xs[N] = x ? x[t] : 0.0;                                      /* x == NULL: the flush's zeros (:151) */
        for (int i = 0; i < M; i++) ys[i] = ys[i + 1];               /* :119-120 */
        double acc = 0.0;
        for (int k = 0; k <= N; k++) { double p = b[k] * xs[N - k]; acc = acc + p; }    /* :122-123 */
        for (int k = 1; k <= M; k++) { double p = a[k] * ys[M - k]; acc = acc - p; }    /* :124-125 */
        ys[M] = acc;
        y[t] = acc;
    }
}

/* ------------------------------------------------------------------ plans ----------------- */

static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

/* shared by all three plans: prototype length from the transition-band rule, then the bank
 * bank[r][c] = scale * h[stride_map(r) + c*phases]  (0 past the prototype end) */
static void build_bank(orc_plan_t *p, int phases, double fc, double scale, int win,
                       int k_override, int lm_for_rows /* M for the <lM>_L row map, 0 = identity */)
{
    double ftrans = 0.15 * fc;                                        /* :134 / :204 */
    int n0 = orc_cof_num(win, ftrans, 90);                            /* Kaiser uses 90 dB (:143,:213) */
    int k = k_override > 0 ? k_override : n0 / (2 * phases);          /* :148 / :218 */
    p->n = 2 * k * phases + 1;
    p->rows = phases;
    p->cols = p->n / phases + 1;                                      /* :151 / :222 */
    orc_fir_design(&p->h, ORC_LPF, p->n, fc, 0.0, win);
    p->bank = (double *)calloc((size_t)p->rows * p->cols, sizeof(double));
    if (scale == 0) scale = 1.0;                                      /* :131-132 / :201-202 */
    for (int r = 0; r < p->rows; r++) {
        int first = lm_for_rows ? (r * lm_for_rows) % phases : r;     /* :246 vs :170 */
        for (int c = 0; c < p->cols; c++) {
            int u = c * phases + first;
            p->bank[(size_t)r * p->cols + c] = (u < p->n) ? scale * p->h[u] : 0.0;
        }
    }
}

int orc_resample_plan(orc_plan_t *p, int L, int M, int win, int k_override)
{
    memset(p, 0, sizeof(*p));
    double ratio = ((double)L) / M;
    if (ratio > 16 || (1. / ratio) > 16)                              /* :375-378 */
        return -1;
    p->L = L; p->M = M;
    double fc = (1. / L < 1. / M) ? 1. / L : 1. / M;                  /* :382 */
    build_bank(p, L, fc, (double)L, win, k_override, M);              /* :388 */
    int g = gcd_i(L, M);
    p->num_in = (L * M) / g;                                          /* :394-396 */
    while (p->num_in < 1024) p->num_in *= 2;
    p->num_out = (p->num_in * L) / M;                                 /* :398 */
    return 0;
}

int orc_decimate_plan(orc_plan_t *p, int M, int win)
{
    memset(p, 0, sizeof(*p));
    if (M > 16) return -1;                                            /* :278-279 */
    p->L = 1; p->M = M;
    build_bank(p, M, 1. / M, 1.0, win, 0, 0);                         /* :289 */
    p->num_out = 1024 / M;                                            /* :291-293 */
    p->num_in = p->num_out * M;
    return 0;
}

int orc_interp_plan(orc_plan_t *p, int L, int win)
{
    memset(p, 0, sizeof(*p));
    if (L > 16) return -1;                                            /* :326-327 */
    p->L = L; p->M = 1;
    build_bank(p, L, 1. / L, (double)L, win, 0, 0);                   /* :336 */
    p->num_in = 1024;                                                 /* :338-339 */
    p->num_out = 1024 * L;
    return 0;
}

void orc_plan_free(orc_plan_t *p)
{
    free(p->h); free(p->bank);
    p->h = p->bank = NULL;
}

/* ------------------------------------------------------------------ int16 data paths ------ */

/* gain, saturate, then C double->short conversion = truncation toward zero (:594-601) */
static int16_t finish_s16(double acc, double gain)
{
    acc = acc * gain;
    if (acc > 32767)  acc = 32767;
    if (acc < -32768) acc = -32768;
    return (int16_t)acc;
}

void orc_resample_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                      int16_t *y, long long m0, long long n_out)
{
    const int L = p->L, M = p->M, Q = p->cols;
    for (long long j = 0; j < n_out; j++) {
        long long m = m0 + j;
        long long base = (m * M) / L;                                 /* :586, 64-bit */
        const double *row = p->bank + (size_t)(m % L) * Q;            /* :587 */
        double acc = 0.0;
        for (int k = 0; k < Q; k++) {                                 /* :590-592 */
            long long s = base - k;
            double xv = (s >= 0 && s < n_in) ? (double)x[s] : 0.0;
            double prod = xv * row[k];
            acc = acc + prod;
        }
        y[j] = finish_s16(acc, gain);
    }
}

void orc_decimate_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                      int16_t *y, long long n_out)
{
    const int M = p->M, K = p->cols, n = p->n;
    for (long long i = 0; i < n_out; i++) {
        double acc = 0.0;
        for (int m = 0; m < M; m++) {                                 /* :467-473, m outer k inner */
            const double *row = p->bank + (size_t)m * K;
            for (int k = 0; k < K; k++) {
                long long s = i * M + m + (long long)M * k - n;
                double xv = (s >= 0 && s < n_in) ? (double)x[s] : 0.0;
                double prod = xv * row[k];
                acc = acc + prod;
            }
        }
        y[i] = finish_s16(acc, gain);
    }
}

void orc_interp_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                    int16_t *y)
{
    const int L = p->L, K = p->cols, F = p->num_in;
    for (long long i = 0; i < n_in; i++) {
        long long in_frame = i % F;
        for (int m = 0; m < L; m++) {                                 /* :520-533 */
            const double *row = p->bank + (size_t)m * K;
            double acc = 0.0;
            for (int k = 0; k < K; k++) {
                double xv = (in_frame + k < F && i + k < n_in) ? (double)x[i + k] : 0.0;
                double prod = xv * row[k];
                acc = acc + prod;
            }
            y[i * L + (L - 1 - m)] = finish_s16(acc, gain);
        }
    }
}

/* ------------------------------------------------------------------ helpers --------------- */

uint64_t orc_fnv64(const void *data, long long nbytes)
{
    const unsigned char *b = (const unsigned char *)data;
    uint64_t hsh = 0xcbf29ce484222325ULL;
    for (long long i = 0; i < nbytes; i++) {
        hsh ^= b[i];
        hsh *= 0x100000001b3ULL;
    }
    return hsh;
}

void orc_lcg_s16(int16_t *x, long long n, uint32_t seed)
{
    uint32_t s = seed;
    for (long long i = 0; i < n; i++) {
        s = s * 1664525u + 1013904223u;
        x[i] = (int16_t)((int32_t)(s >> 17) - 16384);
    }
}

void orc_lcg_f64(double *x, long long n, uint32_t seed)
{
    uint32_t s = seed;
    for (long long i = 0; i < n; i++) {
        s = s * 1664525u + 1013904223u;
        x[i] = (double)((int32_t)(s >> 8) - 8388608) / 8388608.0;
    }
}
