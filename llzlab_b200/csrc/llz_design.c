/*
 * llz_design.c -- host side of libllzfilter_cuda: tap design and polyphase planning, in C.
 *
 * This is the part of the reference that runs once per handle (SURVEY.md section 8a rows
 * a1-a11).  It stays on the host; everything it produces is uploaded to the GPU by the shim.
 * The doubles must be bit-identical to the reference's, so each expression keeps the
 * reference's operand order (cited per function); the code itself is written fresh.
 *
 * Exports the tap-design half of llz_fir.h (windows, estimators, *_cof, llz_conv).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "llz_internal.h"

/* ------------------------------------------------------------------------------------------ */
/* error string (thread-local)                                                                 */
/* ------------------------------------------------------------------------------------------ */
static __thread char g_err[512] = "";

void llz_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

const char *llz_cuda_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------------ */
/* windows: reference llz_fir.c:61-83 (raised-cosine pair), :85-158 (Kaiser)                   */
/* ------------------------------------------------------------------------------------------ */

/* The two cosine windows share one loop: value at i for i <= N-1-i, mirrored onto N-1-i. */
typedef double (*cos_window_fn)(int i, int N);

static double hamming_at(int i, int N)
{
    return 0.54 - 0.46 * cos(2 * M_PI * i / (N - 1));                       /* llz_fir.c:66 */
}

static double blackman_at(int i, int N)
{
    return 0.42 - 0.5 * cos(2 * M_PI * i / (N - 1))
                + 0.08 * cos(4 * M_PI * i / (N - 1));                       /* llz_fir.c:78 */
}

static int fill_mirrored(double *w, int N, cos_window_fn at)
{
    int head = 0, tail = N - 1;
    while (head <= tail) {
        w[head] = at(head, N);
        w[tail] = w[head];
        head++;
        tail--;
    }
    return N;
}

int llz_hamming(double *w, const int N)  { return fill_mirrored(w, N, hamming_at); }
int llz_blackman(double *w, const int N) { return fill_mirrored(w, N, blackman_at); }

/* I0(x) = sum_k ((x/2)^k / k!)^2, truncated when a term falls below 1e-16 of the running sum
 * (llz_fir.c:85-103 with EPS from :20-21) */
static double bessel_i0(double x)
{
    double half_x = 0.5 * x;
    double ratio = 1.0;        /* (x/2)^k / k! */
    double term = 1.0;         /* ratio^2      */
    double sum = 1.0;
    for (int k = 1; term > sum * 1E-16; k++) {
        ratio = ratio * (half_x / k);
        term = ratio * ratio;
        sum = sum + term;
    }
    return sum;
}

/* llz_fir.c:141-158: w[i] = I0(beta*sqrt(1-u^2)) / I0(beta), u = 2i/(N-1) - 1; no mirroring */
int llz_kaiser_beta(double *w, const int N, const double beta)
{
    for (int i = 0; i < N; i++) {
        double i0_beta = bessel_i0(beta);
        double u = (2. * i / (N - 1)) - 1;
        w[i] = bessel_i0(beta * sqrt(1. - u * u)) / i0_beta;
    }
    return N;
}

int llz_kaiser(double *w, const int N)                                      /* llz_fir.c:121-139 */
{
    return llz_kaiser_beta(w, N, 8.96);
}

double llz_kaiser_atten2beta(double atten)                                  /* llz_fir.c:105-118 */
{
    if (atten <= 21.)
        return 0.;
    if (atten < 50.)
        return 0.5842 * pow(atten - 21., 0.4) + 0.07886 * (atten - 21.);
    return 0.1102 * (atten - 8.7);
}

static int make_window(double *w, int N, win_t win)
{
    switch (win) {
    case HAMMING:  return llz_hamming(w, N);
    case BLACKMAN: return llz_blackman(w, N);
    case KAISER:   return llz_kaiser(w, N);
    default:       return -1;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* tap-count estimators: reference llz_fir.c:173-193 (the (int) cast truncates)                */
/* ------------------------------------------------------------------------------------------ */
int llz_hamming_cof_num(double ftrans)  { return (int)(6.2 / ftrans); }
int llz_blackman_cof_num(double ftrans) { return (int)(6.6 / ftrans); }

int llz_kaiser_cof_num(double ftrans, double atten)
{
    if (atten <= 21.)
        return (int)((0.9222 * 2.) / ftrans);
    return (int)(((atten - 7.95) * 2.) / (14.36 * ftrans));
}

/* ------------------------------------------------------------------------------------------ */
/* windowed-sinc designs: reference llz_fir.c:39-59 (sinc), :201-269 (kernels), :271-393       */
/* ------------------------------------------------------------------------------------------ */

/* sin(pi x)/(pi x); exactly 0.0 at non-zero integers, argument reduced with fmod(x,2) */
static double sinc_norm(double x)
{
    if (x == 0.0)
        return 1.0;
    if (x == floor(x))
        return 0.0;
    return sin(M_PI * fmod(x, 2.0)) / (M_PI * x);
}

/* ideal (unwindowed) response at offset d from the centre, per design kind */
static double ideal_tap(int kind, double d, double fc1, double fc2)
{
    switch (kind) {
    case LLZ_CUDA_LPF: return fc1 * sinc_norm(fc1 * d);                                 /* :209 */
    case LLZ_CUDA_HPF: return -fc1 * sinc_norm(fc1 * d);                                /* :226 */
    case LLZ_CUDA_BPF: return fc2 * sinc_norm(fc2 * d) - fc1 * sinc_norm(fc1 * d);      /* :244 */
    default:           return -(fc2 * sinc_norm(fc2 * d) - fc1 * sinc_norm(fc1 * d));   /* :262 */
    }
}

int llz_design_taps(double **h_out, int kind, int N, double fc1, double fc2, win_t win)
{
    if (N < 2 || kind < LLZ_CUDA_LPF || kind > LLZ_CUDA_BSF) {
        llz_set_error("llz_design_taps: bad kind %d or length %d", kind, N);
        return -1;
    }
    if (kind != LLZ_CUDA_LPF && (N % 2) == 0)
        N++;                                   /* only the low-pass tolerates even N (:305-307) */

    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    double *h = (double *)malloc(sizeof(double) * (size_t)N);
    if (!w || !h || make_window(w, N, win) < 0) {
        free(w); free(h);
        llz_set_error("llz_design_taps: allocation failed or bad window %d", (int)win);
        return -1;
    }

    if (kind == LLZ_CUDA_LPF) {
        double mid = (double)(N - 1) / 2;      /* half-sample delay for even N (:206) */
        for (int a = 0, b = N - 1; a <= mid; a++, b--)
            h[b] = h[a] = ideal_tap(kind, a - mid, fc1, fc2) * w[a];
    } else {
        int mid = (N - 1) / 2;
        for (int a = 0, b = N - 1; a <= mid; a++, b--)
            h[b] = h[a] = ideal_tap(kind, (double)(a - mid), fc1, fc2) * w[a];
        /* the centre tap is overwritten without the window (:229, :247, :265) */
        if (kind == LLZ_CUDA_HPF)      h[mid] = 1 - fc1;
        else if (kind == LLZ_CUDA_BPF) h[mid] = fc2 - fc1;
        else                           h[mid] = 1 - (fc2 - fc1);
    }
    free(w);
    *h_out = h;
    return N;
}

int llz_fir_lpf_cof(double **h, int N, double fc, win_t win_type)
{
    return llz_design_taps(h, LLZ_CUDA_LPF, N, fc, 0.0, win_type);
}

int llz_fir_hpf_cof(double **h, int N, double fc, win_t win_type)
{
    return llz_design_taps(h, LLZ_CUDA_HPF, N, fc, 0.0, win_type);
}

int llz_fir_bandpass_cof(double **h, int N, double fc1, double fc2, win_t win_type)
{
    return llz_design_taps(h, LLZ_CUDA_BPF, N, fc1, fc2, win_type);
}

int llz_fir_bandstop_cof(double **h, int N, double fc1, double fc2, win_t win_type)
{
    return llz_design_taps(h, LLZ_CUDA_BSF, N, fc1, fc2, win_type);
}

/* host utility: one output sample, newest input first (reference llz_fir.c:411-426) */
double llz_conv(const double *x, const double *h, int h_len)
{
    double y = 0.0;
    for (int i = 0; i < h_len; i++)
        y += h[i] * x[-i];
    return y;
}

/* ------------------------------------------------------------------------------------------ */
/* polyphase plans: reference llz_resample.c:124-176, 193-255, 271-302, 320-348, 367-407       */
/* ------------------------------------------------------------------------------------------ */
static int gcd_int(int a, int b)
{
    while (b) { int r = a % b; a = b; b = r; }
    return a;
}

static int estimate_proto_len(win_t win, double fc)
{
    double ftrans = 0.15 * fc;                                     /* :134, :204 */
    switch (win) {
    case HAMMING:  return llz_hamming_cof_num(ftrans);
    case BLACKMAN: return llz_blackman_cof_num(ftrans);
    default:       return llz_kaiser_cof_num(ftrans, 90);          /* :143, :213 */
    }
}

void llz_plan_free(llz_plan_t *p)
{
    if (!p) return;
    free(p->proto); free(p->bank); free(p->cbank); free(p->order); free(p->single_tap);
    memset(p, 0, sizeof *p);
}

int llz_plan_build(llz_plan_t *p, int kind, int L, int M, win_t win, int k_override)
{
    memset(p, 0, sizeof *p);
    if (win != HAMMING && win != BLACKMAN && win != KAISER) {
        llz_set_error("unknown window %d", (int)win);
        return -1;
    }
    if (L < 1 || M < 1) {
        llz_set_error("factors must be positive (L=%d M=%d)", L, M);
        return -1;
    }
    /* range checks: llz_resample.c:278, :326, :375-378 */
    if (kind == LLZ_KIND_DECIMATE) {
        L = 1;
        if (M > LLZ_RATIO_MAX) { llz_set_error("decimation factor %d > %d", M, LLZ_RATIO_MAX); return -1; }
    } else if (kind == LLZ_KIND_INTERP) {
        M = 1;
        if (L > LLZ_RATIO_MAX) { llz_set_error("interpolation factor %d > %d", L, LLZ_RATIO_MAX); return -1; }
    } else {
        double ratio = ((double)L) / M;
        if (ratio > LLZ_RATIO_MAX || (1. / ratio) > LLZ_RATIO_MAX) {
            llz_set_error("ratio %d/%d outside [1/%d, %d]", L, M, LLZ_RATIO_MAX, LLZ_RATIO_MAX);
            return -1;
        }
    }
    p->kind = kind; p->L = L; p->M = M;

    /* number of phases, cut-off and bank scale per kind (:283-289, :331-336, :382-388) */
    int phases = (kind == LLZ_KIND_DECIMATE) ? M : L;
    double fc = (kind == LLZ_KIND_DECIMATE) ? 1. / M
              : (kind == LLZ_KIND_INTERP)   ? 1. / L
              : ((1. / L < 1. / M) ? 1. / L : 1. / M);
    double scale = (kind == LLZ_KIND_DECIMATE) ? 1.0 : (double)L;

    int half = k_override > 0 ? k_override : estimate_proto_len(win, fc) / (2 * phases);  /* :148, :218 */
    if (half < 1) { llz_set_error("prototype half-length %d < 1", half); return -1; }
    long long n_ll = 2LL * half * phases + 1;
    if (n_ll > (1 << 24)) { llz_set_error("prototype too long (%lld taps)", n_ll); return -1; }
    p->n = (int)n_ll;
    p->rows = phases;
    p->cols = p->n / phases + 1;                                    /* :151, :222 */
    if (llz_design_taps(&p->proto, LLZ_CUDA_LPF, p->n, fc, 0.0, win) < 0)
        return -1;

    p->bank = (double *)calloc((size_t)p->rows * p->cols, sizeof(double));
    if (!p->bank) { llz_set_error("out of memory"); llz_plan_free(p); return -1; }
    for (int r = 0; r < p->rows; r++) {
        /* first prototype index of row r: <r*M>_L for the L/M bank (:246), r otherwise (:170) */
        int first = (kind == LLZ_KIND_RESAMPLE) ? (r * M) % L : r;
        for (int c = 0; c < p->cols; c++) {
            int u = c * phases + first;
            if (u < p->n)
                p->bank[(size_t)r * p->cols + c] = scale * p->proto[u];
        }
    }

    /* reference frame sizes */
    if (kind == LLZ_KIND_DECIMATE) {
        p->num_out = LLZ_DEFAULT_FRAMELEN / M;                      /* :291-293 */
        p->num_in = p->num_out * M;
    } else if (kind == LLZ_KIND_INTERP) {
        p->num_in = LLZ_DEFAULT_FRAMELEN;                           /* :338-339 */
        p->num_out = LLZ_DEFAULT_FRAMELEN * L;
    } else {
        p->num_in = (L * M) / gcd_int(L, M);                        /* :394-398 */
        while (p->num_in < LLZ_DEFAULT_FRAMELEN) p->num_in *= 2;
        p->num_out = (p->num_in * L) / M;
    }

    /* ---- canonical shape (see llz_internal.h) ---- */
    p->crows = L;
    int order_len;
    if (kind == LLZ_KIND_RESAMPLE) {
        p->ctaps = p->cols; p->shift = 0; p->frame_len = 0;
        order_len = p->ctaps;
    } else if (kind == LLZ_KIND_DECIMATE) {
        p->ctaps = p->n + 1; p->shift = 0; p->frame_len = 0;
        order_len = p->n;
    } else {
        p->ctaps = p->cols; p->shift = p->cols - 1; p->frame_len = p->num_in;
        order_len = p->ctaps;
    }
    p->hist_len = p->ctaps - 1 - p->shift;
    p->cbank = (double *)calloc((size_t)p->crows * p->ctaps, sizeof(double));
    p->order = (int *)calloc((size_t)p->ctaps, sizeof(int));
    p->single_tap = (int *)calloc((size_t)p->crows, sizeof(int));
    if (!p->cbank || !p->order || !p->single_tap) {
        llz_set_error("out of memory"); llz_plan_free(p); return -1;
    }

    if (kind == LLZ_KIND_RESAMPLE) {
        memcpy(p->cbank, p->bank, sizeof(double) * (size_t)p->crows * p->ctaps);
        for (int k = 0; k < p->ctaps; k++) p->order[k] = k;                       /* :590-592 */
    } else if (kind == LLZ_KIND_DECIMATE) {
        /* y[i] = sum_j bank[j%M][j/M] * x[i*M - n + j]  ->  tap index k = n - j  (:467-473) */
        for (int j = 0; j < p->n; j++)
            p->cbank[p->n - j] = p->bank[(size_t)(j % M) * p->cols + j / M];
        int t = 0;
        for (int m = 0; m < M; m++)                    /* reference order: m outer, k inner */
            for (int c = 0; c < p->cols; c++) {
                int j = M * c + m;
                if (j < p->n) p->order[t++] = p->n - j;
            }
        for (; t < p->ctaps; t++) p->order[t] = 0;      /* tap 0 is a structural zero */
    } else {
        /* out[i*L + (L-1-m)] = sum_k bank[m][k] * x[i+k]  (:520-531) */
        for (int r = 0; r < L; r++)
            for (int k = 0; k < p->ctaps; k++)
                p->cbank[(size_t)r * p->ctaps + k] =
                    p->bank[(size_t)(L - 1 - r) * p->cols + (p->cols - 1 - k)];
        for (int k = 0; k < p->ctaps; k++) p->order[k] = p->ctaps - 1 - k;
    }
    (void)order_len;

    p->abs_row_sum = 0.0;
    for (int r = 0; r < p->crows; r++) {
        int nz = 0, where = -1;
        double s = 0.0;
        for (int k = 0; k < p->ctaps; k++) {
            double v = p->cbank[(size_t)r * p->ctaps + k];
            if (v != 0.0) { nz++; where = k; }
            s += fabs(v);
        }
        p->single_tap[r] = (nz == 1) ? where : -1;
        if (s > p->abs_row_sum) p->abs_row_sum = s;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* shard planners (llz_cuda.h); pure integer arithmetic                                        */
/* ------------------------------------------------------------------------------------------ */
static void split_even(long long total, int world, int rank, long long *first, long long *count)
{
    long long base = total / world, extra = total % world;
    *count = base + (rank < extra ? 1 : 0);
    *first = base * rank + (rank < extra ? rank : extra);
}

int llz_cuda_shard_channels(int n_channels, int world, int rank, int *first, int *count)
{
    if (n_channels < 0 || world < 1 || rank < 0 || rank >= world) {
        llz_set_error("shard_channels: bad arguments");
        return -1;
    }
    long long f, c;
    split_even(n_channels, world, rank, &f, &c);
    *first = (int)f; *count = (int)c;
    return 0;
}

int llz_cuda_shard_fir_segments(long long n, int flt_len, int world, int rank,
                                llz_cuda_segment_t *seg)
{
    if (n < 0 || flt_len < 1 || world < 1 || rank < 0 || rank >= world || !seg) {
        llz_set_error("shard_fir_segments: bad arguments");
        return -1;
    }
    split_even(n, world, rank, &seg->in_start, &seg->in_count);
    seg->out_start = seg->in_start;
    seg->out_count = seg->in_count;
    seg->halo = flt_len - 1;
    if (seg->halo > seg->in_start) seg->halo = seg->in_start;   /* before t=0 the signal is zero */
    return 0;
}

int llz_cuda_shard_fir_segments_aligned(long long n, int flt_len, long long granule, int world, int rank,
                                        llz_cuda_segment_t *seg)
{
    if (n < 0 || flt_len < 1 || granule < 1 || world < 1 || rank < 0 || rank >= world || !seg) {
        llz_set_error("shard_fir_segments_aligned: bad arguments");
        return -1;
    }
    long long units = (n + granule - 1) / granule, u0, uc;
    split_even(units, world, rank, &u0, &uc);
    long long start = u0 * granule, end = (u0 + uc) * granule;
    if (start > n) start = n;
    if (end > n) end = n;
    seg->in_start = seg->out_start = start;
    seg->in_count = seg->out_count = end - start;
    seg->halo = flt_len - 1;
    if (seg->halo > seg->in_start) seg->halo = seg->in_start;
    return 0;
}

int llz_cuda_shard_resample_segments(long long n_in, int L, int M, int taps_per_phase,
                                     int frame_in, int world, int rank, llz_cuda_segment_t *seg)
{
    if (n_in < 0 || L < 1 || M < 1 || frame_in < 1 || world < 1 || rank < 0 || rank >= world || !seg) {
        llz_set_error("shard_resample_segments: bad arguments");
        return -1;
    }
    int g = gcd_int(L, M);
    if (frame_in % (M / g) != 0 || n_in % frame_in != 0) {
        llz_set_error("shard_resample_segments: n_in must be whole frames of a multiple of M/gcd");
        return -1;
    }
    long long frames = n_in / frame_in, f0, fc;
    split_even(frames, world, rank, &f0, &fc);
    seg->in_start = f0 * frame_in;
    seg->in_count = fc * frame_in;
    /* frame_in is a multiple of M/g, so frame_in*L/M is an integer multiple of L/g:
     * every segment starts at phase 0 */
    seg->out_start = seg->in_start / (M / g) * (L / g);
    seg->out_count = seg->in_count / (M / g) * (L / g);
    seg->halo = taps_per_phase - 1;
    if (seg->halo > seg->in_start) seg->halo = seg->in_start;
    return 0;
}
