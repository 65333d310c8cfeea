/*
 * llz_oracle.h -- CPU restatement of llzlab's FIR / polyphase-resampling hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  libllzfilter_cuda never links or calls it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py) against
 *   (1) the unmodified reference compiled from /root/reference into oracle/_ref/libllzref.so, and
 *   (2) the FNV-1a-64 known-answer values recorded from that build (tests/golden/kat.json).
 * The reference ships no tests or golden vectors of its own, so (1)+(2) are the only pins.
 *
 * Style: whole-signal closed forms (one call = the whole stream), which the reference's
 * frame-by-frame loops are equivalent to; the citations give the reference lines restated.
 */
#ifndef LLZ_ORACLE_H
#define LLZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_HAMMING = 0, ORC_BLACKMAN = 1, ORC_KAISER = 2 };          /* llz_fir.h:26-32 */
enum { ORC_LPF = 0, ORC_HPF = 1, ORC_BPF = 2, ORC_BSF = 3 };

/* windows and tap-count estimators: llz_fir.c:61-83, 85-158, 173-193 */
int    orc_window(double *w, int N, int win);
int    orc_kaiser_beta(double *w, int N, double beta);
double orc_kaiser_atten2beta(double atten);
int    orc_cof_num(int win, double ftrans, double atten);

/* windowed-sinc designs: llz_fir.c:201-393.  *h is malloc'd; returns the (possibly bumped) N */
int    orc_fir_design(double **h, int kind, int N, double fc1, double fc2, int win);

/* y[t] = sum_{i<N} h[i]*x[t-i], i ascending, separate mul and add: llz_fir.c:411-426, 547-584.
 * hist = the N-1 samples before x[0] (NULL = zeros).  n_out may exceed n_in (flush: x beyond
 * n_in reads as 0, llz_fir.c:590-625). */
/* xs: N+1 doubles, ys: M+1 doubles of stream state (zero for a fresh filter) */
void   orc_iir_run(int M, const double *a, int N, const double *b, double *xs, double *ys,
                   const double *x, long long n, double *y);
void   orc_fir_run(const double *h, int N, const double *hist,
                   const double *x, long long n_in, double *y, long long n_out);

/* polyphase plans: llz_resample.c:124-176 (decimate/interp) and 193-255 (L/M) */
typedef struct {
    int L, M;          /* up / down factors                                             */
    int n;             /* prototype length 2*k*phases+1                                   */
    int rows, cols;    /* bank is rows x cols row-major: L x Q (resample), m x K (others) */
    int num_in;        /* frame size in samples                                            */
    int num_out;
    double *h;         /* prototype, n doubles                                             */
    double *bank;      /* rows*cols doubles                                                */
} orc_plan_t;

/* k_override > 0 replaces k = n0/(2L) (extension used by config C4); 0 = reference-derived */
int    orc_resample_plan(orc_plan_t *p, int L, int M, int win, int k_override);   /* :193-255, 367-407 */
int    orc_decimate_plan(orc_plan_t *p, int M, int win);                          /* :124-176, 271-302 */
int    orc_interp_plan(orc_plan_t *p, int L, int win);                            /* :124-176, 320-348 */
void   orc_plan_free(orc_plan_t *p);

/* y[m] = trunc(clamp(gain * sum_{k<Q} x[floor(m*M/L)-k] * g[m%L][k])), m = m0 .. m0+n_out-1.
 * x[i] for i<0 or i>=n_in reads as 0.  llz_resample.c:544-609 */
void   orc_resample_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                        int16_t *y, long long m0, long long n_out);
/* y[i] = trunc(clamp(gain * sum_{m<M} sum_{k<K} x[i*M+m+M*k-n] * p[m][k])): llz_resample.c:425-491 */
void   orc_decimate_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                        int16_t *y, long long n_out);
/* per frame of num_in samples, no history, reads past the frame as zeros (divergence R4):
 * y[(f*num_in+i)*L + (L-1-m)] = trunc(clamp(gain * sum_k x_f[i+k] * p[m][k])): llz_resample.c:494-541 */
void   orc_interp_run(const orc_plan_t *p, double gain, const int16_t *x, long long n_in,
                      int16_t *y);

uint64_t orc_fnv64(const void *data, long long nbytes);

/* synthetic inputs of SURVEY.md section 8(d)/9 */
void   orc_lcg_s16(int16_t *x, long long n, uint32_t seed);
void   orc_lcg_f64(double *x, long long n, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif
