/*
 * llz_internal.h -- shared between the host C design code and the CUDA shim.  Not installed.
 */
#ifndef LLZ_INTERNAL_H
#define LLZ_INTERNAL_H

#include "llz_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LLZ_KIND_DECIMATE 0      /* the CLI's -t values (example/llz_resample/llz_parseopt.h:31-35) */
#define LLZ_KIND_INTERP   1
#define LLZ_KIND_RESAMPLE 2

/* taps for one of the four designs; returns the tap count actually used, h malloc'd */
int llz_design_taps(double **h, int kind, int N, double fc1, double fc2, win_t win);

/*
 * A polyphase plan in two shapes.
 *
 * (1) reference shape: prototype `proto[n]` and `bank[rows][cols]` exactly as the reference
 *     lays them out (llz_resample.c:167-173 / :238-252) -- used for parity checks.
 *
 * (2) canonical shape, what every kernel consumes:
 *         y[o] = finish( sum_{k<ctaps} cbank[o % L][k] * x[ floor(o*M/L) + shift - k ] )
 *     resample: cbank = bank,                     shift = 0      (llz_resample.c:586-592)
 *     decimate: cbank[0][k] = h[n-k] (k>=1), 0,   shift = 0      (llz_resample.c:467-473)
 *     interp  : cbank[r][k] = bank[L-1-r][K-1-k], shift = K-1    (llz_resample.c:520-531),
 *               samples at or beyond the end of the output's own input frame read as 0.
 *     order[ctaps]: the reference's accumulation order expressed in canonical tap indices
 *     (used by the reference-order recompute).
 */
typedef struct {
    int kind;
    int L, M;
    int n;                 /* prototype length */
    int rows, cols;        /* reference bank shape */
    int num_in, num_out;   /* reference frame */
    double *proto;
    double *bank;

    int crows, ctaps;      /* canonical bank shape: crows = L */
    int shift;
    int frame_len;         /* > 0: frame-local input window (interp) */
    int hist_len;          /* samples of history a stream keeps = ctaps-1-shift (>= 0) */
    double *cbank;         /* crows*ctaps, row-major */
    int *order;            /* ctaps */
    int *single_tap;       /* crows: index of the only non-zero tap of the row, else -1 */
    double abs_row_sum;    /* max_r sum_k |cbank[r][k]| : feeds the guard bound */
} llz_plan_t;

int  llz_plan_build(llz_plan_t *p, int kind, int L, int M, win_t win, int k_override);
void llz_plan_free(llz_plan_t *p);

void llz_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
