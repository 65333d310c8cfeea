/*
 * llz_iir.h -- direct-form IIR filter, the sibling of llz_fir.h.
 *
 * Same four entry points and semantics as the reference's libllzfilter/llz_iir.h:27-30; every sample is computed on
 * the GPU by libllzfilter_cuda (no CPU fallback: without a CUDA device llz_iir_filter_init returns (unsigned long)-1).
 *
 *     y[n] = sum_{k=0..N} b[k] x[n-k]  -  sum_{k=1..M} a[k] y[n-k]          (llz_iir.c:28-33, :103-132)
 *
 * The recurrence is serial in time, so one handle is one GPU thread; the arithmetic is the reference's, operation by
 * operation (b terms first, k ascending, products and sums rounded separately), and the output doubles are
 * bit-identical for finite input.  Throughput comes from the batched entry points of llz_cuda.h
 * (llz_cuda_iir_bank_*: one thread per channel).
 */
#ifndef _LLZ_IIR_H
#define _LLZ_IIR_H

#ifdef __cplusplus
extern "C" {
#endif

/* llz_iir.h:27 / llz_iir.c:38-68: M poles (a[0..M], a[0] taken as 1), N zeros (b[0..N]; b == NULL means all zero, as
 * in the reference).  M, N in [0, 32].                                                                             */
unsigned long llz_iir_filter_init(int M, double *a, int N, double *b);
/* llz_iir.h:28 / llz_iir.c:71-99 */
void          llz_iir_filter_uninit(unsigned long handle);
/* llz_iir.h:29 / llz_iir.c:135-145: filters frame_len samples, returns frame_len (-1 on a CUDA error); x and y may alias */
int           llz_iir_filter(unsigned long handle, double *x, double *y, int frame_len);
/* llz_iir.h:30 / llz_iir.c:147-156: pushes N zeros, writes the N outputs, returns N */
int           llz_iir_filter_flush(unsigned long handle, double *y);

#ifdef __cplusplus
}
#endif

#endif /* _LLZ_IIR_H */
