/*
 * llz_fir.h -- libllzfilter_cuda: drop-in boundary for llzlab's FIR path.
 *
 * Every prototype, enum value and macro below replaces, symbol for symbol, the declaration in
 * the reference header libllzfilter/llz_fir.h (reference lines cited per entry).  Code written
 * against the reference header compiles and links unchanged against libllzfilter_cuda.
 *
 * Division of labour in this library:
 *   - tap design (windows, estimators, *_cof) runs on the host in C, bit-identical doubles;
 *   - llz_fir_filter / _flush run the convolution on the GPU (sm_100a) and are synchronous:
 *     buf_out is valid when the call returns.  There is no CPU fallback: without a CUDA device
 *     the *_init functions return (unsigned long)-1 and llz_cuda_last_error() says why.
 *
 * Batched / device-resident / multi-GPU entry points are in llz_cuda.h, not here.
 */
#ifndef _LLZ_FIR_H
#define _LLZ_FIR_H

#ifndef M_PI
#define M_PI 3.14159265358979323846            /* reference llz_fir.h:16-18 */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* window selector -- reference llz_fir.h:24-32 (values 0,1,2 are ABI) */
typedef int win_t;
enum {
    HAMMING  = 0,
    BLACKMAN = 1,
    KAISER   = 2,
};

/* ---- streaming filter handles ------------------------------------------- llz_fir.h:38-50 --
 * fc, fc1, fc2 are normalised to fs/2.  flt_len even is bumped to flt_len+1 for everything but
 * the low-pass (llz_fir.c:305-307).  The handle keeps the last flt_len-1 input samples on the
 * device between calls.  Failure: (unsigned long)-1.                                          */
unsigned long llz_fir_filter_lpf_init(int frame_len, int flt_len, double fc, win_t win_type);
unsigned long llz_fir_filter_hpf_init(int frame_len, int flt_len, double fc, win_t win_type);
unsigned long llz_fir_filter_bandpass_init(int frame_len, int flt_len,
                                           double fc1, double fc2, win_t win_type);
unsigned long llz_fir_filter_bandstop_init(int frame_len, int flt_len,
                                           double fc1, double fc2, win_t win_type);
void          llz_fir_filter_uninit(unsigned long handle);

/* ---- data path ---------------------------------------------------------- llz_fir.h:56-58 --
 * llz_fir_filter: y[t] = sum_{i<flt_len} h[i]*x[t-i] for one frame (frame_len <= the init-time
 * frame_len); returns frame_len (llz_fir.c:582).  buf_in and buf_out may alias.
 * llz_fir_filter_flush: pushes zeros and emits the flt_len-1 tail samples; returns flt_len-1
 * (llz_fir.c:624).  Both return -1 on a CUDA error instead of asserting.
 * Bit-identical to the reference for FINITE input.  The kernel pads the taps to a multiple of 16 with zeros, which
 * multiply up to 15 samples older than x[t-flt_len+1] that the reference never reads: an Inf or NaN there turns into
 * NaN here (0 * Inf) in outputs the reference leaves finite.  The tolerance-mode banks (overlap-save) spread a
 * non-finite sample over its whole block.                                                       */
int llz_fir_filter(unsigned long handle, double *buf_in, double *buf_out, int frame_len);
int llz_fir_filter_flush(unsigned long handle, double *buf_out);

/* ---- windows ------------------------------------------------------------ llz_fir.h:64-70 -- */
int    llz_hamming(double *w, const int N);
int    llz_blackman(double *w, const int N);
int    llz_kaiser(double *w, const int N);                       /* beta fixed at 8.96 */
int    llz_kaiser_beta(double *w, const int N, const double beta);
double llz_kaiser_atten2beta(double atten);

/* ---- tap-count estimators (C truncation) -------------------------------- llz_fir.h:76-80 -- */
int llz_hamming_cof_num(double ftrans);
int llz_blackman_cof_num(double ftrans);
int llz_kaiser_cof_num(double ftrans, double atten);

/* ---- windowed-sinc designs ---------------------------------------------- llz_fir.h:87-93 --
 * *h is allocated with malloc(); the CALLER frees it.  Return value: the tap count used.      */
int llz_fir_lpf_cof(double **h, int N, double fc, win_t win_type);
int llz_fir_hpf_cof(double **h, int N, double fc, win_t win_type);
int llz_fir_bandpass_cof(double **h, int N, double fc1, double fc2, win_t win_type);
int llz_fir_bandstop_cof(double **h, int N, double fc1, double fc2, win_t win_type);

/* one output sample: x points at the newest input x[n]; returns sum_i h[i]*x[n-i]
 * (host utility, reference llz_fir.h:95 / llz_fir.c:411-426)                                  */
double llz_conv(const double *x, const double *h, int h_len);

#ifdef __cplusplus
}
#endif

#endif /* _LLZ_FIR_H */
