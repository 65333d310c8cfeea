/*
 * llz_resample.h -- libllzfilter_cuda: drop-in boundary for llzlab's polyphase resampler.
 *
 * Replaces libllzfilter/llz_resample.h of the reference symbol for symbol (lines cited below):
 * one mono int16 PCM stream per handle, fixed frame size chosen by the handle, samples passed
 * as bytes.  The multiply-accumulate runs on the GPU (sm_100a); results are bit-identical to
 * the reference's double-precision loop, including its truncation toward zero.  Calls are
 * synchronous.  No CPU fallback: without a CUDA device *_init returns (unsigned long)-1.
 *
 * Algorithm: Crochiere & Rabiner, "Interpolation and Decimation of Digital Signals -- A
 * Tutorial Review", Proc. IEEE 69(3), 1981 (cited by the reference at llz_resample.h:11-16).
 */
#ifndef _LLZ_RESAMPLE_H
#define _LLZ_RESAMPLE_H

#include "llz_fir.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference llz_resample.h:32-35 */
#define LLZ_DEFAULT_FRAMELEN 1024               /* minimum samples per input frame            */
#define LLZ_FRAMELEN_MAX     (160*147+8192)     /* caller-side buffer bound used by the CLI   */
#define LLZ_RATIO_MAX        16                 /* 1/16 <= L/M <= 16                          */

/* ---- handles ------------------------------------------------------ llz_resample.h:37-44 ---
 * gain multiplies every output before saturation.  Out-of-range factors return
 * (unsigned long)-1 (llz_resample.c:278-279, 326-327, 376-378).  Any *_uninit accepts any
 * resampler handle (the reference CLI calls llz_resample_filter_uninit on all three kinds).   */
unsigned long llz_decimate_init(int M, double gain, win_t win_type);
void          llz_decimate_uninit(unsigned long handle);

unsigned long llz_interp_init(int L, double gain, win_t win_type);
void          llz_interp_uninit(unsigned long handle);

unsigned long llz_resample_filter_init(int L, int M, double gain, win_t win_type);
void          llz_resample_filter_uninit(unsigned long handle);

/* ---- data path ---------------------------------------------------- llz_resample.h:46-52 ---
 * sample_in_size must equal llz_get_resample_framelen_bytes(handle); *sample_out_size receives
 * the bytes written.  Return 0 on success, -1 on a size mismatch or CUDA error (the reference
 * asserts).  llz_interp treats the K-1 samples the reference reads past the frame as zeros.   */
int llz_get_resample_framelen_bytes(unsigned long handle);
int llz_decimate(unsigned long handle, unsigned char *sample_in, int sample_in_size,
                 unsigned char *sample_out, int *sample_out_size);
int llz_interp(unsigned long handle, unsigned char *sample_in, int sample_in_size,
               unsigned char *sample_out, int *sample_out_size);
int llz_resample(unsigned long handle, unsigned char *sample_in, int sample_in_size,
                 unsigned char *sample_out, int *sample_out_size);

#ifdef __cplusplus
}
#endif

#endif /* _LLZ_RESAMPLE_H */
