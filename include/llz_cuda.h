/*
 * llz_cuda.h -- libllzfilter_cuda extensions (NOT in the reference).
 *
 * The reference API (llz_fir.h / llz_resample.h) is one mono stream per handle with host
 * buffers.  The entry points here are what a B200 deployment needs on top of that, kept in a
 * separate header so the two drop-in headers stay identical to the reference's:
 *
 *   1. banks: n_channels independent mono streams behind one handle, planar layout
 *      (channel c = base + c*stride elements, each channel contiguous) -- the same thing as
 *      n_channels reference handles driven in lock-step;
 *   2. device-resident runs on caller-owned device pointers and a caller-owned cudaStream_t,
 *      asynchronous (so throughput can be timed on the device without PCIe);
 *   3. whole-signal calls: one call of any length is equivalent to looping the reference's
 *      frames (closed forms: llz_fir.c:547-584, llz_resample.c:544-609);
 *   4. arithmetic selectors (exact FP64 / reference-order FP64 / fast FP32);
 *   5. explicit prototype length for the resampler (k_override);
 *   6. host-side shard planners for multi-GPU runs (channel shards, time segments with halo).
 *
 * All functions return 0 / a positive count on success and -1 on failure unless stated;
 * llz_cuda_last_error() returns the message of the calling thread's last failure.
 * A handle is bound to the CUDA device that was current when it was created.  Calls on one handle may use different
 * streams from call to call: the library orders a call after the previous call's work on the handle's stream state.
 * Inputs must be finite: the bit-identical modes reproduce the reference for finite samples (a NaN / Inf sample
 * spreads over up to 15 more outputs of the strict FIR kernel and over a whole block of the overlap-save kernels).
 */
#ifndef _LLZ_CUDA_H
#define _LLZ_CUDA_H

#include <stddef.h>

#include "llz_fir.h"
#include "llz_resample.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *llz_cuda_stream_t;            /* a cudaStream_t; NULL = the legacy default stream */

/* ---- runtime --------------------------------------------------------------------------- */
const char *llz_cuda_last_error(void);
int         llz_cuda_device_count(void);    /* 0 when no usable CUDA device */
const char *llz_cuda_build_info(void);      /* "libllzfilter_cuda <ver> sm_100a ..." */
/* Measurement knobs (process-wide; not for production code paths).  The environment is read once, when the first
 * handle is created; this call overrides a value afterwards.  Keys: "pipe_slot_mib" (staging-slot size of the
 * *_run_host pipelines, default 64), "slide_ru" (force a tile variant of the decimating kernel: 11, 7, 5, 3; 0 = auto),
 * "fft8k_skew" / "fft16k_skew" (cycles, < 0 = measured default), "fir_algo" (default family of AUTO banks: 0, 1, 2),
 * "umma_slab_mib" (expanded-row workspace per slab of the tcgen05 phase-bank kernel). */
int         llz_cuda_tune(const char *key, double value);
/* page-locked host memory for the *_run_host pipelines (pageable buffers work, but slower) */
void       *llz_cuda_host_alloc(size_t bytes);
void        llz_cuda_host_free(void *p);

/* ---- selectors ------------------------------------------------------------------------- */
enum {                                      /* FIR design kind (llz_fir.c:201-269) */
    LLZ_CUDA_LPF = 0, LLZ_CUDA_HPF = 1, LLZ_CUDA_BPF = 2, LLZ_CUDA_BSF = 3,
};
enum {                                      /* FIR sample type + arithmetic */
    LLZ_CUDA_F64        = 0,  /* double I/O, FP64 arithmetic, any order (|err| <= 1e-12 rel. full scale);
                                 kernel family per LLZ_CUDA_FIR_ALGO_* below                        */
    LLZ_CUDA_F64_STRICT = 1,  /* double I/O, separate mul+add in llz_conv's order: bit-identical   */
    LLZ_CUDA_F32        = 2,  /* float I/O, FP32 arithmetic (>= 120 dB SNR); the direct kernel folds
                                 partial sums every 128 taps                                        */
};
enum {                                      /* FIR kernel family (tolerance-mode banks only; STRICT is always direct) */
    LLZ_CUDA_FIR_ALGO_AUTO   = 0, /* overlap-save for 48..12289 taps, direct form otherwise (default)        */
    LLZ_CUDA_FIR_ALGO_DIRECT = 1, /* register-blocked sliding MAC: 2*N flop per output                       */
    LLZ_CUDA_FIR_ALGO_FFT    = 2, /* overlap-save in the bank's own type: a 1024-point transform per warp up
                                     to 544 taps (~35 FMA-pipe instructions per output at 127 taps), an
                                     8192-point transform per CTA up to 3328 (f64) / 2304 (f32) taps, a
                                     16384-point transform per CTA up to 12289 taps (~57 at 4095 taps);
                                     |err| ~1e-15 (f64) / ~3e-7 (f32) of full scale against the direct sum,
                                     not bit-identical                                                      */
};
enum {                                      /* tile family of the phase-bank kernels (L >= 16); all give the same bytes */
    LLZ_CUDA_TILES_AUTO        = 0, /* exact mode: INT8 tensor cores (tcgen05 for large calls, mma.sync for
                                       frame-sized ones); fast mode: FP16 tensor cores (default)              */
    LLZ_CUDA_TILES_INT8        = 1, /* exact mode: integer evaluation on mma.sync IMMA + two-level guard      */
    LLZ_CUDA_TILES_FP64_TENSOR = 2, /* exact mode: DMMA tiles + guard                                         */
    LLZ_CUDA_TILES_CUDA_CORE   = 3, /* DFMA / FFMA register tiles                                             */
    LLZ_CUDA_TILES_INT8_TCGEN05 = 4, /* exact mode: integer evaluation on tcgen05.mma.kind::i8, accumulators in
                                        TMEM, whatever the size of the call                                   */
};
enum {                                      /* resampler accumulator */
    LLZ_CUDA_ACC_F64        = 0,  /* FP64 FMA + near-integer guard -> bit-identical int16 (default) */
    LLZ_CUDA_ACC_F64_STRICT = 1,  /* reference order, separate mul+add: bit-identical by construction */
    LLZ_CUDA_ACC_F32        = 2,  /* FP32 FMA: |diff| <= 1 LSB, exact at single-tap (knife-edge) phases */
};

/* ======================================================================================== */
/* FIR banks  (rows a12-a16 of SURVEY.md section 8a; replaces N x llz_fir_filter_*_init)      */
/* ======================================================================================== */
unsigned long llz_cuda_fir_bank_init(int kind, int flt_len, double fc1, double fc2,
                                     win_t win_type, int n_channels, int dtype);
/* explicit taps (h[0] multiplies the newest sample, as in llz_conv) */
unsigned long llz_cuda_fir_bank_init_taps(const double *h, int flt_len, int n_channels, int dtype);
void          llz_cuda_fir_bank_uninit(unsigned long handle);

int llz_cuda_fir_bank_flt_len(unsigned long handle);
/* choose the kernel family; fails (-1) for FFT on a STRICT bank or beyond 12289 taps */
int llz_cuda_fir_bank_set_algo(unsigned long handle, int algo);
/* the family the next _run will use: LLZ_CUDA_FIR_ALGO_DIRECT or LLZ_CUDA_FIR_ALGO_FFT */
int llz_cuda_fir_bank_get_algo(unsigned long handle);
/* transform length of the overlap-save family: 0 = by tap count (default), 1024, 8192 or 16384; fails when the
 * filter does not fit.  Crossover measurements and tests use it; results agree to rounding between lengths.
 * A bank on the 16384-point kernel owns a scratch in device memory (256 KB per resident CTA, twice: 76 MB on a
 * B200), allocated by its first _run and freed with the handle: the half of every work item that does not fit
 * shared memory waits there, in L2.                                                                             */
int llz_cuda_fir_bank_set_fft_size(unsigned long handle, int fft_size);
/* samples per work item of the kernel the next _run will use: 2*(1024 - halo), 2*(8192 - halo) or 2*(16384 - halo) for the
 * overlap-save kernels, 1 for the direct form.  A stream cut at multiples of this length (time segments with
 * their flt_len-1 halo, pipeline chunks) reproduces the one-shot result bit for bit.                        */
long long llz_cuda_fir_bank_block_len(unsigned long handle);
int llz_cuda_fir_bank_copy_taps(unsigned long handle, double *h_out);      /* host copy, flt_len doubles */

/* stream state = the last flt_len-1 samples of every channel (llz_fir.c:562-564) */
int llz_cuda_fir_bank_reset(unsigned long handle, llz_cuda_stream_t stream);
/* d_hist: device, planar, flt_len-1 samples per channel (the halo of a time segment) */
int llz_cuda_fir_bank_set_history(unsigned long handle, const void *d_hist, long long stride,
                                  llz_cuda_stream_t stream);

/* y[c][t] = sum_i h[i]*x[c][t-i], t in [0,n), continuing from the stored history; then the
 * history becomes the last flt_len-1 inputs.  d_in/d_out: device, dtype of the bank, strides in
 * elements.  Asynchronous on `stream`.  d_in and d_out must not overlap.                     */
int llz_cuda_fir_bank_run(unsigned long handle, const void *d_in, long long in_stride,
                          void *d_out, long long out_stride, long long n,
                          llz_cuda_stream_t stream);
/* flt_len-1 tail samples per channel (llz_fir.c:590-625); history becomes zeros */
int llz_cuda_fir_bank_flush(unsigned long handle, void *d_out, long long out_stride,
                            llz_cuda_stream_t stream);
/* same as _run but h_in/h_out are HOST buffers (pinned or pageable); chunked H2D / kernel /
 * D2H pipeline on three internal streams; synchronous.                                       */
int llz_cuda_fir_bank_run_host(unsigned long handle, const void *h_in, long long in_stride,
                               void *h_out, long long out_stride, long long n);

/* ======================================================================================== */
/* Resampler banks  (rows a9-a11, a17-a20)                                                   */
/* ======================================================================================== */
typedef struct {
    int kind;            /* 0 decimate, 1 interp, 2 resample (the CLI's -t values)            */
    int L, M;
    int n;               /* prototype length                                                  */
    int taps_per_phase;  /* Q (resample) or K (decimate/interp)                               */
    int num_in, num_out; /* reference frame size in samples (llz_resample.c:291-295,338-341,394-400) */
    int n_channels;
    int acc;
} llz_cuda_resample_info_t;

/* k_override > 0 forces the prototype to n = 2*k_override*L + 1 taps (Q = 2*k_override+1 per
 * phase for L>1); 0 = the reference's rule n0/(2L) (llz_resample.c:204-222).                 */
unsigned long llz_cuda_resample_bank_init(int L, int M, double gain, win_t win_type,
                                          int k_override, int n_channels, int acc);
unsigned long llz_cuda_decimate_bank_init(int M, double gain, win_t win_type,
                                          int n_channels, int acc);
unsigned long llz_cuda_interp_bank_init(int L, double gain, win_t win_type,
                                        int n_channels, int acc);
void          llz_cuda_resample_bank_uninit(unsigned long handle);

int llz_cuda_resample_bank_info(unsigned long handle, llz_cuda_resample_info_t *info);
/* host copies for parity checks: prototype (n doubles) and bank (rows*taps_per_phase doubles) */
int llz_cuda_resample_bank_copy_proto(unsigned long handle, double *h_out);
int llz_cuda_resample_bank_copy_bank(unsigned long handle, double *bank_out);

/* number of outputs the next _run(n_in) will produce (resample: ceil((consumed+n_in)*L/M) -
 * produced; decimate: floor; interp: n_in*L) */
long long llz_cuda_resample_bank_out_len(unsigned long handle, long long n_in);

int llz_cuda_resample_bank_reset(unsigned long handle, llz_cuda_stream_t stream);
/* halo of a time segment: taps_per_phase-1 (resample) / n (decimate) int16 samples per channel
 * that precede the segment; also rewinds the phase to output index 0.                        */
int llz_cuda_resample_bank_set_history(unsigned long handle, const short *d_hist,
                                       long long stride, llz_cuda_stream_t stream);

/* int16 planar in, int16 planar out, device pointers, asynchronous.  *n_out (host) receives the
 * outputs per channel.  Equivalent to feeding the reference frame by frame when n_in is a
 * multiple of num_in; any n_in is accepted for resample/decimate (interp: multiple of num_in,
 * because the reference restarts its window at every frame, llz_resample.c:515-523).         */
int llz_cuda_resample_bank_run(unsigned long handle, const short *d_in, long long in_stride,
                               long long n_in, short *d_out, long long out_stride,
                               long long *n_out, llz_cuda_stream_t stream);
int llz_cuda_resample_bank_run_host(unsigned long handle, const short *h_in, long long in_stride,
                                    long long n_in, short *h_out, long long out_stride,
                                    long long *n_out);
/* Interleaved PCM frames in (formats: LLZ_CUDA_PCM_* below, channels = the bank's n_channels):
 * d_frames holds n_frames frames; channel c of the bank filters sample c of every frame, converted to the int16 the
 * resampler works on exactly as llz_cuda_pcm_deinterleave does (s16 as is, s24 >> 8, f32 trunc(clamp(x * 2^15))).
 * The output is planar int16 like llz_cuda_resample_bank_run's.  For calls large enough for the tcgen05 kernel the
 * de-interleave is FUSED into the kernel's load stage (its pre-pass gathers the channels straight out of the frames:
 * no planar copy of the input is written); frame-sized calls de-interleave into a scratch buffer first.
 * _host: file-sized jobs from and to host memory, interleaved frames out (any LLZ_CUDA_PCM_* format).               */
int llz_cuda_resample_bank_run_pcm(unsigned long handle, const void *d_frames, int pcm_format, long long n_frames,
                                   short *d_out, long long out_stride, long long *n_out, llz_cuda_stream_t stream);
int llz_cuda_resample_bank_run_pcm_host(unsigned long handle, const void *h_frames, int in_format, long long n_frames,
                                        void *h_out_frames, int out_format, long long out_cap_frames,
                                        long long *n_out_frames);
/* kernel launches of the handle's last run call and the name of the kernel that did the filtering (reporting only) */
int llz_cuda_resample_bank_last_run(unsigned long handle, int *launches, char *kernel, int kernel_cap);
/* outputs that took the reference-order recompute (near-integer guard) since init */
long long llz_cuda_resample_bank_guard_count(unsigned long handle);
/* Verification knobs.  set_tiles picks the tile family of the phase-bank kernels (LLZ_CUDA_TILES_*).
 * set_guard_scale multiplies the guard bands of the exact mode (1 <= scale <= 1e12): the outputs stay bit-identical
 * -- the recompute IS the reference's own sum -- but a measurable share of them takes the recompute path, so tests
 * can exercise the branch that the production band (about 1e-8 of the outputs) almost never enters.             */
int llz_cuda_resample_bank_set_tiles(unsigned long handle, int tiles);
int llz_cuda_resample_bank_set_guard_scale(unsigned long handle, double scale);

/* ======================================================================================== */
/* Shard planners (host integer arithmetic; no CUDA needed)                                  */
/* ======================================================================================== */
/* contiguous channel shards, remainder spread over the first ranks */
int llz_cuda_shard_channels(int n_channels, int world, int rank, int *first, int *count);

typedef struct {
    long long in_start;    /* first input sample owned by this rank                           */
    long long in_count;    /* owned input samples                                             */
    long long halo;        /* samples before in_start this rank must also read (<= in_start)  */
    long long out_start;   /* first output sample produced by this rank                       */
    long long out_count;
} llz_cuda_segment_t;

/* FIR: split n samples into `world` contiguous segments; halo = flt_len-1 */
int llz_cuda_shard_fir_segments(long long n, int flt_len, int world, int rank,
                                llz_cuda_segment_t *seg);
/* same, with every internal boundary at a multiple of `granule` samples (llz_cuda_fir_bank_block_len): the
 * concatenated segment outputs are then byte-identical to the one-shot run for the overlap-save kernels too */
int llz_cuda_shard_fir_segments_aligned(long long n, int flt_len, long long granule, int world, int rank,
                                        llz_cuda_segment_t *seg);
/* resample: split n_in (a multiple of `frame_in`, the handle's num_in) into whole-frame runs so
 * that every segment starts at output index == 0 mod L (phase 0) and input index == 0 mod M/gcd;
 * halo = taps_per_phase-1 */
int llz_cuda_shard_resample_segments(long long n_in, int L, int M, int taps_per_phase,
                                     int frame_in, int world, int rank, llz_cuda_segment_t *seg);

/* ======================================================================================== */
/* Multi-GPU contexts and sharded jobs (SURVEY.md 8b extension 4, 8e)                        */
/* ======================================================================================== */
/* The reference is one mono stream per handle; what it offers for partitioning is its frame streaming with a history
 * prefix (llz_fir.c:561-566, llz_resample.c:570-576).  A job here is ONE whole-signal call over n_channels planar
 * channels, cut either by channel (independent handles) or into time segments that carry their flt_len-1 / Q-1 halo;
 * the concatenated result is byte-identical to the single-GPU call.  No exchange happens during the computation; the
 * optional gather puts the whole planar result into one rank's device memory.
 *
 * Two process models share the API: ONE process driving n GPUs (init_all) or one process PER GPU (init_rank with a
 * shared 128-byte NCCL id: torchrun / MPI).  A context has `local_count` local slots (n or 1); per-slot arguments are
 * arrays of local_count entries.  NCCL is loaded at run time (libnccl.so.2) when the first context is created.        */
enum { LLZ_CUDA_SHARD_CHANNEL = 0, LLZ_CUDA_SHARD_TIME = 1 };
enum {
    LLZ_CUDA_GATHER_NONE = 0,  /* every rank keeps its shard in d_out                                               */
    LLZ_CUDA_GATHER_NCCL = 1,  /* shards are computed in `chunks` pieces; piece c travels to the root by grouped
                                  ncclSend/ncclRecv on a second stream while piece c+1 is computed                   */
    LLZ_CUDA_GATHER_PEER = 2,  /* the kernels store straight into the root's buffer over NVLink (peer mapping / CUDA
                                  IPC): compute and gather are one kernel, nothing is staged                         */
    LLZ_CUDA_GATHER_COPY = 3,  /* pieces like GATHER_NCCL, pushed into the root's buffer through the peer mapping by the
                                  rank's copy engine on a second stream: needs no SM, so it overlaps kernels that
                                  occupy every SM (the persistent tcgen05 and overlap-save kernels)                  */
};
typedef struct {
    int first_channel, n_channels;   /* channels owned by the rank                                                  */
    llz_cuda_segment_t seg;          /* samples owned by the rank (the whole signal for channel shards)             */
} llz_cuda_shard_t;

int           llz_cuda_mgpu_unique_id(unsigned char id[128]);               /* rank 0 makes it, everybody gets a copy */
unsigned long llz_cuda_mgpu_init_rank(const unsigned char id[128], int world, int rank);   /* uses the current device */
unsigned long llz_cuda_mgpu_init_all(int n_gpus, const int *devices /* NULL = 0..n_gpus-1 */);
void          llz_cuda_mgpu_uninit(unsigned long ctx);
int           llz_cuda_mgpu_world(unsigned long ctx);
int           llz_cuda_mgpu_local_count(unsigned long ctx);
int           llz_cuda_mgpu_local_rank(unsigned long ctx, int local_idx);
int           llz_cuda_mgpu_local_device(unsigned long ctx, int local_idx);
/* The gathered result lives in a buffer of `bytes` on rank `root`'s device, owned by the context (collective call:
 * every process, same arguments).  result_ptr is the address under which local slot `local_idx` reaches it: the
 * allocation itself on the root, a peer / IPC mapping elsewhere (NULL when the device has no P2P path to the root).  */
int           llz_cuda_mgpu_result_alloc(unsigned long ctx, int root, size_t bytes);
void         *llz_cuda_mgpu_result_ptr(unsigned long ctx, int local_idx);
int           llz_cuda_mgpu_result_free(unsigned long ctx);

/* one bank per local slot, sized for the slot's shard (channel mode: its share of n_channels; time mode: all of them) */
unsigned long llz_cuda_mgpu_fir_init(unsigned long ctx, int kind, int flt_len, double fc1, double fc2,
                                     win_t win_type, int n_channels, int dtype, int shard_mode);
unsigned long llz_cuda_mgpu_resample_init(unsigned long ctx, int L, int M, double gain, win_t win_type,
                                          int k_override, int n_channels, int acc, int shard_mode);
void          llz_cuda_mgpu_job_uninit(unsigned long job);
unsigned long llz_cuda_mgpu_job_bank(unsigned long job, int local_idx);      /* the slot's bank (set_algo, info ...)  */
/* what `rank` owns of a job over n_total input samples per channel, and the outputs per channel of the whole job;
 * time mode: FIR boundaries fall on multiples of the bank's work-item length, resampler boundaries on whole frames
 * (n_total must be a multiple of num_in)                                                                             */
int           llz_cuda_mgpu_job_plan(unsigned long job, long long n_total, int rank, llz_cuda_shard_t *shard);
long long     llz_cuda_mgpu_job_out_len(unsigned long job, long long n_total);
/* One whole-signal call, sharded.  Per local slot i (rank r, shard s = job_plan(r)):
 *   d_in[i]   the rank's input: channel s.first_channel, sample s.seg.in_start - s.seg.halo (the halo comes with the
 *             rank's own slice: no exchange), in_stride[i] elements between channels;
 *   d_out[i]  the rank's shard [s.n_channels][s.seg.out_count] (GATHER_NONE; GATHER_NCCL: dense, out_stride ==
 *             out_count; ignored by the root and by GATHER_PEER);
 *   streams[i] the caller's stream on the slot's device; everything the call enqueues -- on the root including the
 *             arrival of every other rank's outputs -- is ordered before later work on that stream.
 * gather != NONE: the whole planar result [n_channels][result_stride] lands in the context's result buffer.
 * chunks: pieces per shard for GATHER_NCCL (1..8, 0 = 4).  Asynchronous; multi-process: collective.                   */
int           llz_cuda_mgpu_job_run(unsigned long job, long long n_total, const void *const *d_in,
                                    const long long *in_stride, void *const *d_out, const long long *out_stride,
                                    long long result_stride, int gather, int chunks,
                                    const llz_cuda_stream_t *streams);

/* ======================================================================================== */
/* IIR banks (drop-in: llz_iir.h)                                        */
/* ======================================================================================== */
/* n_channels independent direct-form IIR filters with the same coefficients (llz_iir.c:103-132), planar doubles on
 * device pointers, asynchronous on `stream`; one GPU thread per channel (the recurrence is serial in time), the
 * reference's operation order, bit-identical doubles.  The last N inputs and M outputs of every channel stay in the
 * handle, so consecutive calls continue the streams.  a[0..M] (a[0] taken as 1), b[0..N] (NULL = zeros); M, N <= 32.
 * d_x == NULL feeds zeros (the flush).  d_x and d_y may alias.                                                      */
unsigned long llz_cuda_iir_bank_init(int M, const double *a, int N, const double *b, int n_channels);
void          llz_cuda_iir_bank_uninit(unsigned long handle);
int           llz_cuda_iir_bank_reset(unsigned long handle, llz_cuda_stream_t stream);
int           llz_cuda_iir_bank_run(unsigned long handle, const double *d_x, long long x_stride, double *d_y,
                                    long long y_stride, long long n, llz_cuda_stream_t stream);

/* ======================================================================================== */
/* Interleaved PCM frames <-> planar channels                           */
/* ======================================================================================== */
/* The reference filters a multi-channel WAV as one interleaved mono stream (quirk R7,
 * example/llz_resample/main.c:60-62); real PCM frames are interleaved while the banks above are
 * planar.  These HBM-bound kernels convert between the two layouts on the device, fused with
 * the sample-format conversion.  Device pointers, asynchronous on `stream`.
 *   s16/s24 -> f32/f64: x / 2^15 (2^23);  -> s16: copy / arithmetic shift by 8
 *   f32 -> s16 and f32/f64 -> s16/s24: trunc(clamp(x * 2^15 (2^23)))  (the reference's saturate +
 *   C cast, llz_resample.c:596-601)                                                          */
enum { LLZ_CUDA_PCM_S16 = 0, LLZ_CUDA_PCM_S24 = 1 /* packed 3-byte little endian */, LLZ_CUDA_PCM_F32 = 2 };
enum { LLZ_CUDA_PLANAR_S16 = 0, LLZ_CUDA_PLANAR_F32 = 1, LLZ_CUDA_PLANAR_F64 = 2 };
int llz_cuda_pcm_deinterleave(const void *d_frames, int pcm_format, int n_channels, long long n_frames,
                              void *d_planar, int planar_type, long long planar_stride,
                              llz_cuda_stream_t stream);
int llz_cuda_pcm_interleave(const void *d_planar, int planar_type, long long planar_stride,
                            int n_channels, long long n_frames, void *d_frames, int pcm_format,
                            llz_cuda_stream_t stream);

/* ======================================================================================== */
/* Synthetic signals and machine probes (bench / tests)                                      */
/* ======================================================================================== */
/* per-channel 32-bit LCG of SURVEY.md section 8d, generated on the device with jump-ahead:
 * kind 0: double in [-1,1)   x = ((int)(s>>8) - 8388608) / 8388608.0
 * kind 1: float, same value rounded to float
 * kind 2: int16              x = (short)((int)(s>>17) - 16384)
 * channel c uses seed0 + c; element i is the (i+1)-th LCG state.                             */
int llz_cuda_synth_lcg(void *d_out, long long stride, int n_channels, long long n, int kind,
                       unsigned seed0, llz_cuda_stream_t stream);
/* the same sequence from element `first` on (a rank's time segment of a longer stream) */
int llz_cuda_synth_lcg_at(void *d_out, long long stride, int n_channels, long long first, long long n,
                          int kind, unsigned seed0, llz_cuda_stream_t stream);
/* register-resident FMA throughput of this GPU in TFLOP/s (dtype: LLZ_CUDA_F64 or LLZ_CUDA_F32) */
int llz_cuda_probe_fma(int dtype, double *tflops);

#ifdef __cplusplus
}
#endif

#endif /* _LLZ_CUDA_H */
