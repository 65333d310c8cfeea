// llz_shim.cu -- the C-ABI of libllzfilter_cuda: handles, stream state and host<->device traffic.
//
// Two layers live here:
//   * the llz_cuda_* bank entry points of include/llz_cuda.h (n_channels planar streams behind one
//     handle, device pointers, asynchronous, plus chunked host-buffer pipelines);
//   * the reference's own drop-in entry points (include/llz_fir.h, include/llz_resample.h): one
//     mono stream per handle, host buffers, synchronous.  Each is a one-channel bank plus a pinned
//     staging frame.  They replace libllzfilter/llz_fir.c:442-625 and llz_resample.c:271-617.
//
// Tap design and polyphase planning stay in host C (llz_design.c).  There is no CPU data path: every
// sample goes through the kernels of llz_cuda_fir.cu / llz_cuda_resample.cu, and without a CUDA
// device the *_init functions fail with (unsigned long)-1.
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <vector>

#include <cuda_fp16.h>

#include "llz_cuda_common.cuh"
#include "llz_fft32.cuh"
#include "llz_fir_kernels.h"
#include "llz_imma_tables.h"
#include "llz_poly_kernels.h"
#include "llz_umma_tables.h"
#include "llz_iir_kernels.h"
#include "llz_iir.h"

namespace {

using namespace llz;

constexpr unsigned long kFail = (unsigned long)-1;          // llz_resample.c:279
constexpr uint32_t kMagicFir = 0x4C5A4649u;                 // "LZFI"
constexpr uint32_t kMagicPoly = 0x4C5A5250u;                // "LZRP"

// current-device guard: a handle always runs on the device it was created on
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            switched = true;
        }
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

size_t fir_elem_size(int dtype) { return dtype == LLZ_CUDA_F32 ? sizeof(float) : sizeof(double); }

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- host-buffer pipeline ------------------------------------------------------------------------
// Three streams (H2D, compute, D2H) and kSlots staging slots on the device; chunk i uses slot
// i % kSlots.  Events order slot reuse; the compute stream sees the chunks in time order, so the
// bank's stream state (history, phase) advances exactly as in one long call.
constexpr int kSlots = 3;

struct Pipeline {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_ready[kSlots] = {}, run_done[kSlots] = {}, out_done[kSlots] = {};
    void *d_in[kSlots] = {}, *d_out[kSlots] = {};
    size_t in_bytes = 0, out_bytes = 0;      // per slot
    bool ok = false;

    int create()
    {
        LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
        LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; ++i) {
            LLZ_CUDA_TRY(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
            LLZ_CUDA_TRY(cudaEventCreateWithFlags(&run_done[i], cudaEventDisableTiming));
            LLZ_CUDA_TRY(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
        }
        ok = true;
        return 0;
    }
    int reserve(size_t need_in, size_t need_out)
    {
        if (!ok && create() != 0) return -1;
        if (need_in > in_bytes) {
            for (int i = 0; i < kSlots; ++i) {
                if (d_in[i]) cudaFree(d_in[i]);
                d_in[i] = nullptr;
                LLZ_CUDA_TRY(cudaMalloc(&d_in[i], need_in));
            }
            in_bytes = need_in;
        }
        if (need_out > out_bytes) {
            for (int i = 0; i < kSlots; ++i) {
                if (d_out[i]) cudaFree(d_out[i]);
                d_out[i] = nullptr;
                LLZ_CUDA_TRY(cudaMalloc(&d_out[i], need_out));
            }
            out_bytes = need_out;
        }
        return 0;
    }
    void destroy()
    {
        for (int i = 0; i < kSlots; ++i) {
            if (d_in[i]) cudaFree(d_in[i]);
            if (d_out[i]) cudaFree(d_out[i]);
            if (in_ready[i]) cudaEventDestroy(in_ready[i]);
            if (run_done[i]) cudaEventDestroy(run_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
        *this = Pipeline();
    }
};

// planar (strided rows) copy; cudaMemcpy2DAsync only takes pitches below 2 GiB, longer channels go row by row
int copy_planar(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                cudaMemcpyKind kind, cudaStream_t st)
{
    if (width == 0 || rows == 0) return 0;
    constexpr size_t kMaxPitch = 0x7fffffffu;
    if (rows == 1) {
        LLZ_CUDA_TRY(cudaMemcpyAsync(dst, src, width, kind, st));
    } else if (dpitch <= kMaxPitch && spitch <= kMaxPitch) {
        LLZ_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st));
    } else {
        for (size_t r = 0; r < rows; ++r)
            LLZ_CUDA_TRY(cudaMemcpyAsync(static_cast<unsigned char *>(dst) + r * dpitch,
                                         static_cast<const unsigned char *>(src) + r * spitch, width, kind, st));
    }
    return 0;
}

// Cross-stream ordering of a handle's stream state.  History, counters and frame buffers are written by whatever
// stream the previous call ran on; a call on a different stream (a caller switching streams, *_run_host on the internal
// pipeline streams after a device-resident run, a drop-in frame after a bank run) must not overtake that work.
// `pending` is cleared by the synchronous entry points once they have drained their stream.
struct StreamTrail {
    cudaStream_t last = nullptr;
    bool pending = false;
};

int trail_enter(StreamTrail &t, cudaStream_t st)
{
    if (t.pending && t.last != st) {
        cudaEvent_t ev = nullptr;
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(ev, t.last);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
        if (ev) cudaEventDestroy(ev);                          // released once the recorded work has completed
        if (e != cudaSuccess) {
            // the previous stream no longer exists (destroyed by its owner): everything on the device is older than this call
            cudaGetLastError();
            LLZ_CUDA_TRY(cudaDeviceSynchronize());
        }
    }
    t.last = st;
    t.pending = true;
    return 0;
}

// samples per channel per chunk so that one slot (in + out) stays near the slot budget (tunables().pipe_slot_mib)
long long pick_chunk(long long n, int n_channels, double bytes_per_in_sample)
{
    const double budget = tunables().pipe_slot_mib * 1024 * 1024;
    long long c = (long long)(budget / (bytes_per_in_sample * n_channels));
    c = (c / 4096) * 4096;
    if (c < 4096) c = 4096;
    if (c > n) c = n;
    return c;
}

// ---- FIR bank --------------------------------------------------------------------------------------
struct FirBank {
    uint32_t magic = kMagicFir;
    cudaStream_t s_side = nullptr;      // edge items + history update of the overlap-save path run beside the interior items
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int device = 0;
    int dtype = LLZ_CUDA_F64;
    int n_channels = 1;
    int flt_len = 0;
    int hist_len = 0;            // flt_len - 1
    int ntaps_pad = 0;
    double *h_host = nullptr;    // flt_len doubles (malloc)
    void *d_taps = nullptr;      // ntaps_pad elements of the bank's type, zero padded
    void *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    bool hist_zero = true;       // no launch needed to read zeros
    // overlap-save kernel (llz_cuda_fir_fft.cu): spectrum of the taps and inter-pass twiddles, bank's type
    int algo = LLZ_CUDA_FIR_ALGO_AUTO;
    void *d_fft_H = nullptr, *d_fft_tw = nullptr;
    void *d_fft_tw2 = nullptr, *d_fft_tw3 = nullptr;     // 8192-point kernel (llz_cuda_fir_fft8k.cu)
    void *d_fft_scratch = nullptr;                       // 16384-point kernel: the CTAs' L2-resident halves of their items
    int fft_size = 0;            // transform length the tables were built for: 1024, 8192 or 16384
    int fft_size_want = 0;       // llz_cuda_fir_bank_set_fft_size: 0 = by tap count
    StreamTrail trail;           // which stream touched the stream state last
    bool poisoned = false;       // a host pipeline failed half-way: the stream state is unreliable until a reset
    Pipeline pipe;
    // drop-in (mono, host buffers)
    int frame_len = 0;
    void *pinned = nullptr;      // frame_len elements
    void *d_frame_in = nullptr, *d_frame_in2 = nullptr, *d_frame_out = nullptr;
    int frame_flip = 0;          // which of the two frame buffers the next drop-in frame lands in
    // deferred history (drop-in frames): the last hist_len samples still sit at the end of the previous frame's device
    // buffer; the next call reads them there, so no history kernel runs between full frames
    const void *chain_src = nullptr;
    cudaStream_t s_frame = nullptr;
};

FirBank *as_fir(unsigned long handle)
{
    if (handle == 0 || handle == kFail) { llz_set_error("invalid FIR handle"); return nullptr; }
    FirBank *b = reinterpret_cast<FirBank *>(handle);
    if (b->magic != kMagicFir) { llz_set_error("handle is not a FIR handle"); return nullptr; }
    return b;
}

void fir_destroy(FirBank *b)
{
    if (!b) return;
    DeviceGuard g(b->device);
    b->pipe.destroy();
    if (b->s_side) cudaStreamDestroy(b->s_side);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->d_taps) cudaFree(b->d_taps);
    if (b->d_fft_H) cudaFree(b->d_fft_H);
    if (b->d_fft_tw) cudaFree(b->d_fft_tw);
    if (b->d_fft_tw2) cudaFree(b->d_fft_tw2);
    if (b->d_fft_tw3) cudaFree(b->d_fft_tw3);
    if (b->d_fft_scratch) cudaFree(b->d_fft_scratch);
    if (b->d_hist[0]) cudaFree(b->d_hist[0]);
    if (b->d_hist[1]) cudaFree(b->d_hist[1]);
    if (b->pinned) cudaFreeHost(b->pinned);
    if (b->d_frame_in) cudaFree(b->d_frame_in);
    if (b->d_frame_in2) cudaFree(b->d_frame_in2);
    if (b->d_frame_out) cudaFree(b->d_frame_out);
    if (b->s_frame) cudaStreamDestroy(b->s_frame);
    free(b->h_host);
    b->magic = 0;
    delete b;
}

int require_device(int *dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        llz_set_error("no usable CUDA device (%s); libllzfilter_cuda has no CPU path",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return -1;
    }
    LLZ_CUDA_TRY(cudaGetDevice(dev));
    return 0;
}

// takes ownership of h (malloc'd, flt_len doubles)
unsigned long fir_bank_create(double *h, int flt_len, int n_channels, int dtype)
{
    if (dtype != LLZ_CUDA_F64 && dtype != LLZ_CUDA_F64_STRICT && dtype != LLZ_CUDA_F32) {
        llz_set_error("unknown FIR dtype %d", dtype);
        free(h);
        return kFail;
    }
    if (n_channels < 1 || n_channels > 65535 || flt_len < 1) {
        llz_set_error("bad FIR bank shape: %d channels, %d taps", n_channels, flt_len);
        free(h);
        return kFail;
    }
    int dev = 0;
    if (require_device(&dev) != 0) { free(h); return kFail; }
    FirBank *b = new (std::nothrow) FirBank();
    if (!b) { llz_set_error("out of memory"); free(h); return kFail; }
    b->device = dev;
    b->dtype = dtype;
    b->n_channels = n_channels;
    b->flt_len = flt_len;
    b->hist_len = flt_len - 1;
    b->h_host = h;
    int variant = 0;
    const size_t es = fir_elem_size(dtype);
    b->ntaps_pad = (dtype == LLZ_CUDA_F32) ? fir_pad_taps<float>(flt_len, &variant)
                                           : fir_pad_taps<double>(flt_len, &variant);
    auto fail = [&](const char *what, cudaError_t e) {
        llz_set_error("FIR bank init: %s failed: %s", what, cudaGetErrorString(e));
        fir_destroy(b);
        return kFail;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&b->d_taps, (size_t)b->ntaps_pad * es)) != cudaSuccess) return fail("cudaMalloc(taps)", e);
    {
        std::vector<unsigned char> staged((size_t)b->ntaps_pad * es, 0);
        if (dtype == LLZ_CUDA_F32) {
            float *t = reinterpret_cast<float *>(staged.data());
            for (int i = 0; i < flt_len; ++i) t[i] = (float)h[i];
        } else {
            memcpy(staged.data(), h, sizeof(double) * (size_t)flt_len);
        }
        if ((e = cudaMemcpy(b->d_taps, staged.data(), staged.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
            return fail("cudaMemcpy(taps)", e);
    }
    if (b->hist_len > 0) {
        const size_t hb = (size_t)b->hist_len * n_channels * es;
        for (int i = 0; i < 2; ++i) {
            if ((e = cudaMalloc(&b->d_hist[i], hb)) != cudaSuccess) return fail("cudaMalloc(history)", e);
            if ((e = cudaMemset(b->d_hist[i], 0, hb)) != cudaSuccess) return fail("cudaMemset(history)", e);
        }
    }
    // The uploads above went through the legacy default stream (pageable cudaMemcpy may return before the DMA has
    // landed, cudaMemset is asynchronous), while the kernels run on non-blocking streams that do not synchronise with
    // it: drain it once here so that the first run on any stream sees complete taps and a zeroed history.
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return fail("cudaStreamSynchronize(init uploads)", e);
    return reinterpret_cast<unsigned long>(b);
}

// which kernel family the next run uses: LLZ_CUDA_FIR_ALGO_DIRECT or _FFT (-1: the forced choice is impossible).
// AUTO: overlap-save when the arithmetic is tolerance-mode and the tap count is where it wins (llz_cuda_fir_fft.cu);
// the process default tunables().fir_algo (LLZ_FIR_ALGO=direct|fft, read once) overrides AUTO (A-B measurements).
int fir_effective_algo(const FirBank *b)
{
    const bool fft_ok = b->dtype != LLZ_CUDA_F64_STRICT && b->flt_len <= kFirFft16kMaxTaps;
    int algo = b->algo;
    if (algo == LLZ_CUDA_FIR_ALGO_AUTO) {
        const int dflt = tunables().fir_algo;
        if (dflt == LLZ_CUDA_FIR_ALGO_DIRECT) algo = LLZ_CUDA_FIR_ALGO_DIRECT;
        else if (dflt == LLZ_CUDA_FIR_ALGO_FFT && fft_ok) algo = LLZ_CUDA_FIR_ALGO_FFT;
    }
    if (algo == LLZ_CUDA_FIR_ALGO_AUTO)
        algo = (fft_ok && b->flt_len >= kFirFftMinTapsAuto) ? LLZ_CUDA_FIR_ALGO_FFT : LLZ_CUDA_FIR_ALGO_DIRECT;
    if (algo == LLZ_CUDA_FIR_ALGO_FFT && !fft_ok) {
        llz_set_error("the overlap-save FIR kernels are not available for this bank (strict arithmetic or > %d taps)",
                      kFirFft16kMaxTaps);
        return -1;
    }
    return algo;
}

// transform length the overlap-save path uses for this bank: 1024 (one warp per item), 8192 (one CTA per item) or
// 16384 (one CTA per item in two rounds, half of the item in an L2-resident scratch); llz_cuda_fir_bank_set_fft_size overrides the choice where the tap count allows it
int fir_fft_size(const FirBank *b)
{
    const int want = b->fft_size_want;
    if (want == 16384) return 16384;
    if (want == 8192 && b->flt_len <= kFirFft8kMaxTaps) return 8192;
    if (want == 1024 && b->flt_len <= kFirFftMaxTaps) return 1024;
    if (b->flt_len >= (b->dtype == LLZ_CUDA_F32 ? kFirFft16kMinTapsAutoF32 : kFirFft16kMinTapsAutoF64)) return 16384;
    return b->flt_len >= kFirFft8kMinTapsAuto ? 8192 : 1024;
}

int upload_as(void **dst, const std::vector<double> &src, bool f32)
{
    void *d = nullptr;
    const size_t es = f32 ? sizeof(float) : sizeof(double);
    LLZ_CUDA_TRY(cudaMalloc(&d, src.size() * es));
    cudaError_t e;
    if (f32) {
        std::vector<float> f(src.begin(), src.end());
        e = cudaMemcpy(d, f.data(), f.size() * es, cudaMemcpyHostToDevice);
    } else {
        e = cudaMemcpy(d, src.data(), src.size() * es, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cudaFree(d);
        llz_set_error("upload of an FFT table failed: %s", cudaGetErrorString(e));
        return -1;
    }
    *dst = d;
    return 0;
}

// lazily build and upload the spectrum of the taps (1/N folded in) and the folded twiddle tables
int fir_fft_tables(FirBank *b)
{
    if (b->d_fft_H) return 0;
    const bool f32 = b->dtype == LLZ_CUDA_F32;
    b->fft_size = fir_fft_size(b);
    std::vector<double> tw(2 * kTwistEntries * kFftR);
    fft1024_make_twist_table(tw.data());
    if (upload_as(&b->d_fft_tw, tw, f32) != 0) return -1;
    std::vector<double> H(2 * (size_t)b->fft_size);
    if (b->fft_size == 16384) {
        std::vector<double> t2(2 * 16 * kTwistEntries * kFftR), t3(2 * 16 * 512);
        fft16k_make_twist2(t2.data());
        fft16k_make_twist3(t3.data());
        if (upload_as(&b->d_fft_tw2, t2, f32) != 0 || upload_as(&b->d_fft_tw3, t3, f32) != 0) return -1;
        fft16k_make_spectrum(b->h_host, b->flt_len, H.data());
        const int sms = device_sm_count();
        if (sms <= 0) return -1;
        LLZ_CUDA_TRY(cudaMalloc(&b->d_fft_scratch, f32 ? fir_fft16k_scratch_bytes<float>(sms) : fir_fft16k_scratch_bytes<double>(sms)));
    } else if (b->fft_size == 8192) {
        std::vector<double> t2(2 * 8 * kTwistEntries * kFftR), t3(2 * 16 * 256);
        fft8k_make_twist2(t2.data());
        fft8k_make_twist3(t3.data());
        if (upload_as(&b->d_fft_tw2, t2, f32) != 0 || upload_as(&b->d_fft_tw3, t3, f32) != 0) return -1;
        fft8k_make_spectrum(b->h_host, b->flt_len, H.data());
    } else {
        fft1024_make_spectrum(b->h_host, b->flt_len, H.data());
    }
    if (upload_as(&b->d_fft_H, H, f32) != 0) return -1;
    // the uploads above are pageable copies on the legacy stream; the kernels run on non-blocking streams
    LLZ_CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

// channels [c0, c0 + cc) of the bank; d_in / d_out point at channel c0.  `last` = the call completes the bank's
// step: the history ping-pong flips once every channel has been run (the host pipeline runs channel groups).
template <typename T>
int fir_run_typed(FirBank *b, const void *d_in, long long in_stride, void *d_out, long long out_stride,
                  long long n, cudaStream_t st, int c0, int cc, bool last, bool defer_history)
{
    FirLaunch<T> a{};
    a.x = static_cast<const T *>(d_in);
    a.x_stride = in_stride;
    a.y = static_cast<T *>(d_out);
    a.y_stride = out_stride;
    a.n = n;
    a.hist = b->chain_src ? static_cast<const T *>(b->chain_src)
             : (b->hist_zero || b->hist_len == 0) ? nullptr
                                                   : static_cast<const T *>(b->d_hist[b->cur]) + (size_t)c0 * b->hist_len;
    a.taps = static_cast<const T *>(b->d_taps);
    a.ntaps = b->flt_len;
    a.vec_ok = (d_in == nullptr || aligned16(d_in)) && aligned16(d_out) &&
               (cc == 1 || ((in_stride * sizeof(T)) % 16 == 0 && (out_stride * sizeof(T)) % 16 == 0));
    const int algo = fir_effective_algo(b);
    if (algo < 0) return -1;
    // Overlap-save banks: the edge items and the history update are independent of the interior items (they write other
    // outputs / the other history buffer), so they run on a side stream beside them: fork here, join at the end.  On a
    // strong-scaled shard (C2 at N = 8: 0.24 ms per step) the two extra launches in series were 10 % of the step.
    cudaStream_t side = nullptr;
    if (algo == LLZ_CUDA_FIR_ALGO_FFT && !defer_history) {
        if (!b->s_side) {
            LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&b->s_side, cudaStreamNonBlocking));
            LLZ_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
            LLZ_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
        }
        side = b->s_side;
        LLZ_CUDA_TRY(cudaEventRecord(b->ev_fork, st));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(side, b->ev_fork, 0));
    }
    if (algo == LLZ_CUDA_FIR_ALGO_FFT) {
        if (fir_fft_tables(b) != 0) return -1;
        FirFftLaunch<T> f{};
        f.side = side;
        f.x = a.x; f.x_stride = in_stride; f.y = a.y; f.y_stride = out_stride; f.n = n;
        f.hist = a.hist; f.ntaps = b->flt_len;
        f.H = static_cast<const T *>(b->d_fft_H);
        f.tw = static_cast<const T *>(b->d_fft_tw);
        f.tw2 = static_cast<const T *>(b->d_fft_tw2);
        f.tw3 = static_cast<const T *>(b->d_fft_tw3);
        f.scratch = static_cast<T *>(b->d_fft_scratch);
        const int rc = b->fft_size == 16384 ? fir_fft16k_launch<T>(f, cc, st)
                       : b->fft_size == 8192 ? fir_fft8k_launch<T>(f, cc, st) : fir_fft_launch<T>(f, cc, st);
        if (rc != 0) return -1;
    } else if (fir_launch<T>(a, cc, b->dtype == LLZ_CUDA_F64_STRICT, st) != 0) {
        return -1;
    }
    if (defer_history && b->hist_len > 0 && a.x && n >= b->hist_len && b->n_channels == 1) {
        b->chain_src = a.x + (n - b->hist_len);                // the caller keeps d_in intact until the next call
    } else if (b->hist_len > 0) {
        T *next = static_cast<T *>(b->d_hist[b->cur ^ 1]) + (size_t)c0 * b->hist_len;
        if (fir_update_history<T>(a.x, in_stride, n, a.hist, next, b->hist_len, cc, side ? side : st) != 0) return -1;
        if (last) {
            b->cur ^= 1;
            b->hist_zero = false;
            b->chain_src = nullptr;                            // a deferred history has been folded into d_hist
        }
    }
    if (side) {
        LLZ_CUDA_TRY(cudaEventRecord(b->ev_join, side));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_join, 0));
    }
    return 0;
}

int fir_run_part(FirBank *b, const void *d_in, long long in_stride, void *d_out, long long out_stride, long long n,
                 cudaStream_t st, int c0, int cc, bool last, bool defer_history = false)
{
    if (b->poisoned) { llz_set_error("FIR handle: an earlier host pipeline call failed half-way; reset the handle first"); return -1; }
    if (trail_enter(b->trail, st) != 0) return -1;
    if (b->dtype == LLZ_CUDA_F32)
        return fir_run_typed<float>(b, d_in, in_stride, d_out, out_stride, n, st, c0, cc, last, defer_history);
    return fir_run_typed<double>(b, d_in, in_stride, d_out, out_stride, n, st, c0, cc, last, defer_history);
}

int fir_run(FirBank *b, const void *d_in, long long in_stride, void *d_out, long long out_stride, long long n,
            cudaStream_t st, bool defer_history = false)
{
    if (n < 0) { llz_set_error("negative sample count"); return -1; }
    if (n == 0) return 0;
    if (!d_out) { llz_set_error("null output pointer"); return -1; }
    return fir_run_part(b, d_in, in_stride, d_out, out_stride, n, st, 0, b->n_channels, true, defer_history);
}

// ---- polyphase bank ----------------------------------------------------------------------------------
struct PolyBank {
    uint32_t magic = kMagicPoly;
    int device = 0;
    int n_channels = 1;
    int acc = LLZ_CUDA_ACC_F64;
    double gain = 1.0;
    llz_plan_t plan{};
    // device copies of the plan
    double *d_cbank = nullptr, *d_cbankT64 = nullptr, *d_cbankT64_base = nullptr, *d_slide64 = nullptr;
    float *d_cbankT32 = nullptr, *d_cbankT32_base = nullptr, *d_slide32 = nullptr;
    uint16_t *d_cbankT16h = nullptr, *d_cbankT16l = nullptr, *d_cbankT16h_base = nullptr, *d_cbankT16l_base = nullptr;
    int bank16_exp = 0;
    signed char *d_umma_tiles = nullptr;   // the same digit planes in the tcgen05 kernel's layout (llz_cuda_polybank_umma.cu)
    int umma_nchunks = 0;
    int umma_planes = 0, umma_shift = 0;   // digit planes (5 exact / 3 fast) and the scale 2^-shift of those tables
    double umma_eps = 0.0;                 // bound on |sum_k (g gain - q 2^-shift) x| for |x| <= 32768
    // interleaved PCM input of the call in progress (llz_cuda_resample_bank_run_pcm), nullptr otherwise
    const void *pcm_frames = nullptr;
    int pcm_fmt = 0;
    int16_t *d_pcm_tail = nullptr;         // [channels][hist_len]: the call's last samples de-interleaved, for the history
    int16_t *d_pcm_planar = nullptr;       // scratch: small calls are de-interleaved first (pcm_planar_cap samples per channel)
    long long pcm_planar_cap = 0;
    void *d_pcm_io[2] = {nullptr, nullptr};   // host-frames entry point: device staging of a chunk's input / output frames
    size_t pcm_io_cap[2] = {0, 0};
    int16_t *d_pcm_out = nullptr;          // ... and its planar output
    size_t pcm_out_cap = 0;
    int last_launches = 0;                 // kernel launches and dominant kernel of the last run call
    const char *last_kernel = "";
    bool umma_tried = false;               // poly_umma_prepare has run (tables exist iff umma_nchunks > 0)
    int urep = 1;                          // the tcgen05 kernel sees the bank replicated urep times (llz_umma_tables.h)
    double *d_cbank_u = nullptr;           // [L urep][Q]: the replicated bank (guard recompute, knife-edge taps)
    int *d_single_u = nullptr;             // [L urep]
    int *d_umma_weight = nullptr;          // [phase tiles]: relative tile cost for the kernel's walk
    long long umma_weight_sum = 0;
    unsigned char *d_umma_rows = nullptr;  // workspace: expanded input rows of one slab
    size_t umma_rows_cap = 0;
    signed char *d_imma_tiles = nullptr;   // int8 digit planes of the bank (llz_cuda_polybank_imma.cu)
    int imma_nchunks = 0, imma_shift = 0, imma_planes = 0;
    double imma_eps = 0.0;                 // bound on the tap-rounding error of a full-scale dot product
    int bank_pad = 0;
    int rep = 1;             // device tables hold the bank's rows `rep` times: kernels see L*rep phases, M*rep step
    int *d_order = nullptr, *d_single = nullptr;
    int slide_ntp64 = 0, slide_ntp32 = 0;
    unsigned long long *d_guard = nullptr;
    double guard_thr = 0.0;
    double guard_scale = 1.0;    // llz_cuda_resample_bank_set_guard_scale: widens the guard bands (verification)
    int tiles = LLZ_CUDA_TILES_AUTO;
    StreamTrail trail;           // which stream touched the stream state last
    bool poisoned = false;       // a host pipeline failed half-way: the stream state is unreliable until a reset
    // stream state
    long long consumed = 0, produced = 0;
    int16_t *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    bool hist_zero = true;
    Pipeline pipe;
    // drop-in
    int16_t *pinned_in = nullptr, *pinned_out = nullptr;
    int16_t *d_frame_in = nullptr, *d_frame_in2 = nullptr, *d_frame_out = nullptr;
    int frame_flip = 0;                    // which of the two frame buffers the next drop-in frame lands in
    // deferred history (drop-in frames): the stream's last hist_len samples still sit at the end of the previous
    // frame's device buffer, so the next frame reads them there and no history kernel runs between frames
    const int16_t *chain_src = nullptr;
    cudaStream_t s_frame = nullptr;
};

PolyBank *as_poly(unsigned long handle)
{
    if (handle == 0 || handle == kFail) { llz_set_error("invalid resampler handle"); return nullptr; }
    PolyBank *b = reinterpret_cast<PolyBank *>(handle);
    if (b->magic != kMagicPoly) { llz_set_error("handle is not a resampler handle"); return nullptr; }
    return b;
}

void poly_destroy(PolyBank *b)
{
    if (!b) return;
    DeviceGuard g(b->device);
    b->pipe.destroy();
    cudaFree(b->d_cbank); cudaFree(b->d_cbankT64_base); cudaFree(b->d_slide64);
    cudaFree(b->d_cbankT32_base); cudaFree(b->d_slide32);
    cudaFree(b->d_cbankT16h_base); cudaFree(b->d_cbankT16l_base);
    cudaFree(b->d_imma_tiles);
    cudaFree(b->d_umma_tiles);
    cudaFree(b->d_umma_rows);
    cudaFree(b->d_cbank_u);
    cudaFree(b->d_pcm_tail);
    cudaFree(b->d_pcm_planar);
    cudaFree(b->d_pcm_io[0]);
    cudaFree(b->d_pcm_io[1]);
    cudaFree(b->d_pcm_out);
    cudaFree(b->d_single_u);
    cudaFree(b->d_umma_weight);
    cudaFree(b->d_order); cudaFree(b->d_single); cudaFree(b->d_guard);
    cudaFree(b->d_hist[0]); cudaFree(b->d_hist[1]);
    if (b->pinned_in) cudaFreeHost(b->pinned_in);
    if (b->pinned_out) cudaFreeHost(b->pinned_out);
    cudaFree(b->d_frame_in); cudaFree(b->d_frame_in2); cudaFree(b->d_frame_out);
    if (b->s_frame) cudaStreamDestroy(b->s_frame);
    llz_plan_free(&b->plan);
    b->magic = 0;
    delete b;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    LLZ_CUDA_TRY(cudaMalloc(dst, src.size() * sizeof(T)));
    LLZ_CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Tables of the tcgen05 kernel, built by the first call that can use it (poly_umma_prepare): a drop-in handle that only
// ever sees frame-sized calls never pays for them (C4: 10 MB of digit planes + 10 MB of replicated bank).
// tcgen05.mma.kind::i8 (llz_cuda_polybank_umma.cu): digit planes of g * gain -- five for the exact mode (38-bit taps,
// two-level guard), three for the fast mode (22-bit taps, no guard).  With the gain inside the taps the output value is
// (integer sum) * 2^-s and the kernel finishes it with integer instructions.  The kernel sees the bank replicated urep
// times (rows of the sample operand then start on 16-byte boundaries, llz_umma_tables.h) -- also the decimating banks
// (L = 1), which become 64 phases of a cycle of 64 M samples, and llz_interp (M = 1), whose frames lie one after the
// other in the sample planes, each followed by zeros.
int poly_umma_prepare(PolyBank *b)
{
    if (b->umma_tried) return 0;
    b->umma_tried = true;
    const llz_plan_t &p = b->plan;
    const size_t L0 = (size_t)p.crows, Q = (size_t)p.ctaps;
    if (!((b->acc == LLZ_CUDA_ACC_F64 || b->acc == LLZ_CUDA_ACC_F32) && b->gain != 0.0 && isfinite(b->gain) &&
          (L0 > 1 || p.single_tap[0] < 0)))
        return 0;
    auto upload = [](auto **dst, const auto &v) -> int {
        LLZ_CUDA_TRY(cudaMalloc(dst, v.size() * sizeof(v[0])));
        LLZ_CUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
        return 0;
    };
    b->urep = llz::umma_replication((int)L0, p.M);
    const size_t UL = L0 * (size_t)b->urep;
    const long long UM = (long long)p.M * b->urep;
    const double cspan = 64.0 * p.M / (double)L0;          // sample span of a 64-phase tile: padded work (Q + cspan) / Q
    if (!(UL * Q * sizeof(double) <= (64u << 20) && UM < (1LL << 24) && (Q + cspan) / Q <= 4.0)) return 0;
    std::vector<double> cbu(UL * Q), cbg(UL * Q);
    std::vector<int> single_u(UL);
    for (size_t r = 0; r < UL; ++r) {
        single_u[r] = p.single_tap[r % L0];
        for (size_t k = 0; k < Q; ++k) {
            cbu[r * Q + k] = p.cbank[(r % L0) * Q + k];
            cbg[r * Q + k] = cbu[r * Q + k] * b->gain;
        }
    }
    std::vector<signed char> utiles;
    b->umma_planes = b->acc == LLZ_CUDA_ACC_F64 ? llz::kUPlanesExact : llz::kUPlanesFast;
    double qsum = 0.0;
    int nchunks = llz::poly_umma_build_tables(cbg.data(), (int)UL, (int)UM, (int)Q, b->umma_planes, &utiles, &b->umma_shift, &b->umma_eps, &qsum);
    // the 32.32 fixed-point form of the largest possible sum must fit 64 bits
    if (nchunks > 0 && ldexp(qsum * 32768.0, 32 - b->umma_shift) >= ldexp(1.0, 62)) nchunks = 0;
    if (nchunks > 0 && utiles.size() > (256u << 20)) nchunks = 0;
    if (nchunks <= 0) return 0;
    // relative cost of a phase tile for the kernel's walk
    const int n_ptiles = (int)((UL + llz::kUPB - 1) / llz::kUPB);
    std::vector<int> weight(n_ptiles);
    b->umma_weight_sum = 0;
    for (int pt = 0; pt < n_ptiles; ++pt) {
        const llz::UmmaPhaseTile t = llz::umma_phase_tile((int)UL, (int)UM, (int)Q, pt);
        bool knife = false;
        for (int l = 0; l < t.pbv; ++l) knife = knife || single_u[t.l0 + l] >= 0;
        // In cycles.  Measured (tools/umma_knife.py, per-CTA clocks of the profiling build): a tile's time does not follow its
        // K steps (9 or 10: the three sample chunks are loaded either way), and the knife-edge tiles' slower epilogue shows
        // in the fast mode only (+1000 cycles on 4500: 339 -> 392 Gsamples/s on C4); the exact mode's tiles are
        // MMA-bound and weighting them costs 3 %.
        weight[pt] = 4500 + ((knife && b->umma_planes == llz::kUPlanesFast) ? tunables().umma_knife_cycles : 0);
        b->umma_weight_sum += weight[pt];
    }
    if (upload(&b->d_umma_tiles, utiles) || upload(&b->d_cbank_u, cbu) || upload(&b->d_single_u, single_u) || upload(&b->d_umma_weight, weight))
        return -1;
    LLZ_CUDA_TRY(cudaDeviceSynchronize());                     // pageable uploads: done before any stream reads them
    b->umma_nchunks = nchunks;
    return 0;
}

int poly_upload_plan(PolyBank *b)
{
    const llz_plan_t &p = b->plan;
    // Few phases (2 <= L < 16: 2/1, 3/2, 3/1 ...) would leave the 64-phase tiles of the phase-bank kernels mostly
    // padding.  y[o] = sum_k g[o % L][k] x[floor(o*M/L) - k] is unchanged when L and M are both multiplied by r and
    // the bank's rows are repeated r times (o % (rL) selects the same row, floor(o*rM/(rL)) the same sample), so the
    // device tables carry r = floor(64/L) copies and every kernel sees one nearly full 64-phase tile.
    b->rep = (p.crows >= 2 && p.crows < 16 && p.shift == 0 && p.frame_len == 0) ? 64 / p.crows : 1;
    const size_t L0 = (size_t)p.crows, L = L0 * (size_t)b->rep, Q = (size_t)p.ctaps;
    // transposed bank [Q][L] with `pad` zero rows on both sides: the phase-bank kernel shifts rows per phase by up to
    // 64*M/L + 1 and over-runs the last chunk, and must not need bounds checks (llz_cuda_polybank.cu)
    const size_t pad = (size_t)(64.0 * p.M / p.L) + 2 + 64;
    b->bank_pad = (int)pad;
    std::vector<double> cb(L * Q), t64(L * (Q + 2 * pad), 0.0);
    std::vector<float> t32(L * (Q + 2 * pad), 0.f);
    for (size_t r = 0; r < L; ++r)
        for (size_t k = 0; k < Q; ++k) {
            const double g = p.cbank[(r % L0) * Q + k];
            cb[r * Q + k] = g;
            t64[(k + pad) * L + r] = g;
            t32[(k + pad) * L + r] = (float)g;
        }
    if (upload(&b->d_cbank, cb) || upload(&b->d_cbankT64_base, t64) || upload(&b->d_cbankT32_base, t32)) return -1;
    b->d_cbankT64 = b->d_cbankT64_base + pad * L;
    b->d_cbankT32 = b->d_cbankT32_base + pad * L;
    if (b->acc == LLZ_CUDA_ACC_F32 && L >= 16) {
        // fp16 hi/lo planes for the tensor-core fast mode (llz_cuda_polybank.cu): scale so the largest tap sits in
        // [2^14, 2^15), then g*2^e = hi + lo with hi = fp16(g*2^e), lo = fp16(g*2^e - hi): 22 significant bits
        double gmax = 0.0;
        for (double v : cb) gmax = fmax(gmax, fabs(v));
        int e2 = 0;
        if (gmax > 0.0) { frexp(gmax, &e2); e2 = 15 - e2; }
        b->bank16_exp = e2;
        std::vector<uint16_t> th(L * (Q + 2 * pad), 0), tl(L * (Q + 2 * pad), 0);
        for (size_t r = 0; r < L; ++r)
            for (size_t k = 0; k < Q; ++k) {
                const double g = ldexp(cb[r * Q + k], e2);
                const __half hi = __double2half(g);
                const __half lo = __double2half(g - (double)__half2float(hi));
                th[(k + pad) * L + r] = __half_as_ushort(hi);
                tl[(k + pad) * L + r] = __half_as_ushort(lo);
            }
        if (upload(&b->d_cbankT16h_base, th) || upload(&b->d_cbankT16l_base, tl)) return -1;
        b->d_cbankT16h = b->d_cbankT16h_base + pad * L;
        b->d_cbankT16l = b->d_cbankT16l_base + pad * L;
    }
    if (b->acc == LLZ_CUDA_ACC_F64 && L >= 16 && p.shift == 0 && p.frame_len == 0) {
        // integer tensor cores: int8 digit planes of the taps in the kernel's tile layout (five digits)
        std::vector<signed char> tiles;
        b->imma_planes = 5;
        b->imma_nchunks = llz::poly_imma_build_tables(cb.data(), (int)L, p.M * b->rep, (int)Q, b->imma_planes, &tiles, &b->imma_shift, &b->imma_eps);
        if (b->imma_nchunks > 0 && upload(&b->d_imma_tiles, tiles)) return -1;
    }
    std::vector<int> order(p.order, p.order + Q), single(L);
    for (size_t r = 0; r < L; ++r) single[r] = p.single_tap[r % L0];
    if (upload(&b->d_order, order) || upload(&b->d_single, single)) return -1;

    if (p.L == 1 && p.shift == 0 && p.frame_len == 0 && p.single_tap[0] < 0) {
        // sliding kernel: row rho holds the taps k = i*M + rho, i = 0, 1, ... (llz_cuda_resample.cu); the row
        // length leaves room for the coarsest tap granularity of the tile variants (8 vectors), zero filled
        const int M = p.M;
        const int per = (int)((Q + M - 1) / M);
        b->slide_ntp64 = (per / 16 + 2) * 16;
        b->slide_ntp32 = (per / 32 + 2) * 32;
        std::vector<double> s64((size_t)M * b->slide_ntp64, 0.0);
        std::vector<float> s32((size_t)M * b->slide_ntp32, 0.f);
        for (size_t k = 0; k < Q; ++k) {
            const size_t rho = k % M, i = k / M;
            s64[rho * b->slide_ntp64 + i] = cb[k];
            s32[rho * b->slide_ntp32 + i] = (float)cb[k];
        }
        if (upload(&b->d_slide64, s64) || upload(&b->d_slide32, s32)) return -1;
    }
    LLZ_CUDA_TRY(cudaMalloc(&b->d_guard, sizeof(unsigned long long)));
    LLZ_CUDA_TRY(cudaMemset(b->d_guard, 0, sizeof(unsigned long long)));
    if (p.hist_len > 0) {
        const size_t hb = (size_t)p.hist_len * b->n_channels * sizeof(int16_t);
        for (int i = 0; i < 2; ++i) {
            LLZ_CUDA_TRY(cudaMalloc(&b->d_hist[i], hb));
            LLZ_CUDA_TRY(cudaMemset(b->d_hist[i], 0, hb));
        }
    }
    // Guard threshold for full-scale input (|sample| <= 32768).  Two FP64 evaluations of the same Q-term dot
    // product (any order, fused or not) each differ from the exact value by at most
    // (Q+1)*u*sum|g*x| (u = 2^-53), the gain multiply adds |v|*u; so two evaluations of gain*sum
    // differ by < 2*(Q+2)*u*|gain|*rowsum*peak.  A factor 4 of slack costs nothing (the guard fires
    // on ~1e-9 of the outputs) and keeps the bound safe against second-order terms.
    b->guard_thr = 4.0 * 2.0 * (double)(Q + 2) * ldexp(1.0, -53) * fabs(b->gain) * p.abs_row_sum * 32768.0;
    // pageable uploads and memsets on the legacy default stream must have landed before a kernel on a non-blocking
    // stream reads them (see fir_bank_create)
    LLZ_CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

unsigned long poly_bank_create(int kind, int L, int M, double gain, win_t win, int k_override, int n_channels,
                               int acc)
{
    if (acc != LLZ_CUDA_ACC_F64 && acc != LLZ_CUDA_ACC_F64_STRICT && acc != LLZ_CUDA_ACC_F32) {
        llz_set_error("unknown accumulator mode %d", acc);
        return kFail;
    }
    if (n_channels < 1 || n_channels > 65535) {
        llz_set_error("bad channel count %d", n_channels);
        return kFail;
    }
    llz_plan_t plan;
    if (llz_plan_build(&plan, kind, L, M, win, k_override) != 0) return kFail;   // range checks first: no leak
    int dev = 0;
    if (require_device(&dev) != 0) { llz_plan_free(&plan); return kFail; }
    PolyBank *b = new (std::nothrow) PolyBank();
    if (!b) { llz_set_error("out of memory"); llz_plan_free(&plan); return kFail; }
    b->device = dev;
    b->n_channels = n_channels;
    b->acc = acc;
    b->gain = gain;
    b->plan = plan;
    if (poly_upload_plan(b) != 0) { poly_destroy(b); return kFail; }
    return reinterpret_cast<unsigned long>(b);
}

long long poly_out_len(const PolyBank *b, long long n_in)
{
    const llz_plan_t &p = b->plan;
    const long long total_in = b->consumed + n_in;
    if (p.kind == LLZ_KIND_INTERP) return n_in * p.L;
    if (p.kind == LLZ_KIND_DECIMATE) return total_in / p.M - b->produced;
    // outputs m with floor(m*M/L) <= total_in-1  <=>  m < total_in*L/M
    const long long total_out = (total_in * p.L + p.M - 1) / p.M;
    return total_out - b->produced;
}

// is this call big enough for the tcgen05 kernel (at least 4 tiles per SM), or is the kernel forced?  Needs no tables.
bool poly_umma_wanted(const PolyBank *b, long long outs, int cc)
{
    if (b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05) return true;
    const llz_plan_t &p = b->plan;
    const long long urep = llz::umma_replication(p.crows, p.M);
    const long long UL = (long long)p.L * urep;
    const long long cycles = (b->produced + outs - 1) / UL - b->produced / UL + 1;
    const long long n_tiles = (cycles + llz::kUJB - 1) / llz::kUJB * ((UL + llz::kUPB - 1) / llz::kUPB) * cc;
    const int sms = device_sm_count();
    return sms > 0 && n_tiles >= 4LL * sms;
}

// defer_history (single-channel drop-in frames with n_in >= hist_len): leave the history where it is -- the tail of
// d_in, which the caller keeps intact until the next call -- instead of copying it into d_hist
// channels [c0, c0 + cc) of the bank; d_in / d_out point at channel c0.  `last` = the call completes the bank's step:
// the counters advance and the history ping-pong flips once every channel group has been run (llz_mgpu.inl runs
// groups of whole channels so that each group's output can travel while the next one is computed).
int poly_run_part(PolyBank *b, const int16_t *d_in, long long in_stride, long long n_in, int16_t *d_out,
                  long long out_stride, long long *n_out, cudaStream_t st, int c0, int cc, bool last,
                  bool defer_history = false)
{
    const llz_plan_t &p = b->plan;
    if (n_in < 0) { llz_set_error("negative sample count"); return -1; }
    if (p.frame_len > 0 && n_in % p.frame_len != 0) {
        llz_set_error("interp input must be whole frames of %d samples (got %lld)", p.frame_len, n_in);
        return -1;
    }
    const long long outs = poly_out_len(b, n_in);
    if (n_out) *n_out = outs;
    if (n_in == 0) return 0;
    if (outs > 0 && !d_out) { llz_set_error("null output pointer"); return -1; }
    const bool pcm = b->pcm_frames != nullptr;
    if (pcm && (c0 != 0 || cc != b->n_channels || defer_history)) { llz_set_error("internal: PCM input takes the whole bank"); return -1; }
    if (b->poisoned) { llz_set_error("resampler handle: an earlier host pipeline call failed half-way; reset the handle first"); return -1; }
    if (trail_enter(b->trail, st) != 0) return -1;

    if (b->chain_src && !defer_history) {
        // a deferred history meets an ordinary run: materialise it first
        LLZ_CUDA_TRY(cudaMemcpyAsync(b->d_hist[b->cur], b->chain_src, (size_t)p.hist_len * sizeof(int16_t),
                                     cudaMemcpyDeviceToDevice, st));
        b->hist_zero = false;
        b->chain_src = nullptr;
    }
    PolyLaunch a{};
    a.x = d_in;
    a.x_stride = in_stride;
    a.n_in = n_in;
    a.hist = b->chain_src ? b->chain_src
             : (b->hist_zero || p.hist_len == 0) ? nullptr : b->d_hist[b->cur] + (size_t)c0 * p.hist_len;
    a.hist_len = p.hist_len;
    a.y = d_out;
    a.y_stride = out_stride;
    a.o0 = b->produced;
    a.n_out = outs;
    a.in0 = b->consumed;
    a.L = p.L * b->rep; a.M = p.M * b->rep;          // repeated rows: the same outputs (poly_upload_plan)
    a.ctaps = p.ctaps; a.shift = p.shift; a.frame_len = p.frame_len;
    a.acc = b->acc;
    a.tiles = b->tiles;
    a.gain = b->gain;
    a.guard_thr = b->guard_thr * b->guard_scale;
    a.cbank = b->d_cbank;
    a.cbankT64 = b->d_cbankT64;
    a.cbankT32 = b->d_cbankT32;
    a.bank_pad = b->bank_pad;
    a.cbankT16h = b->d_cbankT16h;
    a.cbankT16l = b->d_cbankT16l;
    a.bank16_exp = b->bank16_exp;
    a.order = b->d_order;
    a.order_len = p.ctaps;
    a.single_tap = b->d_single;
    a.slide64 = b->d_slide64;
    a.slide32 = b->d_slide32;
    a.slide_ntp64 = b->slide_ntp64;
    a.slide_ntp32 = b->slide_ntp32;
    a.guard_count = b->d_guard;
    a.imma_tiles = b->d_imma_tiles;
    a.imma_nchunks = b->imma_nchunks;
    a.imma_planes = b->imma_planes;
    a.imma_scale = ldexp(1.0, -b->imma_shift);
    // first-level band of the integer evaluation: the (scaled) FP64 band plus the taps' rounding bound
    a.imma_thr = b->guard_thr * b->guard_scale + 1.001 * fabs(b->gain) * b->imma_eps + ldexp(fabs(b->gain), -36);
    if (c0 == 0) note_reset();
    // tcgen05 kernel for calls that fill the machine (a frame-sized call keeps the mma.sync / sliding tiles: no pre-pass,
    // no workspace, lower latency); LLZ_CUDA_TILES_INT8_TCGEN05 forces it
    bool launched = false;
    if (outs > 0 && (b->tiles == LLZ_CUDA_TILES_AUTO || b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05) && poly_umma_wanted(b, outs, cc)) {
        if (poly_umma_prepare(b) != 0) return -1;
    }
    if (b->d_umma_tiles && b->umma_nchunks > 0 && outs > 0 && (b->tiles == LLZ_CUDA_TILES_AUTO || b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05)) {
        PolyLaunch u = a;
        if (pcm) {                                             // the pre-pass gathers the channels out of the frames
            const int bps = b->pcm_fmt == LLZ_CUDA_PCM_S16 ? 2 : b->pcm_fmt == LLZ_CUDA_PCM_S24 ? 3 : 4;
            u.x = static_cast<const int16_t *>(b->pcm_frames);
            u.pcm_sample_bytes = bps;
            u.pcm_frame_bytes = bps * b->n_channels;
            u.pcm_fmt = b->pcm_fmt;
        }
        u.L = p.L * b->urep;
        u.M = p.M * b->urep;
        u.cbank = b->d_cbank_u;
        u.single_tap = b->d_single_u;
        const long long cycles = (u.o0 + outs - 1) / u.L - u.o0 / u.L + 1;
        const long long n_tiles = (cycles + llz::kUJB - 1) / llz::kUJB * ((u.L + llz::kUPB - 1) / llz::kUPB) * cc;
        const int sms = device_sm_count();
        if (sms <= 0) return -1;
        if (b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05 || n_tiles >= 4LL * sms) {
            const size_t per_cycle = (size_t)2 * cc * (size_t)u.M;
            long long slab = (long long)(tunables().umma_slab_mib * 1024 * 1024 / (double)per_cycle) / llz::kUJB * llz::kUJB;
            if (slab < llz::kUJB) slab = llz::kUJB;
            if (slab > cycles) slab = (cycles + llz::kUJB - 1) / llz::kUJB * llz::kUJB;
            if (p.frame_len > 0 && u.M > 0 && p.frame_len % u.M == 0) {        // whole frames per slab
                const long long rpf = p.frame_len / u.M;
                if (rpf > llz::kUJB) slab = (slab + rpf - 1) / rpf * rpf;
            }
            const size_t need = poly_bank_umma_rows_bytes(u, cc, slab);
            if (need == 0) goto no_umma;
            if (need > b->umma_rows_cap) {
                LLZ_CUDA_TRY(cudaDeviceSynchronize());                 // an earlier call may still read the old workspace
                cudaFree(b->d_umma_rows);
                b->d_umma_rows = nullptr;
                b->umma_rows_cap = 0;
                LLZ_CUDA_TRY(cudaMalloc(&b->d_umma_rows, need));
                b->umma_rows_cap = need;
            }
            u.umma_tiles = b->d_umma_tiles;
            u.umma_weight = b->d_umma_weight;
            u.umma_weight_sum = b->umma_weight_sum;
            u.umma_nchunks = b->umma_nchunks;
            u.umma_planes = b->umma_planes;
            u.umma_scale = ldexp(1.0, -b->umma_shift);
            u.umma_ush = b->umma_shift - 32;
            // first-level band: the (scaled) FP64 band, the taps' rounding bound, the bits the 32.32 form drops
            const double thr = b->guard_thr * b->guard_scale + 1.001 * b->umma_eps + ldexp(1.0, -31);
            u.umma_thr32 = thr >= 0.4999 ? 0x7fffffffu : (unsigned)ceil(ldexp(thr, 32)) + 2u;
            u.umma_rows = b->d_umma_rows;
            u.umma_slab_cycles = (int)slab;
            const int rc = poly_bank_umma_launch(u, cc, st);
            if (rc < 0) return -1;
            launched = rc == 1;
        }
    }
no_umma:
    if (pcm && !launched) { llz_set_error("internal: interleaved input reached a kernel that needs planar samples"); return -1; }
    if (!launched && poly_launch(a, cc, st) != 0) return -1;
    if (pcm && p.hist_len > 0) {
        // history from interleaved input: de-interleave the call's last samples, then the usual splice
        const long long tail = n_in < p.hist_len ? n_in : p.hist_len;
        const int bps = b->pcm_fmt == LLZ_CUDA_PCM_S16 ? 2 : b->pcm_fmt == LLZ_CUDA_PCM_S24 ? 3 : 4;
        if (!b->d_pcm_tail) LLZ_CUDA_TRY(cudaMalloc(&b->d_pcm_tail, (size_t)p.hist_len * b->n_channels * sizeof(int16_t)));
        const unsigned char *src = static_cast<const unsigned char *>(b->pcm_frames) + (size_t)(n_in - tail) * bps * b->n_channels;
        if (llz_cuda_pcm_deinterleave(src, b->pcm_fmt, b->n_channels, tail, b->d_pcm_tail, LLZ_CUDA_PLANAR_S16, p.hist_len, st) != 0) return -1;
        d_in = b->d_pcm_tail;
        in_stride = p.hist_len;
        const int16_t *old = b->hist_zero ? nullptr : b->d_hist[b->cur];
        if (poly_update_history(d_in, in_stride, tail, old, b->d_hist[b->cur ^ 1], p.hist_len, cc, st) != 0) return -1;
        b->cur ^= 1;
        b->hist_zero = false;
    } else
    if (defer_history && p.hist_len > 0 && n_in >= p.hist_len && b->n_channels == 1) {
        b->chain_src = d_in + (n_in - p.hist_len);
    } else if (p.hist_len > 0) {
        const int16_t *old = b->hist_zero ? nullptr : b->d_hist[b->cur] + (size_t)c0 * p.hist_len;
        if (poly_update_history(d_in, in_stride, n_in, old, b->d_hist[b->cur ^ 1] + (size_t)c0 * p.hist_len, p.hist_len, cc,
                                st) != 0)
            return -1;
        if (last) {
            b->cur ^= 1;
            b->hist_zero = false;
        }
    }
    if (last) {
        b->consumed += n_in;
        b->produced += outs;
    }
    b->last_launches = noted_launches();
    b->last_kernel = noted_kernel();
    return 0;
}

int poly_run(PolyBank *b, const int16_t *d_in, long long in_stride, long long n_in, int16_t *d_out,
             long long out_stride, long long *n_out, cudaStream_t st, bool defer_history = false)
{
    return poly_run_part(b, d_in, in_stride, n_in, d_out, out_stride, n_out, st, 0, b->n_channels, true, defer_history);
}

int poly_reset(PolyBank *b)
{
    b->consumed = 0;
    b->produced = 0;
    b->hist_zero = true;
    b->chain_src = nullptr;
    b->poisoned = false;
    return 0;
}

}  // namespace

// ====================================================================================================
// runtime
// ====================================================================================================
extern "C" int llz_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *llz_cuda_build_info(void)
{
    return "libllzfilter_cuda 0.1 sm_100a (CUDA " LLZ_STR(CUDART_VERSION) "), built " __DATE__;
}

extern "C" void *llz_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        llz_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void llz_cuda_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ====================================================================================================
// FIR banks
// ====================================================================================================
extern "C" unsigned long llz_cuda_fir_bank_init(int kind, int flt_len, double fc1, double fc2, win_t win_type,
                                                int n_channels, int dtype)
{
    double *h = nullptr;
    const int n = llz_design_taps(&h, kind, flt_len, fc1, fc2, win_type);
    if (n < 0) return kFail;
    return fir_bank_create(h, n, n_channels, dtype);
}

extern "C" unsigned long llz_cuda_fir_bank_init_taps(const double *h, int flt_len, int n_channels, int dtype)
{
    if (!h || flt_len < 1) { llz_set_error("fir_bank_init_taps: null or empty taps"); return kFail; }
    double *copy = (double *)malloc(sizeof(double) * (size_t)flt_len);
    if (!copy) { llz_set_error("out of memory"); return kFail; }
    memcpy(copy, h, sizeof(double) * (size_t)flt_len);
    return fir_bank_create(copy, flt_len, n_channels, dtype);
}

extern "C" void llz_cuda_fir_bank_uninit(unsigned long handle)
{
    FirBank *b = as_fir(handle);
    if (b) fir_destroy(b);
}

extern "C" int llz_cuda_fir_bank_flt_len(unsigned long handle)
{
    FirBank *b = as_fir(handle);
    return b ? b->flt_len : -1;
}

extern "C" int llz_cuda_fir_bank_set_algo(unsigned long handle, int algo)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (algo != LLZ_CUDA_FIR_ALGO_AUTO && algo != LLZ_CUDA_FIR_ALGO_DIRECT && algo != LLZ_CUDA_FIR_ALGO_FFT) {
        llz_set_error("unknown FIR algorithm %d", algo);
        return -1;
    }
    const int prev = b->algo;
    b->algo = algo;
    if (fir_effective_algo(b) < 0) { b->algo = prev; return -1; }
    return 0;
}

extern "C" int llz_cuda_fir_bank_get_algo(unsigned long handle)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    return fir_effective_algo(b);
}

extern "C" int llz_cuda_fir_bank_set_fft_size(unsigned long handle, int fft_size)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (fft_size != 0 && fft_size != 1024 && fft_size != 8192 && fft_size != 16384) {
        llz_set_error("overlap-save transform length must be 0 (automatic), 1024, 8192 or 16384 (got %d)", fft_size);
        return -1;
    }
    if ((fft_size == 1024 && b->flt_len > kFirFftMaxTaps) || (fft_size == 8192 && b->flt_len > kFirFft8kMaxTaps) ||
        (fft_size == 16384 && b->flt_len > kFirFft16kMaxTaps)) {
        llz_set_error("a %d-point transform cannot hold %d taps", fft_size, b->flt_len);
        return -1;
    }
    if (b->d_fft_H && fir_fft_size(b) != b->fft_size) { llz_set_error("internal: stale FFT tables"); return -1; }
    if (b->d_fft_H && fft_size != b->fft_size_want) {
        // tables of another transform length are resident: drop them, the next run rebuilds
        DeviceGuard g(b->device);
        LLZ_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(b->d_fft_H); cudaFree(b->d_fft_tw); cudaFree(b->d_fft_tw2); cudaFree(b->d_fft_tw3); cudaFree(b->d_fft_scratch);
        b->d_fft_H = b->d_fft_tw = b->d_fft_tw2 = b->d_fft_tw3 = b->d_fft_scratch = nullptr;
        b->fft_size = 0;
    }
    b->fft_size_want = fft_size;
    return 0;
}

long long fir_block_len(const FirBank *b)
{
    if (fir_effective_algo(b) != LLZ_CUDA_FIR_ALGO_FFT) return 1;
    const int size = b->fft_size ? b->fft_size : fir_fft_size(b);
    if (size == 16384) return 2LL * (kFft16kN - (b->flt_len - 1 + 511) / 512 * 512);
    if (size == 8192) return 2LL * (kFft8kN - (b->flt_len - 1 + 255) / 256 * 256);
    return 2LL * (kFftN - (b->flt_len - 1 + 31) / 32 * 32);
}

extern "C" long long llz_cuda_fir_bank_block_len(unsigned long handle)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    return fir_block_len(b);
}

extern "C" int llz_cuda_fir_bank_copy_taps(unsigned long handle, double *h_out)
{
    FirBank *b = as_fir(handle);
    if (!b || !h_out) return -1;
    memcpy(h_out, b->h_host, sizeof(double) * (size_t)b->flt_len);
    return b->flt_len;
}

extern "C" int llz_cuda_fir_bank_reset(unsigned long handle, llz_cuda_stream_t stream)
{
    (void)stream;
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    b->hist_zero = true;
    b->chain_src = nullptr;
    b->poisoned = false;
    return 0;
}

extern "C" int llz_cuda_fir_bank_set_history(unsigned long handle, const void *d_hist, long long stride,
                                             llz_cuda_stream_t stream)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (b->hist_len == 0) return 0;
    b->chain_src = nullptr;
    if (!d_hist) { b->hist_zero = true; return 0; }
    DeviceGuard g(b->device);
    b->poisoned = false;
    if (trail_enter(b->trail, (cudaStream_t)stream) != 0) return -1;
    const size_t es = fir_elem_size(b->dtype);
    if (copy_planar(b->d_hist[b->cur], (size_t)b->hist_len * es, d_hist, (size_t)stride * es,
                                   (size_t)b->hist_len * es, (size_t)b->n_channels, cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream) != 0) return -1;
    b->hist_zero = false;
    return 0;
}

extern "C" int llz_cuda_fir_bank_run(unsigned long handle, const void *d_in, long long in_stride, void *d_out,
                                     long long out_stride, long long n, llz_cuda_stream_t stream)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (!d_in && n > 0) { llz_set_error("null input pointer"); return -1; }
    DeviceGuard g(b->device);
    return fir_run(b, d_in, in_stride, d_out, out_stride, n, (cudaStream_t)stream);
}

extern "C" int llz_cuda_fir_bank_flush(unsigned long handle, void *d_out, long long out_stride,
                                       llz_cuda_stream_t stream)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (b->hist_len == 0) return 0;
    DeviceGuard g(b->device);
    // a run over flt_len-1 zero samples: llz_fir.c:604-621
    if (fir_run(b, nullptr, 0, d_out, out_stride, b->hist_len, (cudaStream_t)stream) != 0) return -1;
    b->hist_zero = true;
    return b->hist_len;
}

static int fir_run_host_body(unsigned long handle, const void *h_in, long long in_stride,
                             void *h_out, long long out_stride, long long n)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (n < 0 || (n > 0 && (!h_in || !h_out))) { llz_set_error("run_host: bad arguments"); return -1; }
    if (n == 0) return 0;
    DeviceGuard g(b->device);
    const size_t es = fir_elem_size(b->dtype);
    const int C = b->n_channels;
    Pipeline &P = b->pipe;
    const unsigned char *src = static_cast<const unsigned char *>(h_in);
    unsigned char *dst = static_cast<unsigned char *>(h_out);
    {
        // Many short channels (C2: 1024 x 3.84 MB): pipeline over GROUPS OF WHOLE CHANNELS.  With dense rows a group
        // is one contiguous span, so every copy is 1-D: tools/pcie_probe2d.cu measures 44.2 GB/s H2D for contiguous
        // copies against 37.8 GB/s for the row-wise shape a time chunk of all channels needs (D2H 48.5 / 47.0), and
        // the kernels see whole channels (no per-chunk edge items).  Long channels fall through to time chunks.
        const double budget = tunables().pipe_slot_mib * 1024 * 1024;
        const double per_channel = 2.0 * (double)es * (double)n;
        long long group = (long long)(budget / per_channel);
        if (group >= 2 && C > 2 * group) {
            if (group > C) group = C;
            const size_t row = (size_t)n * es;
            if (P.reserve((size_t)group * row, (size_t)group * row) != 0) return -1;
            const bool dense_in = in_stride == n, dense_out = out_stride == n;
            long long idx = 0;
            for (long long c0 = 0; c0 < C; c0 += group, ++idx) {
                const long long cc = (C - c0 < group) ? C - c0 : group;
                const int s = (int)(idx % kSlots);
                if (idx >= kSlots) {
                    LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_in, P.run_done[s], 0));
                    LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.out_done[s], 0));
                }
                const unsigned char *hs = src + (size_t)c0 * in_stride * es;
                unsigned char *hd = dst + (size_t)c0 * out_stride * es;
                if (dense_in) LLZ_CUDA_TRY(cudaMemcpyAsync(P.d_in[s], hs, (size_t)cc * row, cudaMemcpyHostToDevice, P.s_in));
                else if (copy_planar(P.d_in[s], row, hs, (size_t)in_stride * es, row, (size_t)cc, cudaMemcpyHostToDevice,
                                     P.s_in) != 0) return -1;
                LLZ_CUDA_TRY(cudaEventRecord(P.in_ready[s], P.s_in));
                LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.in_ready[s], 0));
                if (fir_run_part(b, P.d_in[s], n, P.d_out[s], n, n, P.s_run, (int)c0, (int)cc, c0 + cc >= C) != 0) return -1;
                LLZ_CUDA_TRY(cudaEventRecord(P.run_done[s], P.s_run));
                LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_out, P.run_done[s], 0));
                if (dense_out) LLZ_CUDA_TRY(cudaMemcpyAsync(hd, P.d_out[s], (size_t)cc * row, cudaMemcpyDeviceToHost, P.s_out));
                else if (copy_planar(hd, (size_t)out_stride * es, P.d_out[s], row, row, (size_t)cc, cudaMemcpyDeviceToHost,
                                     P.s_out) != 0) return -1;
                LLZ_CUDA_TRY(cudaEventRecord(P.out_done[s], P.s_out));
            }
            LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_out));
            LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_run));
            return 0;
        }
    }
    long long chunk = pick_chunk(n, C, 2.0 * es);
    {
        // whole work items per chunk: the overlap-save kernels then reproduce the one-shot result bit for bit
        const long long blk = fir_block_len(b);
        if (chunk < n && blk > 1) chunk = (chunk >= blk) ? chunk / blk * blk : blk;
    }
    if (P.reserve((size_t)chunk * C * es, (size_t)chunk * C * es) != 0) return -1;
    long long idx = 0;
    for (long long t0 = 0; t0 < n; t0 += chunk, ++idx) {
        const long long len = (n - t0 < chunk) ? n - t0 : chunk;
        const int s = (int)(idx % kSlots);
        if (idx >= kSlots) {
            LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_in, P.run_done[s], 0));     // input slot consumed
            LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.out_done[s], 0));    // output slot drained
        }
        if (copy_planar(P.d_in[s], (size_t)chunk * es, src + (size_t)t0 * es, (size_t)in_stride * es,
                                       (size_t)len * es, (size_t)C, cudaMemcpyHostToDevice, P.s_in) != 0) return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.in_ready[s], P.s_in));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.in_ready[s], 0));
        if (fir_run(b, P.d_in[s], chunk, P.d_out[s], chunk, len, P.s_run) != 0) return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.run_done[s], P.s_run));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_out, P.run_done[s], 0));
        if (copy_planar(dst + (size_t)t0 * es, (size_t)out_stride * es, P.d_out[s], (size_t)chunk * es,
                                       (size_t)len * es, (size_t)C, cudaMemcpyDeviceToHost, P.s_out) != 0) return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.out_done[s], P.s_out));
    }
    LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_out));
    LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_run));
    return 0;
}

// A failure inside the pipeline must not return while copies are still in flight (the caller owns the host buffers
// again on return) and leaves the stream state half-advanced: drain the three streams and poison the handle until a
// reset / set_history.
extern "C" int llz_cuda_fir_bank_run_host(unsigned long handle, const void *h_in, long long in_stride,
                                          void *h_out, long long out_stride, long long n)
{
    const int rc = fir_run_host_body(handle, h_in, in_stride, h_out, out_stride, n);
    FirBank *b = reinterpret_cast<FirBank *>(handle);
    if (handle == 0 || handle == kFail || b->magic != kMagicFir) return rc;
    if (rc != 0 && b->pipe.ok) {
        DeviceGuard g(b->device);
        cudaStreamSynchronize(b->pipe.s_in);
        cudaStreamSynchronize(b->pipe.s_run);
        cudaStreamSynchronize(b->pipe.s_out);
        cudaGetLastError();
        b->poisoned = true;
    }
    b->trail.pending = false;                                  // the pipeline streams are drained on every path
    return rc;
}

// ====================================================================================================
// resampler banks
// ====================================================================================================
extern "C" unsigned long llz_cuda_resample_bank_init(int L, int M, double gain, win_t win_type, int k_override,
                                                     int n_channels, int acc)
{
    return poly_bank_create(LLZ_KIND_RESAMPLE, L, M, gain, win_type, k_override, n_channels, acc);
}

extern "C" unsigned long llz_cuda_decimate_bank_init(int M, double gain, win_t win_type, int n_channels, int acc)
{
    return poly_bank_create(LLZ_KIND_DECIMATE, 1, M, gain, win_type, 0, n_channels, acc);
}

extern "C" unsigned long llz_cuda_interp_bank_init(int L, double gain, win_t win_type, int n_channels, int acc)
{
    return poly_bank_create(LLZ_KIND_INTERP, L, 1, gain, win_type, 0, n_channels, acc);
}

extern "C" void llz_cuda_resample_bank_uninit(unsigned long handle)
{
    PolyBank *b = as_poly(handle);
    if (b) poly_destroy(b);
}

extern "C" int llz_cuda_resample_bank_info(unsigned long handle, llz_cuda_resample_info_t *info)
{
    PolyBank *b = as_poly(handle);
    if (!b || !info) return -1;
    const llz_plan_t &p = b->plan;
    info->kind = p.kind;
    info->L = p.L;
    info->M = p.M;
    info->n = p.n;
    info->taps_per_phase = p.cols;
    info->num_in = p.num_in;
    info->num_out = p.num_out;
    info->n_channels = b->n_channels;
    info->acc = b->acc;
    return 0;
}

extern "C" int llz_cuda_resample_bank_copy_proto(unsigned long handle, double *h_out)
{
    PolyBank *b = as_poly(handle);
    if (!b || !h_out) return -1;
    memcpy(h_out, b->plan.proto, sizeof(double) * (size_t)b->plan.n);
    return b->plan.n;
}

extern "C" int llz_cuda_resample_bank_copy_bank(unsigned long handle, double *bank_out)
{
    PolyBank *b = as_poly(handle);
    if (!b || !bank_out) return -1;
    const size_t cnt = (size_t)b->plan.rows * b->plan.cols;
    memcpy(bank_out, b->plan.bank, sizeof(double) * cnt);
    return (int)cnt;
}

extern "C" long long llz_cuda_resample_bank_out_len(unsigned long handle, long long n_in)
{
    PolyBank *b = as_poly(handle);
    if (!b || n_in < 0) return -1;
    return poly_out_len(b, n_in);
}

extern "C" int llz_cuda_resample_bank_reset(unsigned long handle, llz_cuda_stream_t stream)
{
    (void)stream;
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    return poly_reset(b);
}

extern "C" int llz_cuda_resample_bank_set_history(unsigned long handle, const short *d_hist, long long stride,
                                                  llz_cuda_stream_t stream)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    poly_reset(b);
    if (b->plan.hist_len == 0 || !d_hist) return 0;
    DeviceGuard g(b->device);
    if (trail_enter(b->trail, (cudaStream_t)stream) != 0) return -1;
    const size_t row = (size_t)b->plan.hist_len * sizeof(int16_t);
    if (copy_planar(b->d_hist[b->cur], row, d_hist, (size_t)stride * sizeof(int16_t), row,
                                   (size_t)b->n_channels, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != 0) return -1;
    b->hist_zero = false;
    return 0;
}

extern "C" int llz_cuda_resample_bank_run(unsigned long handle, const short *d_in, long long in_stride,
                                          long long n_in, short *d_out, long long out_stride, long long *n_out,
                                          llz_cuda_stream_t stream)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (!d_in && n_in > 0) { llz_set_error("null input pointer"); return -1; }
    DeviceGuard g(b->device);
    return poly_run(b, d_in, in_stride, n_in, d_out, out_stride, n_out, (cudaStream_t)stream);
}

// ---- interleaved PCM frames in (fused into the load stage) ----
namespace {

int pcm_sample_bytes(int fmt) { return fmt == LLZ_CUDA_PCM_S16 ? 2 : fmt == LLZ_CUDA_PCM_S24 ? 3 : fmt == LLZ_CUDA_PCM_F32 ? 4 : 0; }

// will poly_run_part put this call on the tcgen05 kernel (whose pre-pass reads the frames directly)?
bool poly_pcm_fused(PolyBank *b, long long n_in)
{
    const llz_plan_t &p = b->plan;
    if (!(b->tiles == LLZ_CUDA_TILES_AUTO || b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05)) return false;
    const long long outs = poly_out_len(b, n_in);
    if (outs <= 0 || !poly_umma_wanted(b, outs, b->n_channels)) return false;
    if (poly_umma_prepare(b) != 0 || b->umma_nchunks <= 0) return false;
    PolyLaunch u{};
    u.L = p.L * b->urep; u.M = p.M * b->urep; u.ctaps = p.ctaps; u.frame_len = p.frame_len;
    if (poly_bank_umma_rows_bytes(u, b->n_channels, llz::kUJB) == 0) return false;
    const long long cycles = (b->produced + outs - 1) / u.L - b->produced / u.L + 1;
    const long long n_tiles = (cycles + llz::kUJB - 1) / llz::kUJB * ((u.L + llz::kUPB - 1) / llz::kUPB) * b->n_channels;
    const int sms = device_sm_count();
    return b->tiles == LLZ_CUDA_TILES_INT8_TCGEN05 || (sms > 0 && n_tiles >= 4LL * sms);
}

}  // namespace

extern "C" int llz_cuda_resample_bank_run_pcm(unsigned long handle, const void *d_frames, int pcm_format, long long n_frames,
                                              short *d_out, long long out_stride, long long *n_out, llz_cuda_stream_t stream)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (pcm_sample_bytes(pcm_format) == 0) { llz_set_error("unknown PCM format %d", pcm_format); return -1; }
    if (!d_frames && n_frames > 0) { llz_set_error("null input pointer"); return -1; }
    if (n_frames < 0) { llz_set_error("negative frame count"); return -1; }
    DeviceGuard g(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_frames > 0 && poly_pcm_fused(b, n_frames)) {
        b->pcm_frames = d_frames;
        b->pcm_fmt = pcm_format;
        const int rc = poly_run(b, nullptr, 0, n_frames, d_out, out_stride, n_out, st);
        b->pcm_frames = nullptr;
        return rc;
    }
    // frame-sized calls: de-interleave first (llz_cuda_pcm.cu), then the planar kernels
    if (n_frames > b->pcm_planar_cap) {
        LLZ_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(b->d_pcm_planar);
        b->d_pcm_planar = nullptr;
        b->pcm_planar_cap = 0;
        LLZ_CUDA_TRY(cudaMalloc(&b->d_pcm_planar, (size_t)n_frames * b->n_channels * sizeof(int16_t)));
        b->pcm_planar_cap = n_frames;
    }
    if (n_frames > 0 && llz_cuda_pcm_deinterleave(d_frames, pcm_format, b->n_channels, n_frames, b->d_pcm_planar, LLZ_CUDA_PLANAR_S16,
                                                  n_frames, stream) != 0)
        return -1;
    return poly_run(b, b->d_pcm_planar, n_frames, n_frames, d_out, out_stride, n_out, st);
}

extern "C" int llz_cuda_resample_bank_run_pcm_host(unsigned long handle, const void *h_frames, int in_format, long long n_frames,
                                                   void *h_out_frames, int out_format, long long out_cap_frames,
                                                   long long *n_out_frames)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    const int ibs = pcm_sample_bytes(in_format), obs = pcm_sample_bytes(out_format);
    if (!ibs || !obs) { llz_set_error("unknown PCM format"); return -1; }
    if (n_frames < 0 || (n_frames > 0 && (!h_frames || !h_out_frames))) { llz_set_error("run_pcm_host: bad arguments"); return -1; }
    const llz_plan_t &p = b->plan;
    const int C = b->n_channels;
    const long long total_out = poly_out_len(b, n_frames);
    if (n_out_frames) *n_out_frames = total_out;
    if (total_out > out_cap_frames) { llz_set_error("run_pcm_host: %lld output frames, room for %lld", total_out, out_cap_frames); return -1; }
    if (n_frames == 0) return 0;
    DeviceGuard g(b->device);
    // chunks of whole reference frames (keeps interp legal); one stream: this entry point is for file-sized jobs
    long long chunk = (long long)(256.0 * 1024 * 1024 / ((double)ibs * C));
    chunk = chunk / p.num_in * p.num_in;
    if (chunk < p.num_in) chunk = p.num_in;
    if (chunk > n_frames) chunk = n_frames;
    const long long chunk_out = (long long)ceil((double)chunk * p.L / p.M) + 2;
    const size_t need_in = (size_t)chunk * C * ibs, need_out = (size_t)chunk_out * C * obs, need_pl = (size_t)chunk_out * C * sizeof(int16_t);
    if (need_in > b->pcm_io_cap[0]) { cudaFree(b->d_pcm_io[0]); b->pcm_io_cap[0] = 0; LLZ_CUDA_TRY(cudaMalloc(&b->d_pcm_io[0], need_in)); b->pcm_io_cap[0] = need_in; }
    if (need_out > b->pcm_io_cap[1]) { cudaFree(b->d_pcm_io[1]); b->pcm_io_cap[1] = 0; LLZ_CUDA_TRY(cudaMalloc(&b->d_pcm_io[1], need_out)); b->pcm_io_cap[1] = need_out; }
    if (need_pl > b->pcm_out_cap) { cudaFree(b->d_pcm_out); b->pcm_out_cap = 0; LLZ_CUDA_TRY(cudaMalloc(&b->d_pcm_out, need_pl)); b->pcm_out_cap = need_pl; }
    long long done_out = 0;
    for (long long f0 = 0; f0 < n_frames; f0 += chunk) {
        const long long nf = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
        LLZ_CUDA_TRY(cudaMemcpyAsync(b->d_pcm_io[0], static_cast<const unsigned char *>(h_frames) + (size_t)f0 * C * ibs, (size_t)nf * C * ibs,
                                     cudaMemcpyHostToDevice, nullptr));
        long long outs = 0;
        if (llz_cuda_resample_bank_run_pcm(handle, b->d_pcm_io[0], in_format, nf, b->d_pcm_out, chunk_out, &outs, nullptr) != 0) return -1;
        if (outs > chunk_out) { llz_set_error("internal: chunk produced %lld outputs", outs); return -1; }
        if (outs > 0) {
            if (llz_cuda_pcm_interleave(b->d_pcm_out, LLZ_CUDA_PLANAR_S16, chunk_out, C, outs, b->d_pcm_io[1], out_format, nullptr) != 0) return -1;
            LLZ_CUDA_TRY(cudaMemcpyAsync(static_cast<unsigned char *>(h_out_frames) + (size_t)done_out * C * obs, b->d_pcm_io[1],
                                         (size_t)outs * C * obs, cudaMemcpyDeviceToHost, nullptr));
        }
        LLZ_CUDA_TRY(cudaStreamSynchronize(nullptr));
        done_out += outs;
    }
    if (done_out != total_out) { llz_set_error("internal: %lld of %lld output frames", done_out, total_out); return -1; }
    return 0;
}

static int poly_run_host_body(unsigned long handle, const short *h_in, long long in_stride,
                              long long n_in, short *h_out, long long out_stride,
                              long long *n_out)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (n_in < 0 || (n_in > 0 && (!h_in || !h_out))) { llz_set_error("run_host: bad arguments"); return -1; }
    if (n_out) *n_out = poly_out_len(b, n_in);
    if (n_in == 0) return 0;
    DeviceGuard g(b->device);
    const llz_plan_t &p = b->plan;
    const int C = b->n_channels;
    const double ratio = (double)p.L / p.M;
    // chunk: whole reference frames (keeps interp legal and every chunk's output count exact)
    long long chunk = pick_chunk(n_in, C, 2.0 * (1.0 + ratio));
    if (chunk < n_in) {
        chunk = chunk / p.num_in * p.num_in;
        if (chunk < p.num_in) chunk = p.num_in;
    }
    const long long chunk_out = (long long)ceil((double)chunk * ratio) + 2;
    Pipeline &P = b->pipe;
    if (P.reserve((size_t)chunk * C * 2, (size_t)chunk_out * C * 2) != 0) return -1;
    long long idx = 0, out_pos = 0;
    for (long long t0 = 0; t0 < n_in; t0 += chunk, ++idx) {
        const long long len = (n_in - t0 < chunk) ? n_in - t0 : chunk;
        const int s = (int)(idx % kSlots);
        if (idx >= kSlots) {
            LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_in, P.run_done[s], 0));
            LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.out_done[s], 0));
        }
        if (copy_planar(P.d_in[s], (size_t)chunk * 2, h_in + t0, (size_t)in_stride * 2, (size_t)len * 2,
                                       (size_t)C, cudaMemcpyHostToDevice, P.s_in) != 0) return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.in_ready[s], P.s_in));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_run, P.in_ready[s], 0));
        long long outs = 0;
        if (poly_run(b, (const int16_t *)P.d_in[s], chunk, len, (int16_t *)P.d_out[s], chunk_out, &outs, P.s_run) != 0)
            return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.run_done[s], P.s_run));
        LLZ_CUDA_TRY(cudaStreamWaitEvent(P.s_out, P.run_done[s], 0));
        if (outs > 0)
            if (copy_planar(h_out + out_pos, (size_t)out_stride * 2, P.d_out[s], (size_t)chunk_out * 2,
                                           (size_t)outs * 2, (size_t)C, cudaMemcpyDeviceToHost, P.s_out) != 0) return -1;
        LLZ_CUDA_TRY(cudaEventRecord(P.out_done[s], P.s_out));
        out_pos += outs;
    }
    LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_out));
    LLZ_CUDA_TRY(cudaStreamSynchronize(P.s_run));
    return 0;
}

extern "C" int llz_cuda_resample_bank_run_host(unsigned long handle, const short *h_in, long long in_stride,
                                               long long n_in, short *h_out, long long out_stride,
                                               long long *n_out)
{
    const int rc = poly_run_host_body(handle, h_in, in_stride, n_in, h_out, out_stride, n_out);
    PolyBank *b = reinterpret_cast<PolyBank *>(handle);
    if (handle == 0 || handle == kFail || b->magic != kMagicPoly) return rc;
    if (rc != 0 && b->pipe.ok) {                               // see llz_cuda_fir_bank_run_host
        DeviceGuard g(b->device);
        cudaStreamSynchronize(b->pipe.s_in);
        cudaStreamSynchronize(b->pipe.s_run);
        cudaStreamSynchronize(b->pipe.s_out);
        cudaGetLastError();
        b->poisoned = true;
    }
    b->trail.pending = false;
    return rc;
}

extern "C" int llz_cuda_resample_bank_set_tiles(unsigned long handle, int tiles)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (tiles < LLZ_CUDA_TILES_AUTO || tiles > LLZ_CUDA_TILES_INT8_TCGEN05) { llz_set_error("unknown tile family %d", tiles); return -1; }
    b->tiles = tiles;
    return 0;
}

extern "C" int llz_cuda_resample_bank_set_guard_scale(unsigned long handle, double scale)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (!(scale >= 1.0) || scale > 1e12) { llz_set_error("guard scale must be in [1, 1e12] (got %g)", scale); return -1; }
    b->guard_scale = scale;
    return 0;
}

extern "C" int llz_cuda_resample_bank_last_run(unsigned long handle, int *launches, char *kernel, int kernel_cap)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (launches) *launches = b->last_launches;
    if (kernel && kernel_cap > 0) snprintf(kernel, (size_t)kernel_cap, "%s", b->last_kernel);
    return 0;
}

extern "C" long long llz_cuda_resample_bank_guard_count(unsigned long handle)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    DeviceGuard g(b->device);
    unsigned long long v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return -1; }
    LLZ_CUDA_TRY(cudaMemcpy(&v, b->d_guard, sizeof v, cudaMemcpyDeviceToHost));
    return (long long)v;
}

// ====================================================================================================
// IIR banks and drop-in handles (llz_iir.h, llz_cuda.h): replaces llz_iir.c:38-156
// ====================================================================================================
namespace {

constexpr uint32_t kMagicIir = 0x4C5A4949u;                 // "LZII"

struct IirBank {
    uint32_t magic = kMagicIir;
    int device = 0;
    int M = 0, N = 0, n_channels = 1;
    double a[llz::kIirMaxOrder + 1] = {0}, b[llz::kIirMaxOrder + 1] = {0};
    double *d_state = nullptr;          // [channels][2 * kIirMaxOrder]
    // drop-in frames (one channel): page-locked staging + device frame, grown on demand; own stream
    double *pinned = nullptr, *d_frame = nullptr;
    int frame_cap = 0;
    cudaStream_t s_frame = nullptr;
};

IirBank *as_iir(unsigned long handle)
{
    if (handle == 0 || handle == kFail) { llz_set_error("invalid IIR handle"); return nullptr; }
    IirBank *b = reinterpret_cast<IirBank *>(handle);
    if (b->magic != kMagicIir) { llz_set_error("handle is not an IIR handle"); return nullptr; }
    return b;
}

void iir_destroy(IirBank *b)
{
    DeviceGuard g(b->device);
    cudaFree(b->d_state);
    cudaFree(b->d_frame);
    if (b->pinned) cudaFreeHost(b->pinned);
    if (b->s_frame) cudaStreamDestroy(b->s_frame);
    b->magic = 0;
    delete b;
}

int iir_run(IirBank *b, const double *d_x, long long x_stride, double *d_y, long long y_stride, long long n, cudaStream_t st)
{
    if (n < 0) { llz_set_error("negative sample count"); return -1; }
    if (n == 0) return 0;
    if (!d_y) { llz_set_error("null output pointer"); return -1; }
    llz::IirLaunch a{};
    a.M = b->M; a.N = b->N; a.n_channels = b->n_channels;
    memcpy(a.a, b->a, sizeof a.a);
    memcpy(a.b, b->b, sizeof a.b);
    a.x = d_x; a.x_stride = x_stride; a.y = d_y; a.y_stride = y_stride; a.n = n;
    a.state = b->d_state; a.state_stride = 2 * llz::kIirMaxOrder;
    return llz::iir_launch(a, st);
}

int iir_frame_room(IirBank *b, int n)
{
    if (n <= b->frame_cap) return 0;
    if (!b->s_frame) LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&b->s_frame, cudaStreamNonBlocking));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    if (b->pinned) cudaFreeHost(b->pinned);
    cudaFree(b->d_frame);
    b->pinned = b->d_frame = nullptr;
    b->frame_cap = 0;
    const int cap = n < 4096 ? 4096 : n;
    LLZ_CUDA_TRY(cudaHostAlloc(&b->pinned, sizeof(double) * (size_t)cap, cudaHostAllocDefault));
    LLZ_CUDA_TRY(cudaMalloc(&b->d_frame, sizeof(double) * (size_t)cap));
    b->frame_cap = cap;
    return 0;
}

}  // namespace

extern "C" unsigned long llz_cuda_iir_bank_init(int M, const double *a, int N, const double *b, int n_channels)
{
    if (M < 0 || N < 0 || M > llz::kIirMaxOrder || N > llz::kIirMaxOrder) {
        llz_set_error("IIR orders must be in [0, %d] (got M = %d, N = %d)", llz::kIirMaxOrder, M, N);
        return kFail;
    }
    if (!a && M > 0) { llz_set_error("null pole coefficients"); return kFail; }
    if (n_channels < 1 || n_channels > (1 << 24)) { llz_set_error("bad channel count %d", n_channels); return kFail; }
    int dev = 0;
    if (require_device(&dev) != 0) return kFail;
    IirBank *h = new (std::nothrow) IirBank();
    if (!h) { llz_set_error("out of memory"); return kFail; }
    h->device = dev; h->M = M; h->N = N; h->n_channels = n_channels;
    for (int i = 0; i <= M && a; ++i) h->a[i] = a[i];                      // a[0] is taken as 1 (llz_iir.c:52)
    for (int i = 0; i <= N && b; ++i) h->b[i] = b[i];                      // b == NULL: all zero (llz_iir.c:56-60)
    const size_t bytes = (size_t)n_channels * 2 * llz::kIirMaxOrder * sizeof(double);
    cudaError_t e;
    if ((e = cudaMalloc(&h->d_state, bytes)) != cudaSuccess || (e = cudaMemset(h->d_state, 0, bytes)) != cudaSuccess ||
        (e = cudaDeviceSynchronize()) != cudaSuccess) {
        llz_set_error("IIR bank init: %s", cudaGetErrorString(e));
        iir_destroy(h);
        return kFail;
    }
    return reinterpret_cast<unsigned long>(h);
}

extern "C" void llz_cuda_iir_bank_uninit(unsigned long handle)
{
    IirBank *b = as_iir(handle);
    if (b) iir_destroy(b);
}

extern "C" int llz_cuda_iir_bank_reset(unsigned long handle, llz_cuda_stream_t stream)
{
    IirBank *b = as_iir(handle);
    if (!b) return -1;
    DeviceGuard g(b->device);
    LLZ_CUDA_TRY(cudaMemsetAsync(b->d_state, 0, (size_t)b->n_channels * 2 * llz::kIirMaxOrder * sizeof(double), (cudaStream_t)stream));
    return 0;
}

extern "C" int llz_cuda_iir_bank_run(unsigned long handle, const double *d_x, long long x_stride, double *d_y,
                                     long long y_stride, long long n, llz_cuda_stream_t stream)
{
    IirBank *b = as_iir(handle);
    if (!b) return -1;
    DeviceGuard g(b->device);
    return iir_run(b, d_x, x_stride, d_y, y_stride, n, (cudaStream_t)stream);
}

extern "C" unsigned long llz_iir_filter_init(int M, double *a, int N, double *b)
{
    return llz_cuda_iir_bank_init(M, a, N, b, 1);
}

extern "C" void llz_iir_filter_uninit(unsigned long handle)
{
    llz_cuda_iir_bank_uninit(handle);
}

extern "C" int llz_iir_filter(unsigned long handle, double *x, double *y, int frame_len)
{
    IirBank *b = as_iir(handle);
    if (!b) return -1;
    if (b->n_channels != 1) { llz_set_error("llz_iir_filter needs a handle from llz_iir_filter_init"); return -1; }
    if (frame_len < 0 || (frame_len > 0 && (!x || !y))) { llz_set_error("llz_iir_filter: bad arguments"); return -1; }
    if (frame_len == 0) return 0;
    DeviceGuard g(b->device);
    if (iir_frame_room(b, frame_len) != 0) return -1;
    const size_t bytes = sizeof(double) * (size_t)frame_len;
    memcpy(b->pinned, x, bytes);
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->d_frame, b->pinned, bytes, cudaMemcpyHostToDevice, b->s_frame));
    if (iir_run(b, b->d_frame, 0, b->d_frame, 0, frame_len, b->s_frame) != 0) return -1;      // in place: the tile is staged
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->pinned, b->d_frame, bytes, cudaMemcpyDeviceToHost, b->s_frame));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    memcpy(y, b->pinned, bytes);
    return frame_len;                                                       // llz_iir.c:143
}

extern "C" int llz_iir_filter_flush(unsigned long handle, double *y)
{
    IirBank *b = as_iir(handle);
    if (!b) return -1;
    if (b->n_channels != 1) { llz_set_error("llz_iir_filter_flush needs a handle from llz_iir_filter_init"); return -1; }
    if (b->N == 0) return 0;
    if (!y) { llz_set_error("llz_iir_filter_flush: null output"); return -1; }
    DeviceGuard g(b->device);
    if (iir_frame_room(b, b->N) != 0) return -1;
    if (iir_run(b, nullptr, 0, b->d_frame, 0, b->N, b->s_frame) != 0) return -1;               // N zeros in (llz_iir.c:152-153)
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->pinned, b->d_frame, sizeof(double) * (size_t)b->N, cudaMemcpyDeviceToHost, b->s_frame));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    memcpy(y, b->pinned, sizeof(double) * (size_t)b->N);
    return b->N;                                                            // llz_iir.c:155
}

// ====================================================================================================
// drop-in FIR handles (llz_fir.h): replaces llz_fir.c:442-625
// ====================================================================================================
namespace {

unsigned long fir_dropin_init(int kind, int frame_len, int flt_len, double fc1, double fc2, win_t win)
{
    if (frame_len < 1) { llz_set_error("frame_len must be positive (got %d)", frame_len); return kFail; }
    // the reference accumulates with separate multiply and add in tap order; the strict mode does the
    // same on the GPU, so the drop-in output is bit-identical (PCIe, not the FP64 pipe, bounds a mono handle)
    unsigned long h = llz_cuda_fir_bank_init(kind, flt_len, fc1, fc2, win, 1, LLZ_CUDA_F64_STRICT);
    if (h == kFail) return kFail;
    FirBank *b = reinterpret_cast<FirBank *>(h);
    DeviceGuard g(b->device);
    b->frame_len = frame_len;
    const int cap = frame_len > b->hist_len ? frame_len : b->hist_len;     // flush emits flt_len-1 samples
    cudaError_t e;
    if ((e = cudaHostAlloc(&b->pinned, sizeof(double) * (size_t)cap, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaMalloc(&b->d_frame_in, sizeof(double) * (size_t)cap)) != cudaSuccess ||
        (e = cudaMalloc(&b->d_frame_in2, sizeof(double) * (size_t)cap)) != cudaSuccess ||
        (e = cudaMalloc(&b->d_frame_out, sizeof(double) * (size_t)cap)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&b->s_frame, cudaStreamNonBlocking)) != cudaSuccess) {
        llz_set_error("FIR handle init: %s", cudaGetErrorString(e));
        fir_destroy(b);
        return kFail;
    }
    return h;
}

}  // namespace

extern "C" unsigned long llz_fir_filter_lpf_init(int frame_len, int flt_len, double fc, win_t win_type)
{
    return fir_dropin_init(LLZ_CUDA_LPF, frame_len, flt_len, fc, 0.0, win_type);
}

extern "C" unsigned long llz_fir_filter_hpf_init(int frame_len, int flt_len, double fc, win_t win_type)
{
    return fir_dropin_init(LLZ_CUDA_HPF, frame_len, flt_len, fc, 0.0, win_type);
}

extern "C" unsigned long llz_fir_filter_bandpass_init(int frame_len, int flt_len, double fc1, double fc2,
                                                      win_t win_type)
{
    return fir_dropin_init(LLZ_CUDA_BPF, frame_len, flt_len, fc1, fc2, win_type);
}

extern "C" unsigned long llz_fir_filter_bandstop_init(int frame_len, int flt_len, double fc1, double fc2,
                                                      win_t win_type)
{
    return fir_dropin_init(LLZ_CUDA_BSF, frame_len, flt_len, fc1, fc2, win_type);
}

extern "C" void llz_fir_filter_uninit(unsigned long handle)
{
    FirBank *b = as_fir(handle);
    if (b) fir_destroy(b);
}

extern "C" int llz_fir_filter(unsigned long handle, double *buf_in, double *buf_out, int frame_len)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (!b->pinned) { llz_set_error("llz_fir_filter needs a handle from llz_fir_filter_*_init"); return -1; }
    if (frame_len < 0 || frame_len > b->frame_len) {           // the reference asserts (llz_fir.c:559)
        llz_set_error("frame_len %d exceeds the handle's %d", frame_len, b->frame_len);
        return -1;
    }
    if (frame_len == 0) return 0;
    DeviceGuard g(b->device);
    const size_t bytes = sizeof(double) * (size_t)frame_len;
    memcpy(b->pinned, buf_in, bytes);
    // frames alternate between two device buffers: the previous frame's tail stays readable as this frame's history
    void *d_frame = b->frame_flip ? b->d_frame_in2 : b->d_frame_in;
    b->frame_flip ^= 1;
    LLZ_CUDA_TRY(cudaMemcpyAsync(d_frame, b->pinned, bytes, cudaMemcpyHostToDevice, b->s_frame));
    if (fir_run(b, d_frame, 0, b->d_frame_out, 0, frame_len, b->s_frame, /*defer_history=*/true) != 0) return -1;
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->pinned, b->d_frame_out, bytes, cudaMemcpyDeviceToHost, b->s_frame));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    b->trail.pending = false;
    memcpy(buf_out, b->pinned, bytes);
    return frame_len;                                           // llz_fir.c:582
}

extern "C" int llz_fir_filter_flush(unsigned long handle, double *buf_out)
{
    FirBank *b = as_fir(handle);
    if (!b) return -1;
    if (!b->pinned) { llz_set_error("llz_fir_filter_flush needs a handle from llz_fir_filter_*_init"); return -1; }
    if (b->hist_len == 0) return 0;
    DeviceGuard g(b->device);
    const size_t bytes = sizeof(double) * (size_t)b->hist_len;
    if (fir_run(b, nullptr, 0, b->d_frame_out, 0, b->hist_len, b->s_frame) != 0) return -1;
    b->hist_zero = true;
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->pinned, b->d_frame_out, bytes, cudaMemcpyDeviceToHost, b->s_frame));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    b->trail.pending = false;
    memcpy(buf_out, b->pinned, bytes);
    return b->hist_len;                                         // llz_fir.c:624
}

// ====================================================================================================
// drop-in resampler handles (llz_resample.h): replaces llz_resample.c:271-617
// ====================================================================================================
namespace {

unsigned long poly_dropin_init(int kind, int L, int M, double gain, win_t win)
{
    // FP64 FMA + near-integer guard: bit-identical int16 (see llz_cuda_resample.cu)
    unsigned long h = poly_bank_create(kind, L, M, gain, win, 0, 1, LLZ_CUDA_ACC_F64);
    if (h == kFail) return kFail;
    PolyBank *b = reinterpret_cast<PolyBank *>(h);
    DeviceGuard g(b->device);
    const size_t in_b = sizeof(int16_t) * (size_t)b->plan.num_in, out_b = sizeof(int16_t) * (size_t)b->plan.num_out;
    cudaError_t e;
    if ((e = cudaHostAlloc((void **)&b->pinned_in, in_b, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&b->pinned_out, out_b, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaMalloc((void **)&b->d_frame_in, in_b)) != cudaSuccess ||
        (e = cudaMalloc((void **)&b->d_frame_in2, in_b)) != cudaSuccess ||
        (e = cudaMalloc((void **)&b->d_frame_out, out_b)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&b->s_frame, cudaStreamNonBlocking)) != cudaSuccess) {
        llz_set_error("resampler handle init: %s", cudaGetErrorString(e));
        poly_destroy(b);
        return kFail;
    }
    return h;
}

int poly_dropin_frame(unsigned long handle, int kind, unsigned char *sample_in, int sample_in_size,
                      unsigned char *sample_out, int *sample_out_size)
{
    PolyBank *b = as_poly(handle);
    if (!b) return -1;
    if (!b->pinned_in) { llz_set_error("handle was not created by a llz_*_init drop-in"); return -1; }
    if (b->plan.kind != kind) { llz_set_error("handle kind %d used with entry point of kind %d", b->plan.kind, kind); return -1; }
    const int bytes_in = b->plan.num_in * 2, bytes_out = b->plan.num_out * 2;
    if (sample_in_size != bytes_in) {                           // the reference asserts (llz_resample.c:443,507,560)
        llz_set_error("frame is %d bytes, handle expects %d", sample_in_size, bytes_in);
        return -1;
    }
    DeviceGuard g(b->device);
    memcpy(b->pinned_in, sample_in, (size_t)bytes_in);
    // frames alternate between two device buffers: the previous frame's tail stays readable as this frame's history
    int16_t *d_frame = b->frame_flip ? b->d_frame_in2 : b->d_frame_in;
    b->frame_flip ^= 1;
    LLZ_CUDA_TRY(cudaMemcpyAsync(d_frame, b->pinned_in, (size_t)bytes_in, cudaMemcpyHostToDevice, b->s_frame));
    long long outs = 0;
    if (poly_run(b, d_frame, 0, b->plan.num_in, b->d_frame_out, 0, &outs, b->s_frame, /*defer_history=*/true) != 0) return -1;
    if (outs != b->plan.num_out) { llz_set_error("internal: frame produced %lld samples, expected %d", outs, b->plan.num_out); return -1; }
    LLZ_CUDA_TRY(cudaMemcpyAsync(b->pinned_out, b->d_frame_out, (size_t)bytes_out, cudaMemcpyDeviceToHost, b->s_frame));
    LLZ_CUDA_TRY(cudaStreamSynchronize(b->s_frame));
    b->trail.pending = false;
    memcpy(sample_out, b->pinned_out, (size_t)bytes_out);
    if (sample_out_size) *sample_out_size = bytes_out;          // llz_resample.c:605
    return 0;
}

}  // namespace

extern "C" unsigned long llz_decimate_init(int M, double gain, win_t win_type)
{
    return poly_dropin_init(LLZ_KIND_DECIMATE, 1, M, gain, win_type);
}

extern "C" unsigned long llz_interp_init(int L, double gain, win_t win_type)
{
    return poly_dropin_init(LLZ_KIND_INTERP, L, 1, gain, win_type);
}

extern "C" unsigned long llz_resample_filter_init(int L, int M, double gain, win_t win_type)
{
    return poly_dropin_init(LLZ_KIND_RESAMPLE, L, M, gain, win_type);
}

// any *_uninit accepts any resampler handle (the reference CLI calls llz_resample_filter_uninit on all
// three kinds, main.c:125)
extern "C" void llz_decimate_uninit(unsigned long handle) { llz_cuda_resample_bank_uninit(handle); }
extern "C" void llz_interp_uninit(unsigned long handle) { llz_cuda_resample_bank_uninit(handle); }
extern "C" void llz_resample_filter_uninit(unsigned long handle) { llz_cuda_resample_bank_uninit(handle); }

extern "C" int llz_get_resample_framelen_bytes(unsigned long handle)
{
    PolyBank *b = as_poly(handle);
    return b ? b->plan.num_in * 2 : -1;                         // llz_resample.c:612-617
}

extern "C" int llz_decimate(unsigned long handle, unsigned char *sample_in, int sample_in_size,
                            unsigned char *sample_out, int *sample_out_size)
{
    return poly_dropin_frame(handle, LLZ_KIND_DECIMATE, sample_in, sample_in_size, sample_out, sample_out_size);
}

extern "C" int llz_interp(unsigned long handle, unsigned char *sample_in, int sample_in_size,
                          unsigned char *sample_out, int *sample_out_size)
{
    return poly_dropin_frame(handle, LLZ_KIND_INTERP, sample_in, sample_in_size, sample_out, sample_out_size);
}

extern "C" int llz_resample(unsigned long handle, unsigned char *sample_in, int sample_in_size,
                            unsigned char *sample_out, int *sample_out_size)
{
    return poly_dropin_frame(handle, LLZ_KIND_RESAMPLE, sample_in, sample_in_size, sample_out, sample_out_size);
}

// ====================================================================================================
// multi-GPU contexts and sharded jobs (same translation unit: they drive the banks' channel-group entry points)
// ====================================================================================================
#include "llz_mgpu.inl"
