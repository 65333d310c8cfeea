// llz_cuda_pcm.cu -- interleaved PCM frames <-> planar channels on the GPU.
//
// The reference treats every WAV as one mono stream whatever fmt.channels says (quirk R7: main.c:60-62,
// llz_wavfmt.c:82-150), so a stereo file is filtered across channels.  Real PCM is interleaved
// (frame f = C consecutive samples); the filter banks of this library are planar (channel c at
// base + c*stride).  These two kernels are the step either side of the hot path: pure data movement,
// HBM-bound, fused with the sample-format conversion so the data crosses HBM once.
//
// deinterleave: a CTA stages TF frames x C channels in shared memory -- global reads are the contiguous
// interleaved bytes (16-byte vectors when aligned) -- and writes each channel's TF samples as one contiguous
// run (16-byte vectors when aligned).  interleave is the mirror image.
//
// Conversions (exact, defined here because the reference has no multi-format path):
//   s16 -> f32/f64: x / 32768          s24 -> f32/f64: x / 8388608        f32 -> f32/f64: x
//   s16 -> s16: copy                   s24 -> s16: arithmetic shift by 8   f32 -> s16: trunc(clamp(x * 32768))
//   and back (interleave): f32/f64 -> s16: trunc(clamp(x * 32768)) (the reference's saturate + C cast,
//   llz_resample.c:596-601), -> s24 likewise at 24 bits, -> f32: (float)x.
#include "llz_cuda_common.cuh"

namespace llz {

namespace {

constexpr int kPcmThreads = 256;

template <int FMT> struct PcmIn;                                   // interleaved sample readers
template <> struct PcmIn<LLZ_CUDA_PCM_S16> {
    static constexpr int BYTES = 2;
    static __device__ __forceinline__ float to_unit_scale() { return 1.0f / 32768.0f; }
    static __device__ __forceinline__ int load_int(const unsigned char *p) { return *reinterpret_cast<const int16_t *>(p); }
};
template <> struct PcmIn<LLZ_CUDA_PCM_S24> {
    static constexpr int BYTES = 3;
    static __device__ __forceinline__ int load_int(const unsigned char *p)
    {
        return (int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)(int32_t)(int8_t)p[2] << 16));
    }
};
template <> struct PcmIn<LLZ_CUDA_PCM_F32> { static constexpr int BYTES = 4; };

// interleaved sample at p -> planar element type TO
template <int FMT, typename TO>
__device__ __forceinline__ TO pcm_to_planar(const unsigned char *p)
{
    if constexpr (FMT == LLZ_CUDA_PCM_F32) {
        const float v = *reinterpret_cast<const float *>(p);
        if constexpr (sizeof(TO) == 2) {
            float s = v * 32768.0f;
            s = fminf(fmaxf(s, -32768.0f), 32767.0f);
            return (TO)__float2int_rz(s);
        } else {
            return (TO)v;
        }
    } else {
        const int x = PcmIn<FMT>::load_int(p);
        if constexpr (sizeof(TO) == 2) return (TO)(FMT == LLZ_CUDA_PCM_S24 ? (x >> 8) : x);
        else return (TO)x * (TO)(FMT == LLZ_CUDA_PCM_S24 ? 1.0 / 8388608.0 : 1.0 / 32768.0);
    }
}

// planar element -> interleaved sample at p
template <int FMT, typename TI>
__device__ __forceinline__ void planar_to_pcm(TI v, unsigned char *p)
{
    if constexpr (FMT == LLZ_CUDA_PCM_F32) {
        if constexpr (sizeof(TI) == 2) *reinterpret_cast<float *>(p) = (float)v * (1.0f / 32768.0f);
        else *reinterpret_cast<float *>(p) = (float)v;
    } else {
        int x;
        if constexpr (sizeof(TI) == 2) {
            x = (FMT == LLZ_CUDA_PCM_S24) ? ((int)v << 8) : (int)v;
        } else {
            const double full = (FMT == LLZ_CUDA_PCM_S24) ? 8388608.0 : 32768.0;
            double s = (double)v * full;
            s = fmin(fmax(s, -full), full - 1.0);
            x = __double2int_rz(s);
        }
        if constexpr (FMT == LLZ_CUDA_PCM_S16) {
            *reinterpret_cast<int16_t *>(p) = (int16_t)x;
        } else {
            p[0] = (unsigned char)x; p[1] = (unsigned char)(x >> 8); p[2] = (unsigned char)(x >> 16);
        }
    }
}

// shared tile [C][tf_pad] of TP elements; tf_pad chosen so that consecutive channels land in different banks
template <int FMT, typename TP, bool TO_PLANAR>
__global__ void __launch_bounds__(kPcmThreads)
pcm_transpose_kernel(unsigned char *frames, TP *planar, long long planar_stride, int C, long long n_frames, int TF,
                     int tf_pad)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TP *tile = reinterpret_cast<TP *>(smem_raw);
    constexpr int B = PcmIn<FMT>::BYTES;
    constexpr int VP = 16 / (int)sizeof(TP);                       // planar elements per 16-byte vector
    const long long f0 = (long long)blockIdx.x * TF;
    const int nf = (int)min((long long)TF, n_frames - f0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *fbase = frames + (size_t)f0 * C * B;
    const int total = nf * C;                                      // interleaved samples in this tile

    // Element (c, f) lives in row c at column f with its 16-byte vector index XORed by (c / 8) % 8: a warp walking
    // the interleaved order touches channels 8 apart in one instruction, which this spreads over all banks, while
    // a row stays a permutation of whole vectors for the planar side.
    auto col = [&](int c, int f) { return (((f / VP) ^ ((c >> 3) & 7)) * VP) + (f % VP); };
    auto planar_phase = [&]() {                                    // tile <-> planar rows, one warp per channel
        for (int c = warp; c < C; c += kPcmThreads / 32) {
            TP *row = planar + (long long)c * planar_stride + f0;
            TP *trow = tile + (size_t)c * tf_pad;
            const bool vec = (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
            const int nv = vec ? nf / VP : 0;
            const int key = (c >> 3) & 7;
            if (TO_PLANAR) {
                for (int v = lane; v < nv; v += 32)
                    reinterpret_cast<uint4 *>(row)[v] = reinterpret_cast<const uint4 *>(trow)[v ^ key];
            } else {
                // global -> shared: four independent 16-byte loads in flight per lane before the first store
                int v = lane;
                for (; v + 96 < nv; v += 128) {
                    uint4 w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) w[u] = reinterpret_cast<const uint4 *>(row)[v + 32 * u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) reinterpret_cast<uint4 *>(trow)[(v + 32 * u) ^ key] = w[u];
                }
                for (; v < nv; v += 32) reinterpret_cast<uint4 *>(trow)[v ^ key] = reinterpret_cast<const uint4 *>(row)[v];
            }
            for (int f = nv * VP + lane; f < nf; f += 32) {
                if (TO_PLANAR) row[f] = trow[col(c, f)];
                else trow[col(c, f)] = row[f];
            }
        }
    };
    auto frame_phase = [&]() {                                     // tile <-> interleaved frames, contiguous in memory
        constexpr int EPV = (B == 3) ? 0 : 16 / B;                 // samples per 16-byte vector (s24: scalar path)
        int done = 0;
        if (EPV > 0 && (reinterpret_cast<uintptr_t>(fbase) & 15u) == 0) {
            const int nvec = total / (EPV ? EPV : 1);
            for (int v = tid; v < nvec; v += kPcmThreads) {
                uint4 w;
                if (TO_PLANAR) w = reinterpret_cast<const uint4 *>(fbase)[v];
                unsigned char *wb = reinterpret_cast<unsigned char *>(&w);
                int idx = v * EPV;
                int f = idx / C, c = idx - f * C;
#pragma unroll
                for (int i = 0; i < EPV; ++i) {
                    if (TO_PLANAR) tile[(size_t)c * tf_pad + col(c, f)] = pcm_to_planar<FMT, TP>(wb + i * B);
                    else planar_to_pcm<FMT, TP>(tile[(size_t)c * tf_pad + col(c, f)], wb + i * B);
                    if (++c == C) { c = 0; ++f; }
                }
                if (!TO_PLANAR) reinterpret_cast<uint4 *>(fbase)[v] = w;
            }
            done = nvec * EPV;
        }
        for (int idx = done + tid; idx < total; idx += kPcmThreads) {
            const int f = idx / C, c = idx - f * C;
            if (TO_PLANAR) tile[(size_t)c * tf_pad + col(c, f)] = pcm_to_planar<FMT, TP>(fbase + (size_t)idx * B);
            else planar_to_pcm<FMT, TP>(tile[(size_t)c * tf_pad + col(c, f)], fbase + (size_t)idx * B);
        }
    };

    if (TO_PLANAR) { frame_phase(); __syncthreads(); planar_phase(); }
    else { planar_phase(); __syncthreads(); frame_phase(); }
}

template <int FMT, typename TP, bool TO_PLANAR>
int launch_pcm(unsigned char *frames, TP *planar, long long stride, int C, long long n_frames, cudaStream_t st)
{
    // frames per tile: ~48 KiB of planar elements, a multiple of 64, at least 64
    long long tf = (48 * 1024) / ((long long)C * (long long)sizeof(TP));
    tf = tf / 64 * 64;
    if (tf < 64) tf = 64;
    if (tf > 4096) tf = 4096;
    // row pitch: a multiple of 16 bytes (vector access) plus 16 bytes so consecutive channels start 4 banks apart
    const int tf_pad = (int)tf + 16 / (int)sizeof(TP);
    const size_t smem = (size_t)C * tf_pad * sizeof(TP);
    if (smem > 200 * 1024) { llz_set_error("pcm: %d channels do not fit a shared-memory tile", C); return -1; }
    auto kern = pcm_transpose_kernel<FMT, TP, TO_PLANAR>;
    if (smem > 48 * 1024)
        LLZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (n_frames + tf - 1) / tf;
    if (tiles > 0x7fffffffLL) { llz_set_error("pcm: launch too large"); return -1; }
    kern<<<(unsigned)tiles, kPcmThreads, smem, st>>>(frames, planar, stride, C, n_frames, (int)tf, tf_pad);
    LLZ_CUDA_TRY(cudaGetLastError());
    return 0;
}

template <bool TO_PLANAR>
int dispatch_pcm(void *frames, int fmt, int C, long long n_frames, void *planar, int ptype, long long stride, cudaStream_t st)
{
    if (C < 1 || n_frames < 0 || !frames || !planar) { llz_set_error("pcm: bad arguments"); return -1; }
    if (n_frames == 0) return 0;
    unsigned char *fr = static_cast<unsigned char *>(frames);
#define LLZ_PCM_CASE(F, T) return launch_pcm<F, T, TO_PLANAR>(fr, static_cast<T *>(planar), stride, C, n_frames, st)
    switch (fmt * 4 + ptype) {
    case LLZ_CUDA_PCM_S16 * 4 + LLZ_CUDA_PLANAR_S16: LLZ_PCM_CASE(LLZ_CUDA_PCM_S16, int16_t);
    case LLZ_CUDA_PCM_S16 * 4 + LLZ_CUDA_PLANAR_F32: LLZ_PCM_CASE(LLZ_CUDA_PCM_S16, float);
    case LLZ_CUDA_PCM_S16 * 4 + LLZ_CUDA_PLANAR_F64: LLZ_PCM_CASE(LLZ_CUDA_PCM_S16, double);
    case LLZ_CUDA_PCM_S24 * 4 + LLZ_CUDA_PLANAR_S16: LLZ_PCM_CASE(LLZ_CUDA_PCM_S24, int16_t);
    case LLZ_CUDA_PCM_S24 * 4 + LLZ_CUDA_PLANAR_F32: LLZ_PCM_CASE(LLZ_CUDA_PCM_S24, float);
    case LLZ_CUDA_PCM_S24 * 4 + LLZ_CUDA_PLANAR_F64: LLZ_PCM_CASE(LLZ_CUDA_PCM_S24, double);
    case LLZ_CUDA_PCM_F32 * 4 + LLZ_CUDA_PLANAR_S16: LLZ_PCM_CASE(LLZ_CUDA_PCM_F32, int16_t);
    case LLZ_CUDA_PCM_F32 * 4 + LLZ_CUDA_PLANAR_F32: LLZ_PCM_CASE(LLZ_CUDA_PCM_F32, float);
    case LLZ_CUDA_PCM_F32 * 4 + LLZ_CUDA_PLANAR_F64: LLZ_PCM_CASE(LLZ_CUDA_PCM_F32, double);
    default: llz_set_error("pcm: unknown format %d / planar type %d", fmt, ptype); return -1;
    }
#undef LLZ_PCM_CASE
}

}  // namespace

}  // namespace llz

extern "C" int llz_cuda_pcm_deinterleave(const void *d_frames, int pcm_format, int n_channels, long long n_frames,
                                         void *d_planar, int planar_type, long long planar_stride,
                                         llz_cuda_stream_t stream)
{
    return llz::dispatch_pcm<true>(const_cast<void *>(d_frames), pcm_format, n_channels, n_frames, d_planar, planar_type,
                                   planar_stride, (cudaStream_t)stream);
}

extern "C" int llz_cuda_pcm_interleave(const void *d_planar, int planar_type, long long planar_stride, int n_channels,
                                       long long n_frames, void *d_frames, int pcm_format, llz_cuda_stream_t stream)
{
    return llz::dispatch_pcm<false>(d_frames, pcm_format, n_channels, n_frames, const_cast<void *>(d_planar), planar_type,
                                    planar_stride, (cudaStream_t)stream);
}
