// llz_mgpu.inl -- multi-GPU jobs of libllzfilter_cuda (include/llz_cuda.h, "Multi-GPU"); part of llz_shim.cu's translation
// unit because it drives the banks through their channel-group entry points (fir_run_part / poly_run_part).
//
// The reference has one mono stream per handle (libllzfilter/llz_fir.c:23-33, llz_resample.c:54-78) and no notion of
// several devices.  What it does have is frame streaming with a history prefix (llz_fir.c:561-566,
// llz_resample.c:570-576): a frame's outputs depend on the frame and on the last flt_len-1 / Q-1 samples before it.  A
// time segment with that prefix as halo is therefore an exact restatement of the reference's own frame, and independent
// channels are independent handles -- the two shardings of SURVEY.md section 8(e).  Neither needs an exchange during
// the computation; the only collective is the optional gather of the result on one device.
//
// Process models: one process driving n GPUs (llz_cuda_mgpu_init_all: ncclCommInitAll + peer access) or one process
// per GPU (llz_cuda_mgpu_init_rank with a shared ncclUniqueId: torchrun, MPI).  A context has `nlocal` local slots
// (n or 1); every per-slot argument of a job call is an array of nlocal entries.
//
// Gather modes (root = the rank that owns the result buffer):
//   LLZ_CUDA_GATHER_PEER  the kernels of every rank store their outputs straight into the root's buffer through the
//                         NVLink peer mapping (cudaDeviceEnablePeerAccess / cudaIpcOpenMemHandle): compute and
//                         collective are ONE kernel, the transfer overlaps the arithmetic tile by tile, and nothing is
//                         staged; a one-int ncclAllReduce at the end orders "every rank has finished" on every stream.
//   LLZ_CUDA_GATHER_NCCL  each rank computes its shard in chunks (groups of whole channels / runs of whole work items)
//                         into its own buffer; chunk c travels to the root by grouped ncclSend / ncclRecv on a second
//                         stream while chunk c+1 is computed.  The root's own shard is computed in place.
//   LLZ_CUDA_GATHER_COPY  chunks like GATHER_NCCL, but chunk c is PUSHED into the root's buffer through the peer mapping by
//                         the rank's copy engine (cudaMemcpyAsync on the second stream): no SM is needed for the
//                         transfer, so it overlaps the persistent tcgen05 / FFT kernels, which leave no room for an
//                         NCCL kernel beside them (measured at N = 2: compute + NCCL gather = the SUM of the two for
//                         the resampler configs).
// NCCL is loaded at run time (dlopen "libnccl.so.2": inside a PyTorch process that is the copy torch already loaded),
// so the single-GPU drop-in path has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *so = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

// loaded once; nullptr + error message when the library or a symbol is missing (no fallback: the call fails)
const NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        api.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.so) api.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.so) {
            ok = true;
#define LLZ_NCCL_SYM(field, name)                                                  \
    do {                                                                           \
        *(void **)(&api.field) = dlsym(api.so, name);                              \
        if (!api.field) ok = false;                                                \
    } while (0)
            LLZ_NCCL_SYM(GetVersion, "ncclGetVersion");
            LLZ_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            LLZ_NCCL_SYM(CommInitRank, "ncclCommInitRank");
            LLZ_NCCL_SYM(CommInitAll, "ncclCommInitAll");
            LLZ_NCCL_SYM(CommDestroy, "ncclCommDestroy");
            LLZ_NCCL_SYM(Send, "ncclSend");
            LLZ_NCCL_SYM(Recv, "ncclRecv");
            LLZ_NCCL_SYM(AllReduce, "ncclAllReduce");
            LLZ_NCCL_SYM(Broadcast, "ncclBroadcast");
            LLZ_NCCL_SYM(GroupStart, "ncclGroupStart");
            LLZ_NCCL_SYM(GroupEnd, "ncclGroupEnd");
            LLZ_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef LLZ_NCCL_SYM
        }
    }
    if (!ok) {
        llz_set_error("multi-GPU context needs NCCL: %s", api.so ? "libnccl.so.2 lacks a required symbol" : dlerror());
        return nullptr;
    }
    return &api;
}

#define LLZ_NCCL_TRY(api, expr)                                                                       \
    do {                                                                                              \
        ncclResult_t r__ = (expr);                                                                    \
        if (r__ != ncclSuccess) {                                                                     \
            llz_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(r__)); \
            return -1;                                                                                \
        }                                                                                             \
    } while (0)

constexpr uint32_t kMagicMgpu = 0x4C5A4D47u;                 // "LZMG"
constexpr uint32_t kMagicJob = 0x4C5A4A42u;                  // "LZJB"
constexpr int kMaxSlots = 16;
constexpr int kMaxChunks = 8;

struct MgpuSlot {
    int device = 0;
    int rank = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t s_comm = nullptr;                            // gather traffic of this slot
    cudaEvent_t ev_chunk[kMaxChunks] = {};                    // chunk c of the current run has been computed
    cudaEvent_t ev_comm = nullptr, ev_enter = nullptr;
    int *d_flag = nullptr;                                    // two ints: the completion all-reduce
    void *result_ptr = nullptr;                               // the root's result buffer as this slot addresses it
    bool result_ipc = false;                                  // result_ptr came from cudaIpcOpenMemHandle
};

struct MgpuCtx {
    uint32_t magic = kMagicMgpu;
    int world = 0;
    int nlocal = 0;
    bool single_process = false;
    MgpuSlot slot[kMaxSlots];
    int result_root = -1;
    size_t result_bytes = 0;
};

MgpuCtx *as_mgpu(unsigned long h)
{
    if (h == 0 || h == kFail) { llz_set_error("invalid multi-GPU context"); return nullptr; }
    MgpuCtx *c = reinterpret_cast<MgpuCtx *>(h);
    if (c->magic != kMagicMgpu) { llz_set_error("handle is not a multi-GPU context"); return nullptr; }
    return c;
}

int mgpu_slot_setup(MgpuSlot &s)
{
    DeviceGuard g(s.device);
    LLZ_CUDA_TRY(cudaStreamCreateWithFlags(&s.s_comm, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxChunks; ++i) LLZ_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_chunk[i], cudaEventDisableTiming));
    LLZ_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_comm, cudaEventDisableTiming));
    LLZ_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_enter, cudaEventDisableTiming));
    LLZ_CUDA_TRY(cudaMalloc(&s.d_flag, 2 * sizeof(int)));
    LLZ_CUDA_TRY(cudaMemset(s.d_flag, 0, 2 * sizeof(int)));
    LLZ_CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

int mgpu_result_release(MgpuCtx *c)
{
    for (int i = 0; i < c->nlocal; ++i) {
        MgpuSlot &s = c->slot[i];
        if (!s.result_ptr) continue;
        DeviceGuard g(s.device);
        cudaDeviceSynchronize();
        if (s.result_ipc) cudaIpcCloseMemHandle(s.result_ptr);
        else if (s.rank == c->result_root) cudaFree(s.result_ptr);
        s.result_ptr = nullptr;
        s.result_ipc = false;
    }
    c->result_root = -1;
    c->result_bytes = 0;
    cudaGetLastError();
    return 0;
}

void mgpu_destroy(MgpuCtx *c)
{
    if (!c) return;
    const NcclApi *api = nccl_api();
    mgpu_result_release(c);
    for (int i = 0; i < c->nlocal; ++i) {
        MgpuSlot &s = c->slot[i];
        DeviceGuard g(s.device);
        cudaDeviceSynchronize();
        if (s.comm && api) api->CommDestroy(s.comm);
        if (s.s_comm) cudaStreamDestroy(s.s_comm);
        for (int k = 0; k < kMaxChunks; ++k)
            if (s.ev_chunk[k]) cudaEventDestroy(s.ev_chunk[k]);
        if (s.ev_comm) cudaEventDestroy(s.ev_comm);
        if (s.ev_enter) cudaEventDestroy(s.ev_enter);
        if (s.d_flag) cudaFree(s.d_flag);
    }
    cudaGetLastError();
    c->magic = 0;
    delete c;
}

int mgpu_local_of_rank(const MgpuCtx *c, int rank)
{
    for (int i = 0; i < c->nlocal; ++i)
        if (c->slot[i].rank == rank) return i;
    return -1;
}

// ---- jobs ------------------------------------------------------------------------------------------------------------
struct MgpuJob {
    uint32_t magic = kMagicJob;
    MgpuCtx *ctx = nullptr;
    bool is_fir = true;
    int shard_mode = LLZ_CUDA_SHARD_CHANNEL;
    int n_channels = 0;                                       // of the whole job
    unsigned long bank[kMaxSlots] = {};                       // one bank per local slot, on the slot's device
    // FIR
    int flt_len = 0, dtype = LLZ_CUDA_F64;
    // resample
    int L = 1, M = 1, taps_per_phase = 0, frame_in = 0;
};

MgpuJob *as_job(unsigned long h)
{
    if (h == 0 || h == kFail) { llz_set_error("invalid multi-GPU job"); return nullptr; }
    MgpuJob *j = reinterpret_cast<MgpuJob *>(h);
    if (j->magic != kMagicJob) { llz_set_error("handle is not a multi-GPU job"); return nullptr; }
    return j;
}

void job_destroy(MgpuJob *j)
{
    if (!j) return;
    for (int i = 0; i < kMaxSlots; ++i) {
        if (!j->bank[i] || j->bank[i] == kFail) continue;
        if (j->is_fir) llz_cuda_fir_bank_uninit(j->bank[i]);
        else llz_cuda_resample_bank_uninit(j->bank[i]);
    }
    j->magic = 0;
    delete j;
}

size_t job_elem_size(const MgpuJob *j) { return j->is_fir ? fir_elem_size(j->dtype) : sizeof(int16_t); }

// what `rank` owns of a job over n_total input samples per channel (pure integer arithmetic: every rank can plan
// every other rank's shard, which is how the root knows what to receive)
int job_plan_rank(const MgpuJob *j, long long n_total, int rank, llz_cuda_shard_t *out)
{
    const int world = j->ctx->world;
    memset(out, 0, sizeof *out);
    if (j->shard_mode == LLZ_CUDA_SHARD_CHANNEL) {
        if (llz_cuda_shard_channels(j->n_channels, world, rank, &out->first_channel, &out->n_channels) != 0) return -1;
        out->seg.in_start = 0;
        out->seg.in_count = n_total;
        out->seg.halo = 0;
        out->seg.out_start = 0;
        if (j->is_fir) out->seg.out_count = n_total;
        else out->seg.out_count = (n_total * j->L + j->M - 1) / j->M;
        return 0;
    }
    out->first_channel = 0;
    out->n_channels = j->n_channels;
    if (j->is_fir) {
        const long long granule = llz_cuda_fir_bank_block_len(j->bank[0]);
        if (granule < 1) return -1;
        return llz_cuda_shard_fir_segments_aligned(n_total, j->flt_len, granule, world, rank, &out->seg);
    }
    return llz_cuda_shard_resample_segments(n_total, j->L, j->M, j->taps_per_phase, j->frame_in, world, rank, &out->seg);
}

long long job_total_out(const MgpuJob *j, long long n_total)
{
    return j->is_fir ? n_total : (n_total * j->L + j->M - 1) / j->M;
}

// chunk c of `chunks` of a shard: CHANNEL mode cuts the shard's channels, TIME mode its time axis (whole work items /
// whole frames, so that chunked and one-shot results are the same bytes).  Pure arithmetic, evaluated identically by
// the sender and by the root.
struct ChunkPlan {
    int c0, cc;                 // channels of the shard: [c0, c0 + cc)
    long long in0, in_len;      // input samples relative to the segment start
    long long out0, out_len;    // outputs relative to the segment's first output
};

ChunkPlan job_chunk(const MgpuJob *j, const llz_cuda_shard_t &sh, long long granule, int chunks, int c)
{
    ChunkPlan p{};
    if (j->shard_mode == LLZ_CUDA_SHARD_CHANNEL) {
        const int base = sh.n_channels / chunks, rem = sh.n_channels % chunks;
        p.c0 = c * base + (c < rem ? c : rem);
        p.cc = base + (c < rem ? 1 : 0);
        p.in0 = 0; p.in_len = sh.seg.in_count;
        p.out0 = 0; p.out_len = sh.seg.out_count;
        return p;
    }
    p.c0 = 0; p.cc = sh.n_channels;
    const long long units = (sh.seg.in_count + granule - 1) / granule;
    const long long base = units / chunks, rem = units % chunks;
    const long long u0 = c * base + (c < rem ? c : rem), uc = base + (c < rem ? 1 : 0);
    long long a = u0 * granule, b = (u0 + uc) * granule;
    if (a > sh.seg.in_count) a = sh.seg.in_count;
    if (b > sh.seg.in_count) b = sh.seg.in_count;
    p.in0 = a; p.in_len = b - a;
    if (j->is_fir) { p.out0 = a; p.out_len = b - a; }
    else {
        // segment starts at phase 0 and granule is whole frames: outputs of [0, t) input samples = ceil(t*L/M)
        p.out0 = (a * j->L + j->M - 1) / j->M;
        p.out_len = (b * j->L + j->M - 1) / j->M - p.out0;
    }
    return p;
}

long long job_granule(const MgpuJob *j)
{
    if (j->shard_mode == LLZ_CUDA_SHARD_CHANNEL) return 1;
    if (j->is_fir) return llz_cuda_fir_bank_block_len(j->bank[0]);
    return j->frame_in;
}

// run channels [c0, c0+cc) x inputs [in0, in0+in_len) of slot i's shard; `x` points at the shard's first owned sample
// of its first channel, `y` at where the shard's first output goes
int job_run_chunk(MgpuJob *j, int i, const ChunkPlan &p, const unsigned char *x, long long x_stride, unsigned char *y,
                  long long y_stride, cudaStream_t st, bool last_group)
{
    const size_t es = job_elem_size(j);
    if (p.cc <= 0 || p.in_len <= 0) return 0;
    if (j->is_fir) {
        FirBank *b = as_fir(j->bank[i]);
        if (!b) return -1;
        return fir_run_part(b, x + ((size_t)p.c0 * x_stride + p.in0) * es, x_stride, y + ((size_t)p.c0 * y_stride + p.out0) * es,
                            y_stride, p.in_len, st, p.c0, p.cc, last_group);
    }
    PolyBank *b = as_poly(j->bank[i]);
    if (!b) return -1;
    long long outs = 0;
    if (poly_run_part(b, reinterpret_cast<const int16_t *>(x + ((size_t)p.c0 * x_stride + p.in0) * es), x_stride, p.in_len,
                      reinterpret_cast<int16_t *>(y + ((size_t)p.c0 * y_stride + p.out0) * es), y_stride, &outs, st, p.c0,
                      p.cc, last_group) != 0)
        return -1;
    if (outs != p.out_len) {
        llz_set_error("internal: multi-GPU chunk produced %lld outputs, planned %lld", outs, p.out_len);
        return -1;
    }
    return 0;
}

}  // namespace

// ======================================================================================================================
// context
// ======================================================================================================================
extern "C" int llz_cuda_mgpu_unique_id(unsigned char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    const NcclApi *api = nccl_api();
    if (!api || !id) return -1;
    ncclUniqueId u;
    LLZ_NCCL_TRY(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return 0;
}

extern "C" unsigned long llz_cuda_mgpu_init_rank(const unsigned char id[128], int world, int rank)
{
    const NcclApi *api = nccl_api();
    if (!api) return kFail;
    if (!id || world < 1 || world > 4096 || rank < 0 || rank >= world) { llz_set_error("mgpu_init_rank: bad arguments"); return kFail; }
    int dev = 0;
    if (require_device(&dev) != 0) return kFail;
    MgpuCtx *c = new (std::nothrow) MgpuCtx();
    if (!c) { llz_set_error("out of memory"); return kFail; }
    c->world = world;
    c->nlocal = 1;
    c->single_process = false;
    c->slot[0].device = dev;
    c->slot[0].rank = rank;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = api->CommInitRank(&c->slot[0].comm, world, u, rank);
    if (r != ncclSuccess) {
        llz_set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        c->slot[0].comm = nullptr;
        mgpu_destroy(c);
        return kFail;
    }
    if (mgpu_slot_setup(c->slot[0]) != 0) { mgpu_destroy(c); return kFail; }
    return reinterpret_cast<unsigned long>(c);
}

extern "C" unsigned long llz_cuda_mgpu_init_all(int n_gpus, const int *devices)
{
    const NcclApi *api = nccl_api();
    if (!api) return kFail;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) { cudaGetLastError(); llz_set_error("no usable CUDA device"); return kFail; }
    if (n_gpus < 1 || n_gpus > kMaxSlots || n_gpus > count) {
        llz_set_error("mgpu_init_all: %d GPUs requested, %d visible (at most %d per context)", n_gpus, count, kMaxSlots);
        return kFail;
    }
    MgpuCtx *c = new (std::nothrow) MgpuCtx();
    if (!c) { llz_set_error("out of memory"); return kFail; }
    c->world = n_gpus;
    c->nlocal = n_gpus;
    c->single_process = true;
    int devs[kMaxSlots];
    ncclComm_t comms[kMaxSlots];
    for (int i = 0; i < n_gpus; ++i) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= count) { llz_set_error("mgpu_init_all: bad device %d", devs[i]); delete c; return kFail; }
        c->slot[i].device = devs[i];
        c->slot[i].rank = i;
    }
    ncclResult_t r = api->CommInitAll(comms, n_gpus, devs);
    if (r != ncclSuccess) { llz_set_error("ncclCommInitAll failed: %s", api->GetErrorString(r)); delete c; return kFail; }
    for (int i = 0; i < n_gpus; ++i) c->slot[i].comm = comms[i];
    for (int i = 0; i < n_gpus; ++i) {
        if (mgpu_slot_setup(c->slot[i]) != 0) { mgpu_destroy(c); return kFail; }
        // peer mappings for the PEER gather and the result buffer (already-enabled is fine)
        DeviceGuard g(devs[i]);
        for (int k = 0; k < n_gpus; ++k) {
            if (k == i) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[k]);
            if (can) cudaDeviceEnablePeerAccess(devs[k], 0);
            cudaGetLastError();
        }
    }
    return reinterpret_cast<unsigned long>(c);
}

extern "C" void llz_cuda_mgpu_uninit(unsigned long ctx)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (c) mgpu_destroy(c);
}

extern "C" int llz_cuda_mgpu_world(unsigned long ctx) { MgpuCtx *c = as_mgpu(ctx); return c ? c->world : -1; }
extern "C" int llz_cuda_mgpu_local_count(unsigned long ctx) { MgpuCtx *c = as_mgpu(ctx); return c ? c->nlocal : -1; }
extern "C" int llz_cuda_mgpu_local_rank(unsigned long ctx, int i)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c || i < 0 || i >= c->nlocal) return -1;
    return c->slot[i].rank;
}
extern "C" int llz_cuda_mgpu_local_device(unsigned long ctx, int i)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c || i < 0 || i >= c->nlocal) return -1;
    return c->slot[i].device;
}

// collective over the context: every process calls it with the same arguments
extern "C" int llz_cuda_mgpu_result_alloc(unsigned long ctx, int root, size_t bytes)
{
    MgpuCtx *c = as_mgpu(ctx);
    const NcclApi *api = nccl_api();
    if (!c || !api) return -1;
    if (root < 0 || root >= c->world || bytes == 0) { llz_set_error("mgpu_result_alloc: bad arguments"); return -1; }
    mgpu_result_release(c);
    c->result_root = root;
    c->result_bytes = bytes;
    const int li = mgpu_local_of_rank(c, root);
    void *base = nullptr;
    if (li >= 0) {
        DeviceGuard g(c->slot[li].device);
        LLZ_CUDA_TRY(cudaMalloc(&base, bytes));
        c->slot[li].result_ptr = base;
    }
    if (c->single_process) {
        // one address space: the root's allocation is addressable from every device with peer access enabled
        for (int i = 0; i < c->nlocal; ++i) {
            if (i == li) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->slot[i].device, c->slot[li].device);
            cudaGetLastError();
            c->slot[i].result_ptr = can ? base : nullptr;      // no peer path: the PEER gather is refused, NCCL still works
        }
        return 0;
    }
    // one process per GPU: the root exports an IPC handle; it travels through the communicator itself
    MgpuSlot &s = c->slot[0];
    DeviceGuard g(s.device);
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof h);
    if (li >= 0) LLZ_CUDA_TRY(cudaIpcGetMemHandle(&h, base));
    void *d_h = nullptr;
    LLZ_CUDA_TRY(cudaMalloc(&d_h, sizeof h));
    LLZ_CUDA_TRY(cudaMemcpyAsync(d_h, &h, sizeof h, cudaMemcpyHostToDevice, s.s_comm));
    LLZ_NCCL_TRY(api, api->Broadcast(d_h, d_h, sizeof h, ncclChar, root, s.comm, s.s_comm));
    LLZ_CUDA_TRY(cudaMemcpyAsync(&h, d_h, sizeof h, cudaMemcpyDeviceToHost, s.s_comm));
    LLZ_CUDA_TRY(cudaStreamSynchronize(s.s_comm));
    cudaFree(d_h);
    if (li < 0) {
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) {
            s.result_ptr = p;
            s.result_ipc = true;
        } else {
            cudaGetLastError();                               // no peer path from this device: PEER gather refused later
            s.result_ptr = nullptr;
        }
    }
    return 0;
}

extern "C" void *llz_cuda_mgpu_result_ptr(unsigned long ctx, int local_idx)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c || local_idx < 0 || local_idx >= c->nlocal) return nullptr;
    return c->slot[local_idx].result_ptr;
}

extern "C" int llz_cuda_mgpu_result_free(unsigned long ctx)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c) return -1;
    return mgpu_result_release(c);
}

// ======================================================================================================================
// jobs
// ======================================================================================================================
extern "C" unsigned long llz_cuda_mgpu_fir_init(unsigned long ctx, int kind, int flt_len, double fc1, double fc2,
                                                win_t win_type, int n_channels, int dtype, int shard_mode)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c) return kFail;
    if (shard_mode != LLZ_CUDA_SHARD_CHANNEL && shard_mode != LLZ_CUDA_SHARD_TIME) { llz_set_error("unknown shard mode %d", shard_mode); return kFail; }
    if (shard_mode == LLZ_CUDA_SHARD_CHANNEL && n_channels < c->world) {
        llz_set_error("channel sharding needs at least one channel per rank (%d channels, %d ranks)", n_channels, c->world);
        return kFail;
    }
    MgpuJob *j = new (std::nothrow) MgpuJob();
    if (!j) { llz_set_error("out of memory"); return kFail; }
    j->ctx = c;
    j->is_fir = true;
    j->shard_mode = shard_mode;
    j->n_channels = n_channels;
    j->dtype = dtype;
    for (int i = 0; i < c->nlocal; ++i) {
        int first = 0, cnt = n_channels;
        if (shard_mode == LLZ_CUDA_SHARD_CHANNEL) llz_cuda_shard_channels(n_channels, c->world, c->slot[i].rank, &first, &cnt);
        DeviceGuard g(c->slot[i].device);
        j->bank[i] = llz_cuda_fir_bank_init(kind, flt_len, fc1, fc2, win_type, cnt, dtype);
        if (j->bank[i] == kFail) { job_destroy(j); return kFail; }
    }
    j->flt_len = llz_cuda_fir_bank_flt_len(j->bank[0]);
    return reinterpret_cast<unsigned long>(j);
}

extern "C" unsigned long llz_cuda_mgpu_resample_init(unsigned long ctx, int L, int M, double gain, win_t win_type,
                                                     int k_override, int n_channels, int acc, int shard_mode)
{
    MgpuCtx *c = as_mgpu(ctx);
    if (!c) return kFail;
    if (shard_mode != LLZ_CUDA_SHARD_CHANNEL && shard_mode != LLZ_CUDA_SHARD_TIME) { llz_set_error("unknown shard mode %d", shard_mode); return kFail; }
    if (shard_mode == LLZ_CUDA_SHARD_CHANNEL && n_channels < c->world) {
        llz_set_error("channel sharding needs at least one channel per rank (%d channels, %d ranks)", n_channels, c->world);
        return kFail;
    }
    MgpuJob *j = new (std::nothrow) MgpuJob();
    if (!j) { llz_set_error("out of memory"); return kFail; }
    j->ctx = c;
    j->is_fir = false;
    j->shard_mode = shard_mode;
    j->n_channels = n_channels;
    for (int i = 0; i < c->nlocal; ++i) {
        int first = 0, cnt = n_channels;
        if (shard_mode == LLZ_CUDA_SHARD_CHANNEL) llz_cuda_shard_channels(n_channels, c->world, c->slot[i].rank, &first, &cnt);
        DeviceGuard g(c->slot[i].device);
        j->bank[i] = llz_cuda_resample_bank_init(L, M, gain, win_type, k_override, cnt, acc);
        if (j->bank[i] == kFail) { job_destroy(j); return kFail; }
    }
    llz_cuda_resample_info_t info;
    llz_cuda_resample_bank_info(j->bank[0], &info);
    j->L = info.L; j->M = info.M; j->taps_per_phase = info.taps_per_phase; j->frame_in = info.num_in;
    return reinterpret_cast<unsigned long>(j);
}

extern "C" void llz_cuda_mgpu_job_uninit(unsigned long job)
{
    MgpuJob *j = as_job(job);
    if (j) job_destroy(j);
}

extern "C" unsigned long llz_cuda_mgpu_job_bank(unsigned long job, int local_idx)
{
    MgpuJob *j = as_job(job);
    if (!j || local_idx < 0 || local_idx >= j->ctx->nlocal) return kFail;
    return j->bank[local_idx];
}

extern "C" int llz_cuda_mgpu_job_plan(unsigned long job, long long n_total, int rank, llz_cuda_shard_t *shard)
{
    MgpuJob *j = as_job(job);
    if (!j || !shard) return -1;
    if (rank < 0 || rank >= j->ctx->world || n_total < 0) { llz_set_error("mgpu_job_plan: bad arguments"); return -1; }
    return job_plan_rank(j, n_total, rank, shard);
}

extern "C" long long llz_cuda_mgpu_job_out_len(unsigned long job, long long n_total)
{
    MgpuJob *j = as_job(job);
    if (!j || n_total < 0) return -1;
    return job_total_out(j, n_total);
}

extern "C" int llz_cuda_mgpu_job_run(unsigned long job, long long n_total, const void *const *d_in,
                                     const long long *in_stride, void *const *d_out, const long long *out_stride,
                                     long long result_stride, int gather, int chunks,
                                     const llz_cuda_stream_t *streams)
{
    MgpuJob *j = as_job(job);
    const NcclApi *api = nccl_api();
    if (!j || !api) return -1;
    MgpuCtx *c = j->ctx;
    if (!d_in || !in_stride || n_total < 0) { llz_set_error("mgpu_job_run: bad arguments"); return -1; }
    if (gather != LLZ_CUDA_GATHER_NONE && gather != LLZ_CUDA_GATHER_NCCL && gather != LLZ_CUDA_GATHER_PEER &&
        gather != LLZ_CUDA_GATHER_COPY) {
        llz_set_error("unknown gather mode %d", gather);
        return -1;
    }
    if (gather == LLZ_CUDA_GATHER_NONE && (!d_out || !out_stride)) { llz_set_error("mgpu_job_run: no output buffers"); return -1; }
    const bool staged = gather == LLZ_CUDA_GATHER_NCCL || gather == LLZ_CUDA_GATHER_COPY;   // shards staged in d_out, moved chunk by chunk
    if (staged && (!d_out || !out_stride)) { llz_set_error("mgpu_job_run: the chunked gathers stage each shard in d_out"); return -1; }
    if (chunks < 1) chunks = 4;
    if (chunks > kMaxChunks) chunks = kMaxChunks;
    if (!staged) chunks = 1;
    const size_t es = job_elem_size(j);
    const long long total_out = job_total_out(j, n_total);
    const int root = c->result_root;
    if (gather != LLZ_CUDA_GATHER_NONE) {
        if (root < 0) { llz_set_error("mgpu_job_run: call llz_cuda_mgpu_result_alloc first"); return -1; }
        if (result_stride < total_out) { llz_set_error("mgpu_job_run: result_stride %lld < %lld outputs per channel", result_stride, total_out); return -1; }
        if ((size_t)j->n_channels * (size_t)result_stride * es > c->result_bytes) {
            llz_set_error("mgpu_job_run: the result buffer holds %zu bytes, the job needs %zu", c->result_bytes,
                          (size_t)j->n_channels * (size_t)result_stride * es);
            return -1;
        }
    }
    const long long granule = job_granule(j);
    if (granule < 1) return -1;

    llz_cuda_shard_t sh[kMaxSlots];
    cudaStream_t st[kMaxSlots];
    const unsigned char *x_own[kMaxSlots];      // first owned sample of the shard's first channel
    unsigned char *y_base[kMaxSlots];           // where output 0 of the shard's first channel goes
    long long y_stride[kMaxSlots];
    for (int i = 0; i < c->nlocal; ++i) {
        MgpuSlot &s = c->slot[i];
        if (job_plan_rank(j, n_total, s.rank, &sh[i]) != 0) return -1;
        st[i] = streams ? (cudaStream_t)streams[i] : nullptr;
        if (!d_in[i]) { llz_set_error("mgpu_job_run: null input for local slot %d", i); return -1; }
        x_own[i] = static_cast<const unsigned char *>(d_in[i]) + (size_t)sh[i].seg.halo * es;
        const bool to_result = gather == LLZ_CUDA_GATHER_PEER || (staged && s.rank == root);
        if (to_result) {
            if (!s.result_ptr) {
                llz_set_error("mgpu_job_run: rank %d has no peer mapping of the result buffer (no NVLink / P2P path); use LLZ_CUDA_GATHER_NCCL", s.rank);
                return -1;
            }
            y_base[i] = static_cast<unsigned char *>(s.result_ptr) + ((size_t)sh[i].first_channel * result_stride + sh[i].seg.out_start) * es;
            y_stride[i] = result_stride;
        } else {
            if (!d_out[i]) { llz_set_error("mgpu_job_run: null output for local slot %d", i); return -1; }
            if (gather == LLZ_CUDA_GATHER_COPY && !s.result_ptr) {
                llz_set_error("mgpu_job_run: rank %d has no peer mapping of the result buffer (no NVLink / P2P path); use LLZ_CUDA_GATHER_NCCL", s.rank);
                return -1;
            }
            if (gather == LLZ_CUDA_GATHER_NCCL && out_stride[i] != sh[i].seg.out_count) {
                llz_set_error("mgpu_job_run: the NCCL gather needs dense shard outputs (out_stride %lld, shard has %lld outputs)",
                              out_stride[i], sh[i].seg.out_count);
                return -1;
            }
            y_base[i] = static_cast<unsigned char *>(d_out[i]);
            y_stride[i] = out_stride[i];
        }
    }

    // ---- stream state: a job run is a one-shot whole-signal call (reset), time segments load their halo ----
    for (int i = 0; i < c->nlocal; ++i) {
        MgpuSlot &s = c->slot[i];
        DeviceGuard g(s.device);
        if (j->is_fir) {
            if (llz_cuda_fir_bank_reset(j->bank[i], st[i]) != 0) return -1;
            if (sh[i].seg.halo > 0) {
                if (sh[i].seg.halo != j->flt_len - 1) { llz_set_error("internal: FIR halo %lld", sh[i].seg.halo); return -1; }
                if (llz_cuda_fir_bank_set_history(j->bank[i], d_in[i], in_stride[i], st[i]) != 0) return -1;
            }
        } else {
            if (llz_cuda_resample_bank_reset(j->bank[i], st[i]) != 0) return -1;
            if (sh[i].seg.halo > 0) {
                if (sh[i].seg.halo != j->taps_per_phase - 1) { llz_set_error("internal: resampler halo %lld", sh[i].seg.halo); return -1; }
                if (llz_cuda_resample_bank_set_history(j->bank[i], static_cast<const short *>(d_in[i]), in_stride[i], st[i]) != 0) return -1;
            }
        }
        if (staged) {
            // the gather stream must not start before the caller's stream has reached this call (buffer reuse)
            LLZ_CUDA_TRY(cudaEventRecord(s.ev_enter, st[i]));
            LLZ_CUDA_TRY(cudaStreamWaitEvent(s.s_comm, s.ev_enter, 0));
        }
    }

    // ---- chunks: compute chunk k on every local slot, then move chunk k while chunk k+1 is computed ----
    for (int k = 0; k < chunks; ++k) {
        for (int i = 0; i < c->nlocal; ++i) {
            MgpuSlot &s = c->slot[i];
            DeviceGuard g(s.device);
            const ChunkPlan p = job_chunk(j, sh[i], granule, chunks, k);
            // channel groups: the bank's step completes with the last NON-EMPTY group (the remainder goes to the first ones)
            bool last_group = true;
            if (j->shard_mode == LLZ_CUDA_SHARD_CHANNEL)
                for (int k2 = k + 1; k2 < chunks; ++k2)
                    if (job_chunk(j, sh[i], granule, chunks, k2).cc > 0) last_group = false;
            if (job_run_chunk(j, i, p, x_own[i], in_stride[i], y_base[i], y_stride[i], st[i], last_group) != 0) return -1;
            if (staged) {
                LLZ_CUDA_TRY(cudaEventRecord(s.ev_chunk[k], st[i]));
                LLZ_CUDA_TRY(cudaStreamWaitEvent(s.s_comm, s.ev_chunk[k], 0));
            }
            if (gather == LLZ_CUDA_GATHER_COPY && s.rank != root && p.cc > 0 && p.out_len > 0) {
                // push the chunk into its place in the root's buffer: copy engine, second stream
                const unsigned char *src = y_base[i] + ((size_t)p.c0 * y_stride[i] + p.out0) * es;
                unsigned char *dst = static_cast<unsigned char *>(s.result_ptr) +
                                     ((size_t)(sh[i].first_channel + p.c0) * result_stride + sh[i].seg.out_start + p.out0) * es;
                if (copy_planar(dst, (size_t)result_stride * es, src, (size_t)y_stride[i] * es, (size_t)p.out_len * es, (size_t)p.cc,
                                    cudaMemcpyDeviceToDevice, s.s_comm) != 0)
                    return -1;
            }
        }
        if (gather != LLZ_CUDA_GATHER_NCCL) continue;
        LLZ_NCCL_TRY(api, api->GroupStart());
        for (int i = 0; i < c->nlocal; ++i) {
            MgpuSlot &s = c->slot[i];
            if (s.rank == root) {
                // receive chunk k of every other rank straight into its place in the result
                for (int r = 0; r < c->world; ++r) {
                    if (r == root) continue;
                    llz_cuda_shard_t rs;
                    if (job_plan_rank(j, n_total, r, &rs) != 0) { api->GroupEnd(); return -1; }
                    const ChunkPlan p = job_chunk(j, rs, granule, chunks, k);
                    if (p.cc <= 0 || p.out_len <= 0) continue;
                    unsigned char *dst = static_cast<unsigned char *>(s.result_ptr) +
                                         ((size_t)(rs.first_channel + p.c0) * result_stride + rs.seg.out_start + p.out0) * es;
                    if (result_stride == p.out_len) {
                        LLZ_NCCL_TRY(api, api->Recv(dst, (size_t)p.cc * p.out_len * es, ncclChar, r, s.comm, s.s_comm));
                    } else {
                        for (int ch = 0; ch < p.cc; ++ch)
                            LLZ_NCCL_TRY(api, api->Recv(dst + (size_t)ch * result_stride * es, (size_t)p.out_len * es, ncclChar, r, s.comm, s.s_comm));
                    }
                }
            } else {
                const ChunkPlan p = job_chunk(j, sh[i], granule, chunks, k);
                if (p.cc <= 0 || p.out_len <= 0) continue;
                const unsigned char *src = y_base[i] + ((size_t)p.c0 * y_stride[i] + p.out0) * es;
                // the same message shapes as the root posts: whole chunk when the DESTINATION rows are dense
                if (result_stride == p.out_len) {
                    if (y_stride[i] != p.out_len) { api->GroupEnd(); llz_set_error("internal: dense result needs dense shards"); return -1; }
                    LLZ_NCCL_TRY(api, api->Send(src, (size_t)p.cc * p.out_len * es, ncclChar, root, s.comm, s.s_comm));
                } else {
                    for (int ch = 0; ch < p.cc; ++ch)
                        LLZ_NCCL_TRY(api, api->Send(src + (size_t)ch * y_stride[i] * es, (size_t)p.out_len * es, ncclChar, root, s.comm, s.s_comm));
                }
            }
        }
        LLZ_NCCL_TRY(api, api->GroupEnd());
    }

    // ---- completion ----
    if (gather == LLZ_CUDA_GATHER_NCCL) {
        for (int i = 0; i < c->nlocal; ++i) {
            MgpuSlot &s = c->slot[i];
            DeviceGuard g(s.device);
            LLZ_CUDA_TRY(cudaEventRecord(s.ev_comm, s.s_comm));
            LLZ_CUDA_TRY(cudaStreamWaitEvent(st[i], s.ev_comm, 0));
        }
    } else if ((gather == LLZ_CUDA_GATHER_PEER || gather == LLZ_CUDA_GATHER_COPY) && c->world > 1) {
        if (gather == LLZ_CUDA_GATHER_COPY) {
            for (int i = 0; i < c->nlocal; ++i) {              // the pushes of this rank precede its "done"
                MgpuSlot &s = c->slot[i];
                DeviceGuard g(s.device);
                LLZ_CUDA_TRY(cudaEventRecord(s.ev_comm, s.s_comm));
                LLZ_CUDA_TRY(cudaStreamWaitEvent(st[i], s.ev_comm, 0));
            }
        }
        // every rank's kernels have stored into the root's buffer: a one-int all-reduce on the callers' streams orders
        // "all ranks done" after each rank's kernels (stores to peer memory are performed at kernel completion)
        LLZ_NCCL_TRY(api, api->GroupStart());
        for (int i = 0; i < c->nlocal; ++i) {
            MgpuSlot &s = c->slot[i];
            LLZ_NCCL_TRY(api, api->AllReduce(s.d_flag, s.d_flag + 1, 1, ncclInt, ncclSum, s.comm, st[i]));
        }
        LLZ_NCCL_TRY(api, api->GroupEnd());
    }
    return 0;
}
