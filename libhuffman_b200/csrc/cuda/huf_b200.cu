// huf_b200.cu — the C-ABI shim declared in <huffman/b200.h>: context, device workspace,
// kernel launches.  Host code above this file is plain C; nothing below it runs on the CPU
// except launch orchestration.  Compiled for sm_100a only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <huffman/b200.h>

#include "dec_kernels.cuh"
#include "dec_fast.cuh"
#include "enc_kernels.cuh"
#include "enc_build.cuh"
#include "enc_pack.cuh"
#include "host_pipe.cuh"

using namespace hufb200;

#ifndef HUF_EMU
#define HUF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// Launch through the context: counts the launch and, when kernel timing is on, brackets it
// with events on the launching stream.
#define CTX_LAUNCH(c, kernel, grid, block, smem, stream, ...)                      \
    do {                                                                           \
        huf_b200_ctx::Timed *t__ = nullptr;                                        \
        if ((c)->timing && (c)->ntimed < 64) {                                     \
            t__ = &(c)->timed[(c)->ntimed++];                                      \
            t__->name = #kernel;                                                   \
            cudaEventCreate(&t__->t0);                                             \
            cudaEventCreate(&t__->t1);                                             \
            cudaEventRecord(t__->t0, (stream));                                    \
        }                                                                          \
        HUF_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__);                \
        if (t__) cudaEventRecord(t__->t1, (stream));                               \
        if ((c)->debug) {                                                          \
            cudaError_t le__ = cudaPeekAtLastError();                              \
            if (le__ != cudaSuccess)                                               \
                fprintf(stderr, "huf_b200: launch of %s failed: %s\n", #kernel,    \
                        cudaGetErrorString(le__));                                 \
        }                                                                          \
        (c)->launches++;                                                           \
    } while (0)

#ifndef HUF_EMU
// (load time: ask for enough hardware work queues for the encoder's pass pipeline, unless the
// application has chosen a value itself)
__attribute__((constructor)) static void huf_b200_on_load() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
#endif

namespace {

constexpr uint32_t kSegMax = 16384;            // bytes per segment (u16 counters suffice)
constexpr uint64_t kMaxPassBlocks = 1u << 18;  // blocks per encode pass (bounds the workspace)
constexpr uint64_t kW32MaxBlock = 4u << 20;    // blocks up to 4 MiB use 32-bit merge keys
constexpr int kEncSlots = 8;                   // passes of a large encode call in flight (a workspace slot each)
constexpr int kEncHistStreams = 2, kEncBuildStreams = 8, kEncPackStreams = 4, kEncWideStreams = 2;
constexpr uint64_t kEncPipeMinBytes = 48ull << 20;  // smaller calls run as one pass on the caller's stream
constexpr uint64_t kEncPipeMinPass = 4ull << 20;    // a pass of the pipeline covers at least this many input bytes

struct Arena {
    uint8_t *base = nullptr;
    size_t cap = 0;
    size_t used = 0;

    bool reserve(size_t bytes)
    {
        used = 0;
        if (bytes <= cap) return true;
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
        // grow with headroom so that repeated calls of similar size do not reallocate
        size_t want = bytes + bytes / 8 + (1u << 20);
        if (cudaMalloc(&base, want) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cap = want;
        return true;
    }
    template <typename T>
    T *take(size_t count)
    {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

}  // namespace

struct huf_b200_ctx {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t cur = nullptr;
    Arena enc_ws, dec_ws;
    uint32_t *d_status = nullptr;   // [4]
    uint64_t *d_result = nullptr;   // [16]
    uint64_t *h_result = nullptr;   // pinned mirror, [32]
    uint64_t launches = 0;
    int accept_1025 = 0;
    bool debug = false;             // HUF_B200_DEBUG: report failing launches on stderr

    // optional per-kernel timing (HUF_B200_OPT_KERNEL_TIMING): events around every launch
    bool timing = false;
    bool timeline = false;          // timing value 2: keep the pass pipeline on, report start offsets too
    struct Timed {
        const char *name;
        cudaEvent_t t0, t1;
    };
    Timed timed[64];
    int ntimed = 0;

    // encode call in flight
    // streams of the pipelined encode: histograms (low priority), code builds (high: short
    // grids of latency-bound warps that must not queue behind thousands of CTAs), packing
    cudaStream_t enc_hist_st[kEncHistStreams] = {}, enc_build_st[kEncBuildStreams] = {}, enc_pack_st[kEncPackStreams] = {},
                 enc_wide_st[kEncWideStreams] = {};
    cudaEvent_t enc_ev_hist[kEncSlots], enc_ev_build[kEncSlots], enc_ev_pack[kEncSlots], enc_ev_wide[kEncSlots], enc_ev_fork;
    bool enc_side_ready = false;
    bool no_overlap = false;        // HUF_B200_OPT_NO_OVERLAP / HUF_B200_NO_OVERLAP=1: passes one after the other on one stream
    uint64_t enc_pipe_min = kEncPipeMinBytes;       // (HUF_B200_ENC_PIPE_MIN, HUF_B200_ENC_PIPE_PASS, HUF_B200_ENC_SLOTS:
    uint64_t enc_pipe_min_pass = kEncPipeMinPass;   //  the pipeline's thresholds in bytes and its slot count, for tests)
    int enc_slots = kEncSlots;
    bool enc_pending = false;
    EncArgs enc{};
    uint64_t *d_blk_off = nullptr;  // [nblocks + 1], lives in enc_ws
    uint64_t enc_nblocks = 0;

    // decode call in flight
    bool dec_pending = false;
    uint64_t dec_first = 0;         // offset the pending decode call started at
    uint64_t dec_first_cand = ~0ull;  // offset of the first block the call found (range mode)
    uint32_t *dec_terms = nullptr;    // terminal slots of very large decode calls (sized after the scan)
    uint64_t dec_terms_cap = 0;
    DecArgs dec{};
    bool dec_dense = false;         // stream has many tiny blocks: use the exact two-pass header scan
    const uint64_t *hint_off = nullptr;  // optional block index for the next decode (device pointer)
    uint64_t hint_n = 0;
    uint32_t dec_stage = 0;         // dynamic smem bytes for k_decode_slow
    bool slow_ready = false;        // k_decode_slow attribute set
    bool fast_ready = false;        // k_decode attributes set
    bool lut_or = true;             // k_decode's table is 8 KB aligned in the shared window (probed)
    bool force_lut_add = false;     // HUF_B200_OPT_FORCE_LUT_ADD: always launch k_decode_unaligned
    int fast_per_sm = 1;
    int fast_pad = 0;
    uint64_t dec_stage_want = 80 * 1024;  // payload bytes of one block staged in shared memory

    pipe::PipeState pipe;           // streams, events and cached buffers of the host-buffer lanes
    uint64_t dec_margin = 1u << 20; // host lane: bytes kept back behind a pass so its last block is whole
};

namespace {

huf_error_t cuda_fail(cudaError_t e, int line = 0)
{
    if (getenv("HUF_B200_DEBUG"))
        fprintf(stderr, "huf_b200: CUDA error %d (%s) at huf_b200.cu:%d\n", (int)e, cudaGetErrorString(e), line);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? HUF_ERROR_MEMORY_ALLOCATION : HUF_ERROR_FATAL;
}

#define CU_TRY(expr)                                   \
    do {                                               \
        cudaError_t e__ = (expr);                      \
        if (e__ != cudaSuccess) return cuda_fail(e__, __LINE__); \
    } while (0)

#define HUF_TRY_CXX(expr)                              \
    do {                                               \
        huf_error_t e__ = (expr);                      \
        if (e__ != HUF_ERROR_SUCCESS) return e__;      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Timed launches of a call whose times were never read (huf_b200_kernel_times): release their
// events before the slots are reused.
void drop_timed(huf_b200_ctx *c)
{
    for (int i = 0; i < c->ntimed; i++) {
        cudaEventDestroy(c->timed[i].t0);
        cudaEventDestroy(c->timed[i].t1);
    }
    c->ntimed = 0;
}

cudaStream_t pick_stream(huf_b200_ctx *c, void *stream)
{
    return stream == HUF_B200_STREAM_PRIVATE ? c->own_stream : (cudaStream_t)stream;
}

uint32_t pick_seg(uint64_t blocksize)
{
    uint64_t s = blocksize < kSegMax ? blocksize : kSegMax;
    s = (s + 15) & ~uint64_t(15);
    return (uint32_t)(s ? s : 16);
}

// ---- staged host <-> device copies ---------------------------------------------------------

constexpr uint64_t kStageChunk = 32ull << 20;  // bytes per pinned bounce buffer
constexpr uint64_t kStageMin = 8ull << 20;     // smaller copies go straight through cudaMemcpy

struct StagePool {
    std::mutex mu;
    bool tried = false, ok = false;
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    cudaStream_t stream = nullptr;
};
StagePool g_stage;

bool stage_ready()
{
#ifdef HUF_EMU
    return false;  // kernel-logic emulation: plain copies
#else
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (g_stage.tried) return g_stage.ok;
    g_stage.tried = true;
    cudaError_t e = cudaStreamCreateWithFlags(&g_stage.stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaHostAlloc(reinterpret_cast<void **>(&g_stage.pin[i]), kStageChunk, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_stage.done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(g_stage.done[i], g_stage.stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    g_stage.ok = true;
    return true;
#endif
}

void parallel_memcpy(void *dst, const void *src, uint64_t bytes)
{
    pipe::CopyPool::get().copy(dst, src, bytes);
}

}  // namespace

extern "C" {

int huf_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

huf_error_t huf_b200_ctx_create(huf_b200_ctx_t **out, int device)
{
    if (!out) return HUF_ERROR_INVALID_ARGUMENT;
    *out = nullptr;
    if (huf_b200_device_count() <= 0) return HUF_ERROR_FATAL;  // no GPU: no CPU path either
    if (device < 0) CU_TRY(cudaGetDevice(&device));

    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return HUF_ERROR_FATAL;  // kernels are built for sm_100a only

    DeviceGuard g(device);
    if (!g.ok) return HUF_ERROR_FATAL;
    huf_b200_ctx *c = new (std::nothrow) huf_b200_ctx();
    if (!c) return HUF_ERROR_MEMORY_ALLOCATION;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    const char *env = getenv("HUF_B200_ACCEPT_1025");
    c->accept_1025 = env && env[0] == '1';
    c->debug = getenv("HUF_B200_DEBUG") != nullptr;
    env = getenv("HUF_B200_NO_OVERLAP");
    c->no_overlap = env && env[0] == '1';
    // The pass pipeline of the encoder uses 16 streams.  With the driver's default of 8 hardware
    // work queues they alias and serialise falsely (measured: 1.53 ms against 1.41 ms on one
    // stream and 1.33 ms with 32 queues), so the pipeline is only used when the process runs
    // with CUDA_DEVICE_MAX_CONNECTIONS >= 16 (this library sets 32 when it is loaded, which
    // takes effect if that happens before CUDA is initialised).
    env = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
    if (!env || atoi(env) < 16) c->no_overlap = true;
    if ((env = getenv("HUF_B200_ENC_PIPE_MIN")) && atoll(env) > 0) c->enc_pipe_min = (uint64_t)atoll(env);
    if ((env = getenv("HUF_B200_ENC_PIPE_PASS")) && atoll(env) > 0) c->enc_pipe_min_pass = (uint64_t)atoll(env);
    if ((env = getenv("HUF_B200_ENC_SLOTS")) && atoi(env) >= 1 && atoi(env) <= kEncSlots) c->enc_slots = atoi(env);
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_status, 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_result, 16 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_result, 32 * sizeof(uint64_t));
    if (e != cudaSuccess) {
        huf_b200_ctx_destroy(&c);
        return cuda_fail(e);
    }
    *out = c;
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_ctx_destroy(huf_b200_ctx_t **ctx)
{
    if (!ctx) return HUF_ERROR_INVALID_ARGUMENT;
    huf_b200_ctx *c = *ctx;
    if (c) {
        DeviceGuard g(c->device);
        cudaDeviceSynchronize();
        drop_timed(c);
        c->pipe.release();
        if (c->enc_ws.base) cudaFree(c->enc_ws.base);
        if (c->dec_ws.base) cudaFree(c->dec_ws.base);
        if (c->dec_terms) cudaFree(c->dec_terms);
        if (c->d_status) cudaFree(c->d_status);
        if (c->d_result) cudaFree(c->d_result);
        if (c->h_result) cudaFreeHost(c->h_result);
        if (c->enc_side_ready) {
            // (the streams belong to the device's pool, see enc_stream_pool)
            for (int i = 0; i < kEncSlots; i++) {
                cudaEventDestroy(c->enc_ev_hist[i]);
                cudaEventDestroy(c->enc_ev_build[i]);
                cudaEventDestroy(c->enc_ev_pack[i]);
                cudaEventDestroy(c->enc_ev_wide[i]);
            }
            cudaEventDestroy(c->enc_ev_fork);
        }
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        delete c;
    }
    *ctx = nullptr;
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_ctx_set_option(huf_b200_ctx_t *ctx, int option, int64_t value)
{
    if (!ctx) return HUF_ERROR_INVALID_ARGUMENT;
    switch (option) {
    case HUF_B200_OPT_ACCEPT_1025:
        ctx->accept_1025 = value != 0;
        return HUF_ERROR_SUCCESS;
    case HUF_B200_OPT_KERNEL_TIMING:
        ctx->timing = value != 0;
        ctx->timeline = value == 2;
        return HUF_ERROR_SUCCESS;
    case HUF_B200_OPT_FORCE_LUT_ADD:
        ctx->force_lut_add = value != 0;
        ctx->fast_ready = false;  // choose the kernel instance again
        return HUF_ERROR_SUCCESS;
    case HUF_B200_OPT_NO_OVERLAP:
        ctx->no_overlap = value != 0;
        return HUF_ERROR_SUCCESS;
    default:
        return HUF_ERROR_INVALID_ARGUMENT;
    }
}

uint64_t huf_b200_block_count(uint64_t length, uint64_t blocksize)
{
    if (!length) return 0;
    if (!blocksize) blocksize = length;
    return (length + blocksize - 1) / blocksize;
}

uint64_t huf_b200_encode_bound(uint64_t length, uint64_t blocksize)
{
    // per block: 10 + 2*1025 header bytes; payload < 10 bits per symbol is the true bound for a
    // 256-symbol alphabet with the extra root bit, rounded up per block
    const uint64_t nb = huf_b200_block_count(length, blocksize);
    return nb * (kHdrFixed + 2 * kMaxTreeElems + 8) + length + length / 4 + 16;
}

uint64_t huf_b200_last_launch_count(const huf_b200_ctx_t *ctx) { return ctx ? ctx->launches : 0; }

uint64_t huf_b200_last_slow_blocks(const huf_b200_ctx_t *ctx) { return ctx ? ctx->h_result[10] : 0; }

huf_error_t huf_b200_kernel_times(huf_b200_ctx_t *c, char *buf, uint64_t buflen)
{
    if (!c || !buf || !buflen) return HUF_ERROR_INVALID_ARGUMENT;
    DeviceGuard g(c->device);
    size_t at = 0;
    buf[0] = 0;
    for (int i = 0; i < c->ntimed; i++) {
        float ms = 0.f, from0 = 0.f;
        cudaEventSynchronize(c->timed[i].t1);
        cudaEventElapsedTime(&ms, c->timed[i].t0, c->timed[i].t1);
        int n;
        if (c->timeline) {
            // (timeline mode: "name@start duration", the start relative to the first launch)
            cudaEventElapsedTime(&from0, c->timed[0].t0, c->timed[i].t0);
            n = snprintf(buf + at, buflen - at, "%s@%.4f %.6f\n", c->timed[i].name, (double)from0, (double)ms);
        } else {
            n = snprintf(buf + at, buflen - at, "%s %.6f\n", c->timed[i].name, (double)ms);
        }
        if (n < 0 || (size_t)n >= buflen - at) break;
        at += (size_t)n;
    }
    for (int i = 0; i < c->ntimed; i++) {
        cudaEventDestroy(c->timed[i].t0);
        cudaEventDestroy(c->timed[i].t1);
    }
    c->ntimed = 0;
    return HUF_ERROR_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------

namespace {
huf_error_t encode_enqueue(huf_b200_ctx_t *c, const void *d_in, uint64_t length, uint64_t blocksize,
                           void *d_out, uint64_t out_capacity, void *stream);
}

huf_error_t huf_b200_encode_async(huf_b200_ctx_t *c, const void *d_in, uint64_t length,
                                  uint64_t blocksize, void *d_out, uint64_t out_capacity,
                                  void *stream)
{
    if (!c || (!d_in && length) || (!d_out && length)) return HUF_ERROR_INVALID_ARGUMENT;
    // (an encode may be enqueued again while one is pending -- several on one stream, one finish
    // for the last -- but not while a decode of this context is in flight)
    if (c->dec_pending) return HUF_ERROR_INVALID_ARGUMENT;
    // the call is pending (encode_finish owed) only once everything was enqueued: a failure on
    // the way (workspace allocation, launch error) leaves the context free for the next call
    const huf_error_t e = encode_enqueue(c, d_in, length, blocksize, d_out, out_capacity, stream);
    c->enc_pending = e == HUF_ERROR_SUCCESS;
    return e;
}

namespace {
// The streams of the encoder's pass pipeline: one set per device for the whole process, created
// once and never destroyed.  (Per context they would be created and destroyed with every codec
// object; the driver hands its hardware work queues to streams as they are created, and after
// a few generations the sixteen streams of a context shared queues with each other: the same
// encode measured 936 and 700 GB/s in two contexts of one process.)
struct EncStreamPool {
    bool ready = false;
    cudaStream_t hist[kEncHistStreams], build[kEncBuildStreams], pack[kEncPackStreams], wide[kEncWideStreams];
};
constexpr int kMaxPoolDevices = 64;
EncStreamPool g_enc_pool[kMaxPoolDevices];
std::mutex g_enc_pool_mu;

huf_error_t enc_stream_pool(huf_b200_ctx *c)
{
    if (c->device < 0 || c->device >= kMaxPoolDevices) return HUF_ERROR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(g_enc_pool_mu);
    EncStreamPool &p = g_enc_pool[c->device];
    if (!p.ready) {
        int prio_lo = 0, prio_hi = 0;  // (numerically: highest priority = smallest value)
        CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const int prio_mid = prio_hi < prio_lo ? prio_hi + 1 : prio_lo;
        for (int i = 0; i < kEncHistStreams; i++)
            CU_TRY(cudaStreamCreateWithPriority(&p.hist[i], cudaStreamNonBlocking, prio_lo));
        for (int i = 0; i < kEncBuildStreams; i++)
            CU_TRY(cudaStreamCreateWithPriority(&p.build[i], cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < kEncPackStreams; i++)
            CU_TRY(cudaStreamCreateWithPriority(&p.pack[i], cudaStreamNonBlocking, prio_mid));
        for (int i = 0; i < kEncWideStreams; i++)
            CU_TRY(cudaStreamCreateWithPriority(&p.wide[i], cudaStreamNonBlocking, prio_mid));
        p.ready = true;
    }
    for (int i = 0; i < kEncHistStreams; i++) c->enc_hist_st[i] = p.hist[i];
    for (int i = 0; i < kEncBuildStreams; i++) c->enc_build_st[i] = p.build[i];
    for (int i = 0; i < kEncPackStreams; i++) c->enc_pack_st[i] = p.pack[i];
    for (int i = 0; i < kEncWideStreams; i++) c->enc_wide_st[i] = p.wide[i];
    return HUF_ERROR_SUCCESS;
}

huf_error_t encode_enqueue(huf_b200_ctx_t *c, const void *d_in, uint64_t length, uint64_t blocksize,
                           void *d_out, uint64_t out_capacity, void *stream)
{
    DeviceGuard g(c->device);
    if (!g.ok) return HUF_ERROR_FATAL;
    cudaStream_t st = pick_stream(c, stream);
    c->cur = st;
    c->launches = 0;
    drop_timed(c);
    if (!blocksize) blocksize = length;

    const uint64_t nblocks = huf_b200_block_count(length, blocksize);
    c->enc_nblocks = nblocks;
    EncArgs &a = c->enc;
    memset(&a, 0, sizeof(a));
    if (!nblocks) return HUF_ERROR_SUCCESS;

    a.in = static_cast<const uint8_t *>(d_in);
    a.length = length;
    a.blocksize = blocksize;
    a.nblocks = nblocks;
    a.seg = pick_seg(blocksize);
    const uint64_t nspb64 = (blocksize + a.seg - 1) / a.seg;
    if (nspb64 > 0xffffffffull) return HUF_ERROR_INVALID_ARGUMENT;
    a.nspb = (uint32_t)nspb64;
    a.out = static_cast<uint8_t *>(d_out);
    a.out_cap = out_capacity;

    // pass size: bounded by block count and by ~1 GiB of segment histograms
    uint64_t per_pass = nblocks < kMaxPassBlocks ? nblocks : kMaxPassBlocks;
    const uint64_t seg_cap = (1ull << 21);  // segments per pass
    if (per_pass * a.nspb > seg_cap) per_pass = seg_cap / a.nspb ? seg_cap / a.nspb : 1;
    // Large calls run as a pipeline of passes (block ranges) over workspace slots.  The code
    // build of a pass (K2: per block a chain of up to 255 dependent merge steps, latency bound
    // whatever the grid: 0.3 ms of a 1.4 ms encode at 5-65 % occupancy) then runs under the
    // histograms and the packing of its neighbours.  Three kinds of streams: the histograms of
    // all passes go first (low priority, in pass order), the builds follow them on high-priority
    // streams (a build is a few hundred latency-bound warps and the one-CTA offset scan: queued
    // at equal priority behind the thousands of CTAs of K1/K3 every one of its launches waited
    // ~0.1 ms for a free slot, measured with the timeline mode of HUF_B200_OPT_KERNEL_TIMING),
    // and the packing of a pass starts when its build is done.  Across passes only the
    // block-offset scan chains (the offsets of a pass start at the total of the pass before).
    // With per-kernel timing on everything stays on the caller's stream so that event times add
    // up (value 2 keeps the pipeline and reports start offsets instead).
    bool overlap = (!c->timing || c->timeline) && !c->no_overlap && length >= c->enc_pipe_min && nblocks >= 2 * (uint64_t)kEncSlots;
    if (overlap) {
        uint64_t want = (nblocks + kEncSlots - 1) / kEncSlots;
        const uint64_t min_blocks = (c->enc_pipe_min_pass + blocksize - 1) / blocksize;
        if (want < min_blocks) want = min_blocks;
        // (very large calls: passes of at most 256 MiB over the eight slots in turn, so that the
        // workspace stays small whatever the call size)
        const uint64_t max_blocks = ((256ull << 20) + blocksize - 1) / blocksize;
        if (want > max_blocks) want = max_blocks;
        if (want < per_pass) per_pass = want;
        if (per_pass >= nblocks) overlap = false;
    }
    const uint64_t nseg_pass = per_pass * a.nspb;
    const uint64_t npasses = (nblocks + per_pass - 1) / per_pass;
    const int nslot = overlap ? (int)(npasses < (uint64_t)c->enc_slots ? npasses : (uint64_t)c->enc_slots) : 1;
    if (overlap && !c->enc_side_ready) {
        HUF_TRY_CXX(enc_stream_pool(c));
        for (int i = 0; i < kEncSlots; i++) {
            CU_TRY(cudaEventCreateWithFlags(&c->enc_ev_hist[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&c->enc_ev_build[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&c->enc_ev_pack[i], cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&c->enc_ev_wide[i], cudaEventDisableTiming));
        }
        CU_TRY(cudaEventCreateWithFlags(&c->enc_ev_fork, cudaEventDisableTiming));
        c->enc_side_ready = true;
    }

    size_t need = 0;
    need += Arena::padded(nblocks * sizeof(uint64_t));           // blk_size
    need += Arena::padded((nblocks + 1) * sizeof(uint64_t));     // blk_off
    size_t per_slot = 0;
    per_slot += Arena::padded(nseg_pass * 256 * sizeof(uint16_t));
    per_slot += Arena::padded(nseg_pass * sizeof(uint64_t));
    per_slot += Arena::padded(per_pass * sizeof(uint64_t));          // blk_bits
    per_slot += Arena::padded(per_pass * 512 * sizeof(uint32_t));    // blk_table
    per_slot += Arena::padded(per_pass * kTreeStride * sizeof(int16_t));
    per_slot += Arena::padded(per_pass * 4 * sizeof(uint32_t));
    per_slot += Arena::padded(per_pass * 256 * sizeof(uint32_t));
    per_slot += Arena::padded(per_pass * 512 * sizeof(uint32_t));
    need += (size_t)nslot * per_slot;
    if (!c->enc_ws.reserve(need)) return HUF_ERROR_MEMORY_ALLOCATION;
    a.blk_size = c->enc_ws.take<uint64_t>(nblocks);
    a.blk_off = c->enc_ws.take<uint64_t>(nblocks + 1);
    a.status = c->d_status;
    c->d_blk_off = a.blk_off;
    EncArgs slot[kEncSlots];
    for (int i = 0; i < nslot; i++) {
        slot[i] = a;
        slot[i].seg_hist = c->enc_ws.take<uint16_t>(nseg_pass * 256);
        slot[i].seg_bitoff = c->enc_ws.take<uint64_t>(nseg_pass);
        slot[i].blk_bits = c->enc_ws.take<uint64_t>(per_pass);
        slot[i].blk_table = c->enc_ws.take<uint32_t>(per_pass * 512);
        slot[i].blk_tree = c->enc_ws.take<int16_t>(per_pass * kTreeStride);
        slot[i].blk_meta = c->enc_ws.take<uint32_t>(per_pass * 4);
        slot[i].blk_keys = c->enc_ws.take<uint32_t>(per_pass * 256);
        slot[i].blk_nodes = c->enc_ws.take<uint32_t>(per_pass * 512);
    }

    CU_TRY(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(uint32_t), st));
    CU_TRY(cudaMemsetAsync(a.blk_off, 0, sizeof(uint64_t), st));
    if (overlap) CU_TRY(cudaEventRecord(c->enc_ev_fork, st));  // the side streams start behind the caller's work

    uint64_t pass = 0;
    for (uint64_t blk0 = 0; blk0 < nblocks; blk0 += per_pass, pass++) {
        const int si = overlap ? (int)(pass % (uint64_t)nslot) : 0;
        const int sprev = overlap ? (int)((pass + (uint64_t)nslot - 1) % (uint64_t)nslot) : 0;
        EncArgs &pa = slot[si];
        cudaStream_t hs = overlap ? c->enc_hist_st[pass % kEncHistStreams] : st;
        cudaStream_t bs = overlap ? c->enc_build_st[pass % kEncBuildStreams] : st;
        cudaStream_t ps = overlap ? c->enc_pack_st[pass % kEncPackStreams] : st;
        cudaStream_t ws = overlap ? c->enc_wide_st[pass % kEncWideStreams] : st;
        pa.blk0 = blk0;
        pa.npass = nblocks - blk0 < per_pass ? nblocks - blk0 : per_pass;
        const uint64_t nseg = pa.npass * pa.nspb;
        const unsigned seg_grid = (unsigned)((nseg + kEncWarps - 1) / kEncWarps);
        const unsigned bld_grid = (unsigned)((pa.npass + kBuildWarps - 1) / kBuildWarps);
        const bool reuse = overlap && pass >= (uint64_t)nslot;  // the slot has served an earlier pass

        // K1.  (A slot's histograms are free once the build of its previous pass has read them.)
        if (overlap && pass < (uint64_t)kEncHistStreams) CU_TRY(cudaStreamWaitEvent(hs, c->enc_ev_fork, 0));
        if (reuse) CU_TRY(cudaStreamWaitEvent(hs, c->enc_ev_build[si], 0));
        CTX_LAUNCH(c, k_seg_hist, seg_grid, kEncWarps * 32, 0, hs, pa);
        if (overlap) {
            CU_TRY(cudaEventRecord(c->enc_ev_hist[si], hs));
            // K2 behind K1; a slot's tables are free once its previous pass has been packed
            CU_TRY(cudaStreamWaitEvent(bs, c->enc_ev_hist[si], 0));
            if (reuse) {
                CU_TRY(cudaStreamWaitEvent(bs, c->enc_ev_pack[si], 0));
                CU_TRY(cudaStreamWaitEvent(bs, c->enc_ev_wide[si], 0));
            }
        }
        if (blocksize <= kW32MaxBlock) {
            // sort (warp per block), exact merge (lane per block), codes + tree (warp per block)
            CTX_LAUNCH(c, k_build_sort, bld_grid, kBuildWarps * 32, 0, bs, pa);
            CTX_LAUNCH(c, k_build_merge, (unsigned)((pa.npass + 31) / 32), 32, kMergeDyn, bs, pa);
            CTX_LAUNCH(c, k_build_codes, bld_grid, kBuildWarps * 32, 0, bs, pa);
        } else
            CTX_LAUNCH(c, k_build<uint64_t>, bld_grid, kBuildWarps * 32, 0, bs, pa);
        // (the offsets of this pass continue where the pass before, on another stream, ended)
        if (overlap && pass > 0) CU_TRY(cudaStreamWaitEvent(bs, c->enc_ev_build[sprev], 0));
        CTX_LAUNCH(c, k_scan_sizes, 1, kScanThreads, 0, bs, pa.blk_size + blk0, pa.blk_off + blk0,
                   pa.npass, pa.out_cap, pa.status);
        if (overlap) {
            CU_TRY(cudaEventRecord(c->enc_ev_build[si], bs));
            CU_TRY(cudaStreamWaitEvent(ps, c->enc_ev_build[si], 0));
            CU_TRY(cudaStreamWaitEvent(ws, c->enc_ev_build[si], 0));
        }
        // (the two packing kernels take disjoint blocks -- code words of up to 16 bits / longer --
        // and nearly always one of them finds nothing to do: side by side, so that an empty grid
        // queueing for SM slots does not hold up the packing stream)
        CTX_LAUNCH(c, k_pack, seg_grid, kEncWarps * 32, 0, ps, pa);
        if (overlap) CU_TRY(cudaEventRecord(c->enc_ev_pack[si], ps));
        CTX_LAUNCH(c, k_pack_wide, seg_grid, kEncWarps * 32, 0, ws, pa);
        if (overlap) CU_TRY(cudaEventRecord(c->enc_ev_wide[si], ws));
    }
    if (overlap) {
        // join: the caller's stream continues behind the packing of every slot's last pass
        // (everything else of the call lies in front of those by the event chain)
        for (int i = 0; i < nslot; i++) {
            CU_TRY(cudaStreamWaitEvent(st, c->enc_ev_pack[i], 0));
            CU_TRY(cudaStreamWaitEvent(st, c->enc_ev_wide[i], 0));
        }
    }
    a = slot[overlap ? (int)((pass - 1) % (uint64_t)nslot) : 0];  // (what block_offsets and the debug helpers look at)
    CU_TRY(cudaGetLastError());
    // result: total size + status, copied to the pinned mirror on the same stream
    CU_TRY(cudaMemcpyAsync(&c->h_result[0], a.blk_off + nblocks, sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&c->h_result[1], c->d_status, 2 * sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, st));
    return HUF_ERROR_SUCCESS;
}
}  // namespace

huf_error_t huf_b200_encode_finish(huf_b200_ctx_t *c, uint64_t *out_len)
{
    if (!c || !out_len) return HUF_ERROR_INVALID_ARGUMENT;
    if (!c->enc_pending) return HUF_ERROR_INVALID_ARGUMENT;
    c->enc_pending = false;
    *out_len = 0;
    if (!c->enc_nblocks) return HUF_ERROR_SUCCESS;
    DeviceGuard g(c->device);
    CU_TRY(cudaStreamSynchronize(c->cur));
    const uint32_t status = (uint32_t)(c->h_result[1] & 0xffffffffu);
    if (status != kOk) return (huf_error_t)status;
    *out_len = c->h_result[0];
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_encode_block_offsets(huf_b200_ctx_t *c, const uint64_t **d_offsets,
                                          uint64_t *nblocks)
{
    if (!c || !d_offsets || !nblocks) return HUF_ERROR_INVALID_ARGUMENT;
    *d_offsets = c->d_blk_off;
    *nblocks = c->enc_nblocks;
    return HUF_ERROR_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------

namespace {

// Enqueue one speculative pass starting at the proven block start `first`.
huf_error_t dec_enqueue(huf_b200_ctx *c, uint64_t first, uint64_t out_base, bool plan_only,
                        uint64_t max_cand_hint, bool first_proven = true)
{
    DecArgs &a = c->dec;
    cudaStream_t st = c->cur;
    a.first = first;
    a.first_proven = first_proven ? 1u : 0u;
    a.out_base = out_base;
    a.accept_1025 = (uint32_t)c->accept_1025;
    a.count_only = plan_only ? 1u : 0u;
    a.mul14 = 1u << 14;

    const uint64_t lim = a.length < a.avail ? a.length : a.avail;
    const uint64_t span = lim > first ? lim - first : 0;
    const uint64_t scan = lim > find_base(first) ? lim - find_base(first) : 0;  // the scan starts aligned
    a.nchunks = (scan + kFindChunk - 1) / kFindChunk;
    if (!a.nchunks) a.nchunks = 1;
    uint64_t max_cand = max_cand_hint ? max_cand_hint : span / 256 + 1024;
    const bool hinted = c->hint_off && c->hint_n && first == 0 && !plan_only;
    if (hinted && max_cand < c->hint_n + 16) max_cand = c->hint_n + 16;
    a.max_cand = max_cand;

    size_t need = 0;
    need += Arena::padded(a.nchunks * sizeof(uint32_t));
    need += Arena::padded(a.nchunks * kFindSlots * sizeof(uint32_t));
    need += Arena::padded((a.nchunks + 1) * sizeof(uint64_t));
    need += 4 * Arena::padded((max_cand + 1) * sizeof(uint64_t));
    need += 3 * Arena::padded(max_cand * sizeof(uint32_t));
    // Terminal slots of the fast lane (3.2 KB per candidate).  Ordinary calls get enough for every
    // block of ~1 KiB and larger up front, so the whole pass is enqueued without a host round
    // trip.  Above 2 GiB of stream that guess would reserve tens of GB: such calls learn the
    // candidate count first (one synchronisation, negligible at that size) and reserve exactly.
    const bool exact_terms = !plan_only && span > (2ull << 30);
    uint64_t term_slots = plan_only || exact_terms ? 0 : span / 1024 + 4096;
    if (term_slots > max_cand) term_slots = max_cand;
    need += Arena::padded(term_slots * kTermStride * sizeof(uint32_t));
    if (!c->dec_ws.reserve(need)) return HUF_ERROR_MEMORY_ALLOCATION;
    a.chunk_cnt = c->dec_ws.take<uint32_t>(a.nchunks);
    a.slots = c->dec_ws.take<uint32_t>(a.nchunks * kFindSlots);
    a.chunk_off = c->dec_ws.take<uint64_t>(a.nchunks + 1);
    a.cand = c->dec_ws.take<uint64_t>(max_cand + 1);
    a.olen = c->dec_ws.take<uint64_t>(max_cand + 1);
    a.out_off = c->dec_ws.take<uint64_t>(max_cand + 1);
    a.end_off = c->dec_ws.take<uint64_t>(max_cand + 1);
    a.blk_status = c->dec_ws.take<uint32_t>(max_cand);
    a.meta = c->dec_ws.take<uint32_t>(2 * max_cand);
    a.terms = c->dec_ws.take<uint32_t>(term_slots * kTermStride);
    a.term_slots = term_slots;
    a.result = c->d_result;

    CU_TRY(cudaMemsetAsync(c->d_result, 0, 16 * sizeof(uint64_t), st));
    const unsigned find_grid = (unsigned)((a.nchunks + kFindWarps - 1) / kFindWarps);
    if (hinted) {
        // block index supplied by the caller: no header scan in this pass
        CTX_LAUNCH(c, k_hint, (unsigned)((c->hint_n + 255) / 256), 256, 0, st, a, c->hint_off, c->hint_n);
        c->hint_off = nullptr;  // a restart after a broken chain scans
        c->hint_n = 0;
    } else if (c->dec_dense) {
        CTX_LAUNCH(c, k_find<0>, find_grid, kFindWarps * 32, 0, st, a);
        CTX_LAUNCH(c, k_scan_chunks, 1, kScanThreads, 0, st, a);
        CTX_LAUNCH(c, k_find<1>, find_grid, kFindWarps * 32, 0, st, a);
    } else {
        CTX_LAUNCH(c, k_find<2>, find_grid, kFindWarps * 32, 0, st, a);
        CTX_LAUNCH(c, k_scan_chunks, 1, kScanThreads, 0, st, a);
        CTX_LAUNCH(c, k_compact, (unsigned)((a.nchunks + 255) / 256), 256, 0, st, a);
    }
    CTX_LAUNCH(c, k_gather, c->sm_count * 4, 256, 0, st, a);
    CTX_LAUNCH(c, k_scan_olen, 1, kScanThreads, 0, st, a);
    if (exact_terms) {
        uint64_t found = 0;
        CU_TRY(cudaMemcpyAsync(&found, c->d_result, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const uint64_t want = (found + 16) * kTermStride * sizeof(uint32_t);
        if (want > c->dec_terms_cap) {
            if (c->dec_terms) cudaFree(c->dec_terms);
            c->dec_terms = nullptr;
            c->dec_terms_cap = 0;
            if (cudaMalloc(&c->dec_terms, want + want / 8) != cudaSuccess) {
                cudaGetLastError();
                return HUF_ERROR_MEMORY_ALLOCATION;
            }
            c->dec_terms_cap = want + want / 8;
        }
        a.terms = c->dec_terms;
        a.term_slots = found + 16;
    }
    if (!plan_only) {
        // dynamic shared memory: payload staging for one block (adapts to the stream's block size)
        uint64_t want = c->dec_stage_want;
        const uint64_t static_smem = sizeof(DecSmem) + 1024;
        const uint64_t max_dyn = (uint64_t)c->max_smem_optin > static_smem
                                     ? (uint64_t)c->max_smem_optin - static_smem : 20480;
        if (want > max_dyn) want = max_dyn;
        if (want < 20480) want = 20480;  // the terminal list of the table build lives here
        want &= ~uint64_t(15);
        // the attribute belongs to the function, not to this context: always allow the maximum
        // (other contexts in the process launch the same kernel with their own sizes)
        if (!c->slow_ready) {
            CU_TRY(cudaFuncSetAttribute(k_decode_slow, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(max_dyn & ~uint64_t(15))));
            c->slow_ready = true;
        }
        c->dec_stage = (uint32_t)want;
        a.stage_cap = c->dec_stage;
        int per_sm = 1;
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode_slow, kDecThreads,
                                                             c->dec_stage));
        if (per_sm < 1) per_sm = 1;
        if (!c->fast_ready) {
            // is the lookup table of k_decode 8 KB aligned in the shared window?  (It is when
            // dynamic shared memory starts at 0x400; ask the device instead of assuming.)
            if (!c->force_lut_add) {
                CTX_LAUNCH(c, k_smem_base, 1, 32, 1024, st, reinterpret_cast<uint32_t *>(c->d_result + 15));
                uint64_t base = 0;
                CU_TRY(cudaMemcpyAsync(&base, c->d_result + 15, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                CU_TRY(cudaStreamSynchronize(st));
                c->lut_or = (((uint32_t)base + (uint32_t)kFastLutOff) & (uint32_t)(kFastLutAlign - 1)) == 0;
#ifdef HUF_EMU
                c->lut_or = true;  // (addresses are host pointers there; both instances add)
#endif
            } else {
                c->lut_or = false;
            }
            // (HUF_B200_DEC_PAD: unused extra shared memory per CTA -- an occupancy experiment knob)
            const char *pad_env = getenv("HUF_B200_DEC_PAD");
            c->fast_pad = pad_env ? atoi(pad_env) & ~15 : 0;
            CU_TRY(cudaFuncSetAttribute(k_decode_unaligned, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kFastDyn + c->fast_pad));
            CU_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kFastDyn + c->fast_pad));
            int fper = 1;
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fper, k_decode, kFT, kFastDyn + c->fast_pad));
            c->fast_per_sm = fper < 1 ? 1 : fper;
            CU_TRY(cudaFuncSetAttribute(k_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, kTreeDyn));
            c->fast_ready = true;
        }
        // fast lane: tree walk (one lane per candidate), chunked region decode; whatever it
        // declines goes through the general lane
        CTX_LAUNCH(c, k_tree, c->sm_count * 6, 32, kTreeDyn, st, a);
        if (c->lut_or)
            CTX_LAUNCH(c, k_decode, c->sm_count * c->fast_per_sm, kFT, kFastDyn + c->fast_pad, st, a);
        else
            CTX_LAUNCH(c, k_decode_unaligned, c->sm_count * c->fast_per_sm, kFT, kFastDyn + c->fast_pad, st, a);
        CTX_LAUNCH(c, k_decode_slow, c->sm_count * per_sm, kDecThreads, c->dec_stage, st, a);
        CTX_LAUNCH(c, k_verify, 1, kScanThreads, 0, st, a);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(c->h_result, c->d_result, 16 * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, st));
    if (plan_only) {
        // total decoded size of the candidate chain = out_off[ncand]; fetched after the sync
    }
    return HUF_ERROR_SUCCESS;
}

}  // namespace

namespace {
huf_error_t decode_start(huf_b200_ctx_t *c, const void *d_in, uint64_t avail, uint64_t length,
                         uint64_t first, bool first_proven, void *d_out, uint64_t out_capacity,
                         void *stream)
{
    if (!c || (!d_in && avail) || (!d_out && out_capacity)) return HUF_ERROR_INVALID_ARGUMENT;
    if (c->enc_pending) return HUF_ERROR_INVALID_ARGUMENT;
    DeviceGuard g(c->device);
    if (!g.ok) return HUF_ERROR_FATAL;
    c->cur = pick_stream(c, stream);
    c->launches = 0;
    drop_timed(c);
    c->dec_dense = false;
    DecArgs &a = c->dec;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const uint8_t *>(d_in);
    a.avail = avail;
    a.length = length;
    a.out = static_cast<uint8_t *>(d_out);
    a.out_cap = out_capacity;
    // (pending only after a successful enqueue, see huf_b200_encode_async)
    const huf_error_t e = length > first ? dec_enqueue(c, first, 0, false, 0, first_proven)
                                         : HUF_ERROR_SUCCESS;  // src/decoder.c:218: nothing to consume
    c->dec_first = first;
    c->dec_first_cand = ~0ull;
    c->dec_pending = e == HUF_ERROR_SUCCESS;
    return e;
}
}  // namespace

huf_error_t huf_b200_decode_async_at(huf_b200_ctx_t *c, const void *d_in, uint64_t avail,
                                     uint64_t length, uint64_t first, void *d_out,
                                     uint64_t out_capacity, void *stream)
{
    return decode_start(c, d_in, avail, length, first, true, d_out, out_capacity, stream);
}

huf_error_t huf_b200_decode_range_async(huf_b200_ctx_t *c, const void *d_in, uint64_t avail,
                                        uint64_t start, uint64_t stop, int start_is_block,
                                        void *d_out, uint64_t out_capacity, void *stream)
{
    if (stop > avail) stop = avail;
    return decode_start(c, d_in, avail, stop, start, start_is_block != 0, d_out, out_capacity, stream);
}

huf_error_t huf_b200_decode_range_finish(huf_b200_ctx_t *c, uint64_t *first, uint64_t *end,
                                         uint64_t *out_len)
{
    if (!first || !end || !out_len) return HUF_ERROR_INVALID_ARGUMENT;
    uint64_t reached = 0;
    const huf_error_t e = huf_b200_decode_finish(c, out_len, &reached);
    *first = c ? c->dec_first_cand : ~0ull;
    *end = reached;
    return e;
}

huf_error_t huf_b200_decode_async(huf_b200_ctx_t *c, const void *d_in, uint64_t avail,
                                  uint64_t length, void *d_out, uint64_t out_capacity,
                                  void *stream)
{
    return huf_b200_decode_async_at(c, d_in, avail, length, 0, d_out, out_capacity, stream);
}

huf_error_t huf_b200_decode_hint_offsets(huf_b200_ctx_t *c, const uint64_t *d_offsets, uint64_t nblocks)
{
    if (!c) return HUF_ERROR_INVALID_ARGUMENT;
    c->hint_off = nblocks ? d_offsets : nullptr;
    c->hint_n = d_offsets ? nblocks : 0;
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_decode_finish(huf_b200_ctx_t *c, uint64_t *out_len, uint64_t *consumed)
{
    if (!c || !out_len) return HUF_ERROR_INVALID_ARGUMENT;
    if (!c->dec_pending) return HUF_ERROR_INVALID_ARGUMENT;
    c->dec_pending = false;
    *out_len = 0;
    if (consumed) *consumed = c->dec_first;
    if (c->dec.length <= c->dec_first) return HUF_ERROR_SUCCESS;
    DeviceGuard g(c->device);

    for (;;) {
        CU_TRY(cudaStreamSynchronize(c->cur));
        const uint64_t *r = c->h_result;
        if (r[8] && !c->dec_dense) {
            // more headers per chunk than the sparse single-pass scan parks: this stream has
            // tiny blocks; rerun (and keep running) with the exact two-pass scan
            c->dec_dense = true;
            huf_error_t e = dec_enqueue(c, c->dec.first, c->dec.out_base, false, 0, c->dec.first_proven != 0);
            if (e != HUF_ERROR_SUCCESS) return e;
            continue;
        }
        if (r[6] > c->dec.max_cand) {
            // candidate workspace too small for this stream: rerun the pass with the exact size
            huf_error_t e = dec_enqueue(c, c->dec.first, c->dec.out_base, false, r[6] + 16, c->dec.first_proven != 0);
            if (e != HUF_ERROR_SUCCESS) return e;
            continue;
        }
        if (c->dec_first_cand == ~0ull) c->dec_first_cand = r[12];  // (restarts begin behind it)
        if (r[9] + 64 > c->dec_stage_want) c->dec_stage_want = r[9] + 64;  // adapt staging to block size
        *out_len = r[4];
        if (consumed) *consumed = r[3];
        if (r[5]) return (huf_error_t)r[2];
        // The speculative chain broke at a block boundary that no candidate marks (foreign
        // header shape or a false positive): restart from the last proven position.  A pass
        // always proves the block at its start or reports its error, so a pass that did not
        // move cannot happen; should it ever, fail like a reader that runs dry instead of
        // spinning with the caller's lock held.
        if (r[3] <= c->dec.first) return HUF_ERROR_READ_WRITE;
        huf_error_t e = dec_enqueue(c, r[3], r[4], false, 0);
        if (e != HUF_ERROR_SUCCESS) return e;
    }
}

namespace {
huf_error_t decode_plan_from(huf_b200_ctx_t *c, const void *d_in, uint64_t avail, uint64_t length,
                             uint64_t first, bool first_proven, uint64_t *out_len,
                             uint64_t *nblocks, void *stream);
}

huf_error_t huf_b200_decode_plan(huf_b200_ctx_t *c, const void *d_in, uint64_t avail,
                                 uint64_t length, uint64_t *out_len, uint64_t *nblocks,
                                 void *stream)
{
    return decode_plan_from(c, d_in, avail, length, 0, true, out_len, nblocks, stream);
}

huf_error_t huf_b200_decode_range_plan(huf_b200_ctx_t *c, const void *d_in, uint64_t avail,
                                       uint64_t start, uint64_t stop, int start_is_block,
                                       uint64_t *out_len, uint64_t *nblocks, void *stream)
{
    if (stop > avail) stop = avail;
    return decode_plan_from(c, d_in, avail, stop, start, start_is_block != 0, out_len, nblocks, stream);
}

namespace {
huf_error_t decode_plan_from(huf_b200_ctx_t *c, const void *d_in, uint64_t avail, uint64_t length,
                             uint64_t first, bool first_proven, uint64_t *out_len,
                             uint64_t *nblocks, void *stream)
{
    if (!c || !out_len || (!d_in && avail)) return HUF_ERROR_INVALID_ARGUMENT;
    if (c->enc_pending || c->dec_pending) return HUF_ERROR_INVALID_ARGUMENT;
    DeviceGuard g(c->device);
    if (!g.ok) return HUF_ERROR_FATAL;
    c->cur = pick_stream(c, stream);
    c->launches = 0;
    *out_len = 0;
    if (nblocks) *nblocks = 0;
    if (length <= first) return HUF_ERROR_SUCCESS;
    c->dec_dense = false;
    DecArgs &a = c->dec;
    memset(&a, 0, sizeof(a));
    a.in = static_cast<const uint8_t *>(d_in);
    a.avail = avail;
    a.length = length;
    uint64_t hint = 0;
    for (;;) {
        huf_error_t e = dec_enqueue(c, first, 0, true, hint, first_proven);
        if (e != HUF_ERROR_SUCCESS) return e;
        CU_TRY(cudaStreamSynchronize(c->cur));
        if (c->h_result[8] && !c->dec_dense) {
            c->dec_dense = true;
            continue;
        }
        if (c->h_result[6] > a.max_cand) {
            hint = c->h_result[6] + 16;
            continue;
        }
        break;
    }
    const uint64_t n = c->h_result[0];
    uint64_t total = 0;
    CU_TRY(cudaMemcpy(&total, a.out_off + n, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *out_len = total;
    if (nblocks) *nblocks = n;
    if (c->h_result[9] + 64 > c->dec_stage_want) c->dec_stage_want = c->h_result[9] + 64;
    return HUF_ERROR_SUCCESS;
}
}  // namespace

// ------------------------------------------------------------------------------------------
// raw device memory helpers
// ------------------------------------------------------------------------------------------

huf_error_t huf_b200_dev_alloc(void **d_ptr, uint64_t bytes)
{
    if (!d_ptr) return HUF_ERROR_INVALID_ARGUMENT;
    CU_TRY(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_dev_free(void *d_ptr)
{
    CU_TRY(cudaFree(d_ptr));
    return HUF_ERROR_SUCCESS;
}

// Host <-> device copies of caller-owned (pageable) memory.  Large copies run as a two-deep
// pipeline over pinned bounce buffers: the DMA of one chunk overlaps the host-side memcpy of
// the next, and that memcpy is split over a few threads (a freshly allocated destination is
// first-touched by all of them instead of page-faulting on one core).
using pipe::now_s;

struct CopyTimer {
    const char *what;
    uint64_t bytes;
    double t0;
    bool on;
    CopyTimer(const char *w, uint64_t b) : what(w), bytes(b), t0(0), on(getenv("HUF_B200_DEBUG") != nullptr)
    {
        if (on) t0 = now_s();
    }
    ~CopyTimer()
    {
        if (on && bytes >= (1u << 20)) {
            const double dt = now_s() - t0;
            fprintf(stderr, "huf_b200: %s %.1f MiB in %.1f ms (%.2f GB/s)\n", what, bytes / 1048576.0, dt * 1e3,
                    bytes / dt / 1e9);
        }
    }
};

huf_error_t huf_b200_copy_h2d(void *d_dst, const void *h_src, uint64_t bytes)
{
    CopyTimer timer("copy_h2d", bytes);
    if (!bytes) return HUF_ERROR_SUCCESS;
    if (bytes < kStageMin || !stage_ready()) {
        CU_TRY(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
        return HUF_ERROR_SUCCESS;
    }
    std::lock_guard<std::mutex> lock(g_stage.mu);
    const uint8_t *src = static_cast<const uint8_t *>(h_src);
    uint8_t *dst = static_cast<uint8_t *>(d_dst);
    int k = 0;
    for (uint64_t at = 0; at < bytes; at += kStageChunk, k ^= 1) {
        const uint64_t len = bytes - at < kStageChunk ? bytes - at : kStageChunk;
        CU_TRY(cudaEventSynchronize(g_stage.done[k]));  // the DMA that last used this buffer
        parallel_memcpy(g_stage.pin[k], src + at, len);
        CU_TRY(cudaMemcpyAsync(dst + at, g_stage.pin[k], len, cudaMemcpyHostToDevice, g_stage.stream));
        CU_TRY(cudaEventRecord(g_stage.done[k], g_stage.stream));
    }
    CU_TRY(cudaStreamSynchronize(g_stage.stream));
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_copy_d2h(void *h_dst, const void *d_src, uint64_t bytes)
{
    CopyTimer timer("copy_d2h", bytes);
    if (!bytes) return HUF_ERROR_SUCCESS;
    if (bytes < kStageMin || !stage_ready()) {
        CU_TRY(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
        return HUF_ERROR_SUCCESS;
    }
    std::lock_guard<std::mutex> lock(g_stage.mu);
    uint8_t *dst = static_cast<uint8_t *>(h_dst);
    const uint8_t *src = static_cast<const uint8_t *>(d_src);
    const uint64_t nchunk = (bytes + kStageChunk - 1) / kStageChunk;
    auto issue = [&](uint64_t c) -> cudaError_t {
        const uint64_t at = c * kStageChunk;
        const uint64_t len = bytes - at < kStageChunk ? bytes - at : kStageChunk;
        cudaError_t e = cudaMemcpyAsync(g_stage.pin[c & 1], src + at, len, cudaMemcpyDeviceToHost, g_stage.stream);
        if (e == cudaSuccess) e = cudaEventRecord(g_stage.done[c & 1], g_stage.stream);
        return e;
    };
    CU_TRY(issue(0));
    for (uint64_t c = 0; c < nchunk; c++) {
        const uint64_t at = c * kStageChunk;
        const uint64_t len = bytes - at < kStageChunk ? bytes - at : kStageChunk;
        CU_TRY(cudaEventSynchronize(g_stage.done[c & 1]));
        if (c + 1 < nchunk) CU_TRY(issue(c + 1));  // next DMA overlaps this memcpy
        parallel_memcpy(dst + at, g_stage.pin[c & 1], len);
    }
    return HUF_ERROR_SUCCESS;
}


// ------------------------------------------------------------------------------------------
// page-locked caller memory
// ------------------------------------------------------------------------------------------

namespace {
struct PinnedRanges {
    std::mutex mu;
    struct R {
        const uint8_t *base;
        uint64_t len;
    };
    std::vector<R> v;
    bool covers(const void *p, uint64_t len)
    {
        const uint8_t *q = static_cast<const uint8_t *>(p);
        std::lock_guard<std::mutex> lock(mu);
        for (const R &r : v)
            if (q >= r.base && q + len <= r.base + r.len) return true;
        return false;
    }
};
PinnedRanges g_pinned;
std::atomic<uint64_t> g_direct_copies{0};  // spans / results the lanes moved without a bounce copy
}  // namespace

uint64_t huf_b200_direct_copy_count(void) { return g_direct_copies.load(); }

huf_error_t huf_b200_host_register(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes) return HUF_ERROR_INVALID_ARGUMENT;
#ifndef HUF_EMU
    if (huf_b200_device_count() <= 0) return HUF_ERROR_FATAL;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return HUF_ERROR_MEMORY_ALLOCATION;
    }
#endif
    std::lock_guard<std::mutex> lock(g_pinned.mu);
    g_pinned.v.push_back({static_cast<const uint8_t *>(ptr), bytes});
    return HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_host_unregister(void *ptr)
{
    bool known = false;
    {
        std::lock_guard<std::mutex> lock(g_pinned.mu);
        for (size_t i = 0; i < g_pinned.v.size(); i++) {
            if (g_pinned.v[i].base == ptr) {
                g_pinned.v.erase(g_pinned.v.begin() + (long)i);
                known = true;
                break;
            }
        }
    }
    if (!known) return HUF_ERROR_INVALID_ARGUMENT;
#ifndef HUF_EMU
    if (cudaHostUnregister(ptr) != cudaSuccess) {
        cudaGetLastError();
        return HUF_ERROR_FATAL;
    }
#endif
    return HUF_ERROR_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// host-buffer lanes (see host_pipe.cuh)
// ------------------------------------------------------------------------------------------

huf_error_t huf_b200_encode_host(huf_b200_ctx_t *c, const huf_b200_source_t *src, uint64_t length,
                                 uint64_t blocksize, const huf_b200_sink_t *dst, uint64_t *consumed)
{
    using namespace pipe;
    if (!c || !src || !dst) return HUF_ERROR_INVALID_ARGUMENT;
    if (c->enc_pending || c->dec_pending) return HUF_ERROR_INVALID_ARGUMENT;
    if (consumed) *consumed = 0;
    if (!length) return HUF_ERROR_SUCCESS;
    DeviceGuard g(c->device);
    if (!g.ok) return HUF_ERROR_FATAL;
    PipeState &ps = c->pipe;
    HUF_TRY_CXX(ps.init());

    const uint64_t bs = blocksize ? blocksize : length;
    uint64_t span = span_bytes() / bs * bs;  // whole blocks per span; a larger block is its own span
    if (!span) span = bs;
    if (span > length) span = length;
    const uint64_t nspans = (length + span - 1) / span;
    const int nslots = nspans < (uint64_t)slot_count() ? (int)nspans : slot_count();
    const uint64_t out_cap = huf_b200_encode_bound(span, bs);
    for (int i = 0; i < nslots; i++) {
        HUF_TRY_CXX(reserve_pinned(ps.pin_in[i], span));
        HUF_TRY_CXX(reserve_pinned(ps.pin_out[i], out_cap));
        HUF_TRY_CXX(reserve_device(ps.d_in[i], span));
        HUF_TRY_CXX(reserve_device(ps.d_out[i], out_cap));
    }

    // Page-locked caller memory is used in place (huf_b200_host_register): a contiguous source is
    // read by the H2D copy engine where it lies, a lending sink receives the D2H copy directly;
    // both save the bounce copy through the lane's own pinned buffers, i.e. half of the host
    // memory traffic of a call.
    const bool src_direct = src->data && g_pinned.covers(src->data, src->size);
    uint8_t *sink_base = nullptr;
    uint64_t sink_avail = 0, sink_off = 0;
    bool sink_direct = false;
    if (dst->room && dst->reserve && dst->commit) {
        void *p = nullptr;
        if (dst->room(dst->arg, &p, &sink_avail) == HUF_ERROR_SUCCESS && p && sink_avail >= (64u << 10) &&
            g_pinned.covers(p, sink_avail)) {
            sink_base = static_cast<uint8_t *>(p);
            sink_direct = true;
        }
    }

    struct Slot {
        uint64_t in_len = 0, out_len = 0;
        bool short_read = false;
        bool direct = false;  // the result went straight into the sink's buffer
    };
    Slot slots[kSlots];
    Chan<int> free_q, in_q, out_q;
    for (int i = 0; i < nslots; i++) free_q.push(i);
    std::atomic<int> fail{HUF_ERROR_SUCCESS};   // first error of any stage
    std::atomic<uint64_t> taken{0};
    StageTimes tm;
    const double t_start = now_s();
    const int device = c->device;

    auto set_fail = [&](huf_error_t e) {
        int ok = HUF_ERROR_SUCCESS;
        fail.compare_exchange_strong(ok, (int)e);
    };

    // ---- stage "in": source -> pinned -> HBM
    auto stage_in = [&]() {
        cudaSetDevice(device);
        for (uint64_t k = 0; k < nspans; k++) {
            const int i = free_q.pop();
            if (fail.load() != HUF_ERROR_SUCCESS) {
                free_q.push(i);
                break;
            }
            Slot &sl = slots[i];
            const uint64_t want = length - k * span < span ? length - k * span : span;
            uint64_t got = 0;
            const double t0 = now_s();
            const uint8_t *from = ps.pin_in[i].p;
            if (src_direct) {
                const uint64_t left = src->size > k * span ? src->size - k * span : 0;
                got = want < left ? want : left;
                from = static_cast<const uint8_t *>(src->data) + k * span;
                g_direct_copies.fetch_add(1);
            } else {
                huf_error_t e = source_fill(*src, k * span, ps.pin_in[i].p, want, &got);
                if (e != HUF_ERROR_SUCCESS) {
                    set_fail(e);
                    free_q.push(i);
                    break;
                }
            }
            tm.fill += now_s() - t0;
            taken.fetch_add(got);
            sl.short_read = got < want;
            sl.in_len = sl.short_read ? got / bs * bs : got;  // a short read ends after whole blocks
            sl.out_len = 0;
            sl.direct = false;
            if (sl.in_len) {
                cudaMemcpyAsync(ps.d_in[i].p, from, sl.in_len, cudaMemcpyHostToDevice, ps.s_h2d);
                cudaEventRecord(ps.ev_h2d[i], ps.s_h2d);
            }
            in_q.push(i);
            if (sl.short_read) break;
        }
        in_q.push(kEnd);
    };

    // ---- stage "out": HBM -> pinned (issued by the kernel stage) -> sink
    auto stage_out = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const int i = out_q.pop();
            if (i == kEnd) break;
            Slot &sl = slots[i];
            if (sl.out_len && fail.load() == HUF_ERROR_SUCCESS) {
                double t0 = now_s();
                if (cudaEventSynchronize(ps.ev_d2h[i]) != cudaSuccess) {
                    cudaGetLastError();
                    set_fail(HUF_ERROR_FATAL);
                } else {
                    tm.d2h_wait += now_s() - t0;
                    t0 = now_s();
                    huf_error_t e = sl.direct ? dst->commit(dst->arg, sl.out_len)
                                              : sink_deliver(*dst, ps.pin_out[i].p, sl.out_len);
                    tm.deliver += now_s() - t0;
                    if (e != HUF_ERROR_SUCCESS) set_fail(e);
                }
            }
            free_q.push(i);
        }
    };

    // ---- kernel stage (this thread)
    auto stage_kernels = [&]() {
        bool short_read = false;
        for (;;) {
            const int i = in_q.pop();
            if (i == kEnd) break;
            Slot &sl = slots[i];
            if (sl.in_len && fail.load() == HUF_ERROR_SUCCESS) {
                const double t0 = now_s();
                cudaStreamWaitEvent(c->own_stream, ps.ev_h2d[i], 0);
                // (the result also has to fit the pinned buffer it is copied to)
                const uint64_t cap = ps.d_out[i].cap < ps.pin_out[i].cap ? ps.d_out[i].cap : ps.pin_out[i].cap;
                huf_error_t e = huf_b200_encode_async(c, ps.d_in[i].p, sl.in_len, bs, ps.d_out[i].p, cap,
                                                      HUF_B200_STREAM_PRIVATE);
                uint64_t n = 0;
                if (e == HUF_ERROR_SUCCESS) e = huf_b200_encode_finish(c, &n);
                tm.kern += now_s() - t0;
                if (e != HUF_ERROR_SUCCESS) {
                    set_fail(e);
                } else {
                    sl.out_len = n;
                    // (results leave in span order, so the place of this one in the sink's buffer
                    // is known; once one does not fit the room that was there at the start the
                    // rest goes through the bounce buffers, whose delivery may move the buffer)
                    sl.direct = sink_direct && sink_off + n <= sink_avail;
                    if (!sl.direct) sink_direct = false;
                    uint8_t *to = sl.direct ? sink_base + sink_off : ps.pin_out[i].p;
                    if (sl.direct) {
                        sink_off += n;
                        g_direct_copies.fetch_add(1);
                    }
                    cudaMemcpyAsync(to, ps.d_out[i].p, n, cudaMemcpyDeviceToHost, ps.s_d2h);
                    cudaEventRecord(ps.ev_d2h[i], ps.s_d2h);
                }
            }
            short_read = short_read || sl.short_read;
            out_q.push(i);
        }
        out_q.push(kEnd);
        return short_read;
    };

    bool short_read;
    if (nspans == 1) {
        // one span: nothing to overlap, no threads
        stage_in();
        short_read = stage_kernels();
        stage_out();
    } else {
        std::thread t_in(stage_in), t_out(stage_out);
        short_read = stage_kernels();
        t_in.join();
        t_out.join();
    }
    cudaStreamSynchronize(ps.s_h2d);
    cudaStreamSynchronize(ps.s_d2h);
    if (consumed) *consumed = taken.load();
    if (debug_on() && length >= (1u << 20)) {
        const double dt = now_s() - t_start;
        fprintf(stderr,
                "huf_b200: encode_host %.1f MiB in %.1f ms (%.2f GB/s): %llu spans of %.1f MiB; busy ms: fill %.1f kernels+sync %.1f "
                "d2h wait %.1f deliver %.1f\n",
                length / 1048576.0, dt * 1e3, length / dt / 1e9, (unsigned long long)nspans, span / 1048576.0, tm.fill * 1e3,
                tm.kern * 1e3, tm.d2h_wait * 1e3, tm.deliver * 1e3);
    }
    const huf_error_t e = (huf_error_t)fail.load();
    if (e != HUF_ERROR_SUCCESS) return e;
    return short_read ? HUF_ERROR_READ_WRITE : HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_decode_host(huf_b200_ctx_t *c, const huf_b200_source_t *src, uint64_t length,
                                 const huf_b200_sink_t *dst, uint64_t *consumed)
{
    using namespace pipe;
    if (!c || !src || !dst) return HUF_ERROR_INVALID_ARGUMENT;
    if (c->enc_pending || c->dec_pending) return HUF_ERROR_INVALID_ARGUMENT;
    if (consumed) *consumed = 0;
    if (!length) return HUF_ERROR_SUCCESS;
    DeviceGuard g(c->device);
    if (!g.ok) return HUF_ERROR_FATAL;
    PipeState &ps = c->pipe;
    HUF_TRY_CXX(ps.init());

    const bool lent = src->data != nullptr;
    // bytes the in-stage brings in on its own: everything a contiguous source has, `length` bytes
    // of a pull source (more are pulled only when a block turns out to need them)
    const uint64_t planned = lent ? src->size : length;
    const uint64_t span = span_bytes();
    const uint64_t nspans = planned ? (planned + span - 1) / span : 0;
    HUF_TRY_CXX(reserve_device(ps.d_stream, planned + 64));
    for (int i = 0; i < 2; i++) HUF_TRY_CXX(reserve_pinned(ps.pin_in[i], span < planned ? span : planned));
    // output slots: a pass decodes about one span of input; the capacity adapts to the data
    uint64_t out_cap = 2 * span;
    const int nslots = nspans <= 1 ? 1 : (slot_count() > 4 ? 4 : slot_count());
    for (int i = 0; i < nslots; i++) {
        HUF_TRY_CXX(reserve_device(ps.d_out[i], out_cap));
        HUF_TRY_CXX(reserve_pinned(ps.pin_out[i], out_cap));
    }

    // (page-locked caller memory is used in place, see huf_b200_encode_host)
    const bool src_direct = lent && g_pinned.covers(src->data, src->size);
    uint8_t *sink_base = nullptr;
    uint64_t sink_avail = 0, sink_off = 0;
    bool sink_direct = false;
    if (dst->room && dst->reserve && dst->commit) {
        void *p = nullptr;
        if (dst->room(dst->arg, &p, &sink_avail) == HUF_ERROR_SUCCESS && p && sink_avail >= (64u << 10) &&
            g_pinned.covers(p, sink_avail)) {
            sink_base = static_cast<uint8_t *>(p);
            sink_direct = true;
        }
    }

    struct Slot {
        uint64_t out_len = 0;
        bool direct = false;
    };
    Slot slots[kSlots];
    Chan<int> free_q, out_q;
    for (int i = 0; i < nslots; i++) free_q.push(i);
    std::atomic<int> fail{HUF_ERROR_SUCCESS};
    StageTimes tm;
    const double t_start = now_s();
    const int device = c->device;
    auto set_fail = [&](huf_error_t e) {
        int ok = HUF_ERROR_SUCCESS;
        fail.compare_exchange_strong(ok, (int)e);
    };

    // resident = compressed bytes in HBM so far; in_done = the in-stage has brought in all it will
    std::mutex mu;
    std::condition_variable cv;
    uint64_t resident = 0;
    bool in_done = false, in_eof = false, stop_in = false;

    auto stage_in = [&]() {
        cudaSetDevice(device);
        uint64_t at = 0;
        cudaEvent_t ev[2] = {ps.ev_h2d[0], ps.ev_h2d[1]};
        uint64_t issued[2] = {0, 0};
        bool eof = false;
        for (uint64_t k = 0; k < nspans && !eof; k++) {
            const int b = (int)(k & 1);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (stop_in) break;
            }
            if (k >= 2) {
                // the copy that last used this pinned buffer is done: publish its bytes
                cudaEventSynchronize(ev[b]);
                std::lock_guard<std::mutex> lk(mu);
                resident = issued[b];
                cv.notify_all();
            }
            const uint64_t want = planned - at < span ? planned - at : span;
            uint64_t got = 0;
            const double t0 = now_s();
            const uint8_t *from = ps.pin_in[b].p;
            if (src_direct) {
                got = want;  // (planned = src->size)
                from = static_cast<const uint8_t *>(src->data) + at;
                g_direct_copies.fetch_add(1);
            } else {
                huf_error_t e = source_fill(*src, at, ps.pin_in[b].p, want, &got);
                if (e != HUF_ERROR_SUCCESS) {
                    set_fail(e);
                    break;
                }
            }
            tm.fill += now_s() - t0;
            if (got) {
                cudaMemcpyAsync(ps.d_stream.p + at, from, got, cudaMemcpyHostToDevice, ps.s_h2d);
                cudaEventRecord(ev[b], ps.s_h2d);
            }
            at += got;
            issued[b] = at;
            eof = got < want;
        }
        cudaStreamSynchronize(ps.s_h2d);
        std::lock_guard<std::mutex> lk(mu);
        resident = at;
        in_done = true;
        in_eof = eof || lent;  // a contiguous source has nothing beyond its bytes
        cv.notify_all();
    };

    auto stage_out = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const int i = out_q.pop();
            if (i == kEnd) break;
            if (slots[i].out_len && fail.load() == HUF_ERROR_SUCCESS) {
                double t0 = now_s();
                if (cudaEventSynchronize(ps.ev_d2h[i]) != cudaSuccess) {
                    cudaGetLastError();
                    set_fail(HUF_ERROR_FATAL);
                } else {
                    tm.d2h_wait += now_s() - t0;
                    t0 = now_s();
                    huf_error_t e = slots[i].direct ? dst->commit(dst->arg, slots[i].out_len)
                                                    : sink_deliver(*dst, ps.pin_out[i].p, slots[i].out_len);
                    tm.deliver += now_s() - t0;
                    if (e != HUF_ERROR_SUCCESS) set_fail(e);
                }
            }
            free_q.push(i);
        }
    };

    uint64_t done_in = 0;       // compressed bytes of the blocks decoded so far
    uint64_t npass = 0;
    huf_error_t result = HUF_ERROR_SUCCESS;

    auto stage_kernels = [&]() {
        double ratio = 1.3;     // decoded bytes per compressed byte, learnt from the passes
        for (;;) {
            if (fail.load() != HUF_ERROR_SUCCESS) break;
            // ---- wait for enough resident bytes for a worthwhile pass
            uint64_t have;
            bool all_in, eof;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return in_done || resident >= done_in + span + c->dec_margin; });
                have = resident;
                all_in = in_done;
                eof = in_eof;
            }
            if (fail.load() != HUF_ERROR_SUCCESS) break;
            // ---- blocks may start in [done_in, stop): everything while bytes are still arriving
            // except a margin in which a block would not be whole yet
            uint64_t stop = length;
            if (!all_in) {
                const uint64_t safe = have > c->dec_margin ? have - c->dec_margin : 0;
                if (safe < stop) stop = safe;
            }
            // keep the pass inside what an output slot holds
            const uint64_t fit = (uint64_t)((double)out_cap * 0.85 / ratio);
            if (stop > done_in + fit && fit >= (1u << 16)) stop = done_in + fit;
            if (stop <= done_in) {
                if (all_in) stop = length;  // (cannot happen with a margin; be safe)
                else continue;
            }
            const int i = free_q.pop();
            const double t0 = now_s();
            const uint64_t cap = ps.d_out[i].cap < ps.pin_out[i].cap ? ps.d_out[i].cap : ps.pin_out[i].cap;
            huf_error_t e = huf_b200_decode_async_at(c, ps.d_stream.p, have, stop, done_in, ps.d_out[i].p, cap,
                                                     HUF_B200_STREAM_PRIVATE);
            uint64_t n = 0, reached = done_in;
            if (e == HUF_ERROR_SUCCESS) e = huf_b200_decode_finish(c, &n, &reached);
            tm.kern += now_s() - t0;
            npass++;
            slots[i].out_len = n;
            slots[i].direct = false;
            if (n) {
                slots[i].direct = sink_direct && sink_off + n <= sink_avail;
                if (!slots[i].direct) sink_direct = false;  // (from here on delivery may move the sink's buffer)
                uint8_t *to = slots[i].direct ? sink_base + sink_off : ps.pin_out[i].p;
                if (slots[i].direct) {
                    sink_off += n;
                    g_direct_copies.fetch_add(1);
                }
                cudaMemcpyAsync(to, ps.d_out[i].p, n, cudaMemcpyDeviceToHost, ps.s_d2h);
                cudaEventRecord(ps.ev_d2h[i], ps.s_d2h);
            }
            out_q.push(i);  // (also returns an unused slot)
            const bool progressed = reached > done_in;
            if (progressed) ratio = 0.5 * ratio + 0.5 * ((double)n / (double)(reached - done_in));
            if (ratio < 0.05) ratio = 0.05;
            done_in = reached;
            if (e == HUF_ERROR_SUCCESS) {
                if (done_in >= length) break;
                continue;  // (a pass that stopped at `stop` < length: go on from there)
            }
            if (e == HUF_ERROR_MEMORY_ALLOCATION) {
                // an output slot was the limit: resume behind what was delivered; a block that does
                // not fit an empty slot needs bigger slots
                if (!progressed) {
                    const uint64_t need = c->h_result[7] + 4096;  // largest orig_len among the candidates
                    if (need <= out_cap) {
                        result = e;
                        break;
                    }
                    // slots in flight must drain before their buffers are replaced
                    int held[kSlots];
                    for (int q = 0; q < nslots; q++) held[q] = free_q.pop();
                    cudaStreamSynchronize(ps.s_d2h);
                    huf_error_t e2 = HUF_ERROR_SUCCESS;
                    for (int q = 0; q < nslots && e2 == HUF_ERROR_SUCCESS; q++) {
                        e2 = reserve_device(ps.d_out[q], need);
                        if (e2 == HUF_ERROR_SUCCESS) e2 = reserve_pinned(ps.pin_out[q], need);
                    }
                    for (int q = 0; q < nslots; q++) free_q.push(held[q]);
                    if (e2 != HUF_ERROR_SUCCESS) {
                        result = e2;
                        break;
                    }
                    out_cap = need;
                }
                continue;
            }
            if (e == HUF_ERROR_READ_WRITE) {
                // the failing block may simply continue in bytes that are not resident yet
                if (!all_in) {
                    // wait for more bytes than this pass saw; keep a wider margin from now on
                    const uint64_t ext = c->h_result[9];  // largest block extent seen
                    if (2 * ext + 65536 > c->dec_margin) c->dec_margin = 2 * ext + 65536;
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return in_done || resident > have; });
                    continue;
                }
                if (!eof) {
                    // pull source: the in-stage is finished, so this thread pulls more itself
                    const uint64_t more = have > 65536 ? have : 65536;
                    huf_error_t e2 = reserve_device(ps.d_stream, have + more + 64, true, have);
                    uint64_t got = 0, at = have;
                    while (e2 == HUF_ERROR_SUCCESS && at < have + more) {
                        const uint64_t want = have + more - at < ps.pin_in[0].cap ? have + more - at : ps.pin_in[0].cap;
                        e2 = source_fill(*src, at, ps.pin_in[0].p, want, &got);
                        if (e2 != HUF_ERROR_SUCCESS || !got) break;
                        cudaMemcpyAsync(ps.d_stream.p + at, ps.pin_in[0].p, got, cudaMemcpyHostToDevice, ps.s_h2d);
                        cudaStreamSynchronize(ps.s_h2d);
                        at += got;
                        if (got < want) break;
                    }
                    if (e2 != HUF_ERROR_SUCCESS) {
                        result = e2;
                        break;
                    }
                    std::lock_guard<std::mutex> lk(mu);
                    in_eof = at < have + more;
                    if (at > resident) {
                        resident = at;
                        continue;
                    }
                }
            }
            result = e;
            break;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            stop_in = true;
        }
        out_q.push(kEnd);
    };

    if (nspans <= 1) {
        stage_in();
        stage_kernels();
        stage_out();
    } else {
        std::thread t_in(stage_in), t_out(stage_out);
        stage_kernels();
        t_in.join();
        t_out.join();
    }
    cudaStreamSynchronize(ps.s_h2d);
    cudaStreamSynchronize(ps.s_d2h);
    if (consumed) *consumed = done_in;
    if (debug_on() && planned >= (1u << 20)) {
        const double dt = now_s() - t_start;
        fprintf(stderr,
                "huf_b200: decode_host %.1f MiB in %.1f ms (%.2f GB/s compressed): %llu passes; busy ms: fill %.1f kernels+sync %.1f "
                "d2h wait %.1f deliver %.1f\n",
                planned / 1048576.0, dt * 1e3, planned / dt / 1e9, (unsigned long long)npass, tm.fill * 1e3, tm.kern * 1e3,
                tm.d2h_wait * 1e3, tm.deliver * 1e3);
    }
    const huf_error_t e = (huf_error_t)fail.load();
    return e != HUF_ERROR_SUCCESS ? e : result;
}


// ------------------------------------------------------------------------------------------
// several GPUs, one call: contiguous block ranges per device for encode,
// contiguous byte ranges of the ONE stream per device for decode; only sizes and block
// positions travel between the device threads, the host places the slabs at the exclusive
// scan of their sizes.  No device-to-device traffic, no collective.
// ------------------------------------------------------------------------------------------

namespace {

// One thread per device meets here between the phases of a call.
class Rendezvous {
public:
    explicit Rendezvous(int n) : n_(n) {}
    void arrive()
    {
        std::unique_lock<std::mutex> lk(mu_);
        const uint64_t gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            gen_++;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0;
    uint64_t gen_ = 0;
};

// host memory -> device memory through the context's two pinned buffers (fill of piece k+1
// overlaps the DMA of piece k)
huf_error_t upload(huf_b200_ctx *c, uint8_t *d_dst, const uint8_t *h_src, uint64_t bytes)
{
    using namespace pipe;
    PipeState &ps = c->pipe;
    const uint64_t piece = span_bytes();
    for (int i = 0; i < 2; i++) HUF_TRY_CXX(reserve_pinned(ps.pin_in[i], piece < bytes ? piece : bytes));
    uint64_t k = 0;
    for (uint64_t at = 0; at < bytes; at += piece, k++) {
        const int b = (int)(k & 1);
        const uint64_t len = bytes - at < piece ? bytes - at : piece;
        if (k >= 2) CU_TRY(cudaEventSynchronize(ps.ev_h2d[b]));
        CopyPool::get().copy(ps.pin_in[b].p, h_src + at, len);
        CU_TRY(cudaMemcpyAsync(d_dst + at, ps.pin_in[b].p, len, cudaMemcpyHostToDevice, ps.s_h2d));
        CU_TRY(cudaEventRecord(ps.ev_h2d[b], ps.s_h2d));
    }
    CU_TRY(cudaStreamSynchronize(ps.s_h2d));
    return HUF_ERROR_SUCCESS;
}

// device memory -> host memory, same scheme the other way round
huf_error_t download(huf_b200_ctx *c, uint8_t *h_dst, const uint8_t *d_src, uint64_t bytes)
{
    using namespace pipe;
    PipeState &ps = c->pipe;
    const uint64_t piece = span_bytes();
    for (int i = 0; i < 2; i++) HUF_TRY_CXX(reserve_pinned(ps.pin_out[i], piece < bytes ? piece : bytes));
    const uint64_t npieces = (bytes + piece - 1) / piece;
    auto issue = [&](uint64_t k) -> cudaError_t {
        const uint64_t at = k * piece;
        const uint64_t len = bytes - at < piece ? bytes - at : piece;
        cudaError_t e = cudaMemcpyAsync(ps.pin_out[k & 1].p, d_src + at, len, cudaMemcpyDeviceToHost, ps.s_d2h);
        if (e == cudaSuccess) e = cudaEventRecord(ps.ev_d2h[k & 1], ps.s_d2h);
        return e;
    };
    if (npieces) CU_TRY(issue(0));
    for (uint64_t k = 0; k < npieces; k++) {
        const uint64_t at = k * piece;
        const uint64_t len = bytes - at < piece ? bytes - at : piece;
        CU_TRY(cudaEventSynchronize(ps.ev_d2h[k & 1]));
        if (k + 1 < npieces) CU_TRY(issue(k + 1));  // next DMA overlaps this copy
        CopyPool::get().copy(h_dst + at, ps.pin_out[k & 1].p, len);
        // (the buffer of piece k is refilled by piece k+2, which is issued after this copy)
    }
    return HUF_ERROR_SUCCESS;
}

}  // namespace

huf_error_t huf_b200_encode_host_multi(huf_b200_ctx_t *const *ctxs, int ndev, const void *h_in,
                                       uint64_t length, uint64_t blocksize,
                                       const huf_b200_sink_t *dst, uint64_t *slab_sizes)
{
    using namespace pipe;
    if (!ctxs || ndev < 1 || ndev > 64 || !dst || !dst->reserve || (!h_in && length)) return HUF_ERROR_INVALID_ARGUMENT;
    if (!length) return HUF_ERROR_SUCCESS;
    const uint64_t bs = blocksize ? blocksize : length;
    const uint64_t nblocks = huf_b200_block_count(length, bs);
    std::vector<uint64_t> size(ndev, 0), off(ndev + 1, 0);
    std::vector<huf_error_t> err(ndev, HUF_ERROR_SUCCESS);
    uint8_t *base = nullptr;
    Rendezvous meet(ndev);

    auto worker = [&](int g) {
        huf_b200_ctx *c = ctxs[g];
        // contiguous block range of device g (ranges differ by at most one block)
        const uint64_t q = nblocks / ndev, r = nblocks % ndev;
        const uint64_t b_lo = g * q + ((uint64_t)g < r ? g : r), b_hi = b_lo + q + ((uint64_t)g < r ? 1 : 0);
        const uint64_t lo = b_lo * bs < length ? b_lo * bs : length, hi = b_hi * bs < length ? b_hi * bs : length;
        const uint64_t n = hi - lo;
        huf_error_t e = HUF_ERROR_SUCCESS;
        DeviceGuard guard(c->device);
        PipeState &ps = c->pipe;
        if (n) {
            e = ps.init();
            if (e == HUF_ERROR_SUCCESS) e = reserve_device(ps.d_in[0], n);
            if (e == HUF_ERROR_SUCCESS) e = reserve_device(ps.d_out[0], huf_b200_encode_bound(n, bs));
            if (e == HUF_ERROR_SUCCESS) e = upload(c, ps.d_in[0].p, static_cast<const uint8_t *>(h_in) + lo, n);
            if (e == HUF_ERROR_SUCCESS)
                e = huf_b200_encode_async(c, ps.d_in[0].p, n, bs, ps.d_out[0].p, ps.d_out[0].cap, HUF_B200_STREAM_PRIVATE);
            if (e == HUF_ERROR_SUCCESS) e = huf_b200_encode_finish(c, &size[g]);
        }
        err[g] = e;
        meet.arrive();
        if (g == 0) {
            // exclusive scan of the slab sizes: where every slab goes in the one stream
            bool ok = true;
            for (int k = 0; k < ndev; k++) {
                ok = ok && err[k] == HUF_ERROR_SUCCESS;
                off[k + 1] = off[k] + size[k];
            }
            if (ok) {
                void *p = nullptr;
                const huf_error_t e2 = dst->reserve(dst->arg, off[ndev], &p);
                if (e2 != HUF_ERROR_SUCCESS || !p) err[0] = e2 != HUF_ERROR_SUCCESS ? e2 : HUF_ERROR_INVALID_ARGUMENT;
                base = static_cast<uint8_t *>(p);
            }
        }
        meet.arrive();
        if (base && size[g]) err[g] = download(c, base + off[g], ps.d_out[0].p, size[g]);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < ndev; g++) th.emplace_back(worker, g);
    worker(0);
    for (auto &t : th) t.join();
    for (int g = 0; g < ndev; g++) {
        if (err[g] != HUF_ERROR_SUCCESS) return err[g];
    }
    if (slab_sizes) {
        for (int g = 0; g < ndev; g++) slab_sizes[g] = size[g];
    }
    return dst->commit ? dst->commit(dst->arg, off[ndev]) : HUF_ERROR_SUCCESS;
}

huf_error_t huf_b200_decode_host_multi(huf_b200_ctx_t *const *ctxs, int ndev, const void *h_in,
                                       uint64_t avail, uint64_t length, const huf_b200_sink_t *dst,
                                       uint64_t *consumed)
{
    using namespace pipe;
    if (!ctxs || ndev < 1 || ndev > 64 || !dst || !dst->reserve || (!h_in && avail)) return HUF_ERROR_INVALID_ARGUMENT;
    if (consumed) *consumed = 0;
    if (!length) return HUF_ERROR_SUCCESS;
    const uint64_t lim = length < avail ? length : avail;  // blocks start in front of lim
    struct Part {
        uint64_t first = ~0ull, end = 0, out = 0;
        uint64_t hi = 0;         // end of the byte range
        bool saw_all = false;    // the device held the stream up to its very end
        huf_error_t err = HUF_ERROR_SUCCESS;
    };
    std::vector<Part> part(ndev);
    std::vector<uint64_t> out_off(ndev + 1, 0);
    uint8_t *base = nullptr;
    bool stitched = false;
    huf_error_t final_err = HUF_ERROR_SUCCESS;
    int last_dev = -1;       // devices 0..last_dev deliver their slabs
    uint64_t reached = 0;
    Rendezvous meet(ndev);

    auto worker = [&](int g) {
        huf_b200_ctx *c = ctxs[g];
        Part &me = part[g];
        // byte range of device g, and the bytes behind it that its last block may need
        const uint64_t lo = lim / ndev * g, hi = g + 1 == ndev ? lim : lim / ndev * (g + 1);
        const uint64_t lo16 = lo & ~uint64_t(15);  // the device copy keeps the stream's 16-byte phase
        uint64_t over = (hi - lo) / 4;
        if (over < (8ull << 20)) over = 8ull << 20;
        const uint64_t up_hi = avail - hi < over ? avail : hi + over;
        me.hi = hi;
        me.saw_all = up_hi == avail;
        DeviceGuard guard(c->device);
        PipeState &ps = c->pipe;
        huf_error_t e = HUF_ERROR_SUCCESS;
        if (hi > lo) {
            e = ps.init();
            if (e == HUF_ERROR_SUCCESS) e = reserve_device(ps.d_stream, up_hi - lo16 + 64);
            if (e == HUF_ERROR_SUCCESS) e = upload(c, ps.d_stream.p, static_cast<const uint8_t *>(h_in) + lo16, up_hi - lo16);
            uint64_t est = 0;
            if (e == HUF_ERROR_SUCCESS)
                e = huf_b200_decode_range_plan(c, ps.d_stream.p, up_hi - lo16, lo - lo16, hi - lo16, lo == 0, &est,
                                               nullptr, HUF_B200_STREAM_PRIVATE);
            if (e == HUF_ERROR_SUCCESS) e = reserve_device(ps.d_out[0], est + 64);
            if (e == HUF_ERROR_SUCCESS)
                e = huf_b200_decode_range_async(c, ps.d_stream.p, up_hi - lo16, lo - lo16, hi - lo16, lo == 0,
                                                ps.d_out[0].p, ps.d_out[0].cap, HUF_B200_STREAM_PRIVATE);
            if (e == HUF_ERROR_SUCCESS) {
                e = huf_b200_decode_range_finish(c, &me.first, &me.end, &me.out);
                if (me.first != ~0ull) me.first += lo16;
                me.end += lo16;
            }
        }
        me.err = e;
        meet.arrive();
        if (g == 0) {
            // chain validation across the ranges: every range must begin where the chain of
            // its predecessors ended; an error of the stream stops the chain there (the output
            // of the blocks before it is still delivered, like src/decoder.c:218-276 does)
            uint64_t expect = 0;
            stitched = true;
            for (int k = 0; k < ndev; k++) {
                const Part &p = part[k];
                out_off[k + 1] = out_off[k];
                if (k > 0 && expect >= p.hi) continue;  // a block of an earlier range covers this one
                if (p.first != expect) {
                    // a header the scan does not recognise, or a false one, at the seam (also:
                    // device error before any block): the serial lane decides
                    stitched = false;
                    break;
                }
                last_dev = k;
                out_off[k + 1] = out_off[k] + p.out;
                expect = p.end;
                if (p.err != HUF_ERROR_SUCCESS) {
                    // (READ_WRITE with bytes left behind the device's copy: the block outgrew the
                    // overlap, not the stream)
                    if (p.err == HUF_ERROR_READ_WRITE && !p.saw_all) stitched = false;
                    final_err = p.err;
                    break;
                }
            }
            for (int k = last_dev + 1; k < ndev; k++) out_off[k + 1] = out_off[last_dev + 1];
            reached = expect;
            if (stitched && final_err == HUF_ERROR_SUCCESS && reached < lim) stitched = false;  // (cannot happen)
            if (stitched && out_off[last_dev + 1]) {
                void *p = nullptr;
                const huf_error_t e2 = dst->reserve(dst->arg, out_off[last_dev + 1], &p);
                if (e2 != HUF_ERROR_SUCCESS || !p) {
                    final_err = e2 != HUF_ERROR_SUCCESS ? e2 : HUF_ERROR_INVALID_ARGUMENT;
                    stitched = false;
                    last_dev = -2;  // nothing to deliver, no fallback either
                }
                base = static_cast<uint8_t *>(p);
            }
        }
        meet.arrive();
        if (stitched && base && g <= last_dev && me.out) {
            const huf_error_t e2 = download(c, base + out_off[g], ps.d_out[0].p, me.out);
            if (e2 != HUF_ERROR_SUCCESS) me.err = e2;
        }
    };
    // (the seam rule needs to know whether device k saw the whole rest of the stream)
    std::vector<std::thread> th;
    for (int g = 1; g < ndev; g++) th.emplace_back(worker, g);
    worker(0);
    for (auto &t : th) t.join();

    if (debug_on())
        fprintf(stderr, "huf_b200: decode_host_multi over %d devices: %s, chain reached %llu of %llu\n", ndev,
                stitched ? "seams validated" : "a seam did not validate: serial lane", (unsigned long long)reached,
                (unsigned long long)lim);
    if (debug_on()) {
        for (int k = 0; k < ndev; k++)
            fprintf(stderr, "huf_b200:   range %d: ends %llu first %lld chain end %llu out %llu err %d\n", k,
                    (unsigned long long)part[k].hi, (long long)part[k].first, (unsigned long long)part[k].end,
                    (unsigned long long)part[k].out, (int)part[k].err);
    }
    if (last_dev == -2) return final_err;
    if (!stitched) {
        // correctness first: the one-device lane decodes the stream serially span by span
        huf_b200_source_t src;
        memset(&src, 0, sizeof(src));
        src.data = h_in;
        src.size = avail;
        return huf_b200_decode_host(ctxs[0], &src, length, dst, consumed);
    }
    for (int g = 0; g <= last_dev; g++) {
        if (part[g].err != HUF_ERROR_SUCCESS && part[g].err != final_err) return part[g].err;
    }
    if (consumed) *consumed = reached;
    if (dst->commit && out_off[last_dev + 1]) {
        const huf_error_t e2 = dst->commit(dst->arg, out_off[last_dev + 1]);
        if (e2 != HUF_ERROR_SUCCESS) return e2;
    }
    if (final_err == HUF_ERROR_SUCCESS && reached >= avail && reached < length) return HUF_ERROR_READ_WRITE;
    return final_err;
}

}  // extern "C"

#ifdef HUF_PHASE_PROF
// Debug builds only (scripts/phase_prof.py): per-phase cycle sums of k_decode.
extern "C" int huf_b200_debug_phase(unsigned long long *out, int reset)
{
    using namespace hufb200;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (out && cudaMemcpyFromSymbol(out, g_fast_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return -2;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_fast_prof, z, sizeof z) != cudaSuccess) return -3;
    }
    return 0;
}
#endif
