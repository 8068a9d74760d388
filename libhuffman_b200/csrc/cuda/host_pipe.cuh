// host_pipe.cuh — the host-buffer lanes of the codec: huf_b200_encode_host / huf_b200_decode_host.
//
// huf_encode / huf_decode hand their reader and writer to these two calls as a byte source and
// a byte sink (reference: the stream callbacks of src/io.c:9-226 behind src/bufio.c:149-287).
// The input is cut into spans of whole blocks and three stages run concurrently on three
// host threads and three CUDA streams:
//
//     source -> pinned buffer -> HBM        (stage thread "in":  pull/memcpy + H2D copy engine)
//     kernels of span k                     (calling thread:     the context's compute stream)
//     HBM -> pinned buffer -> sink          (stage thread "out": D2H copy engine + memcpy/push)
//
// so the PCIe transfers of both directions, the host-side copies and the kernels of neighbouring
// spans overlap.  Host-side copies between caller memory and the pinned buffers are split over
// a persistent pool of copy threads (pageable memory cannot be DMA'd, and registering caller
// memory per call costs more than copying it: cudaHostRegister runs at 3-4 GB/s, the copy pool
// at 50-75 GB/s on the 16-core B200 host).  Pinned and device buffers are cached per process.
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

extern "C" void huf__copy_stream(void *dst, const void *src, size_t n);  // csrc/host/ntcopy.c

namespace hufb200 {
namespace pipe {

// HUF_B200_NT_COPY=0 switches the staging copies back to plain memcpy (measurement aid).
inline void stage_copy(void *dst, const void *src, size_t n)
{
    static const bool nt = [] {
        const char *e = getenv("HUF_B200_NT_COPY");
        return !(e && e[0] == '0');
    }();
    if (nt) huf__copy_stream(dst, src, n);
    else memcpy(dst, src, n);
}

inline double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

inline bool debug_on()
{
    static const bool on = getenv("HUF_B200_DEBUG") != nullptr;
    return on;
}

// ---- persistent copy threads ------------------------------------------------------------------

class CopyPool {
public:
    static CopyPool &get()
    {
        static CopyPool pool;
        return pool;
    }

    // memcpy split over the pool; the caller works on slices too and returns when all are done
    void copy(void *dst, const void *src, uint64_t bytes)
    {
        constexpr uint64_t kSlice = 2ull << 20;
        if (bytes < 2 * kSlice || nthreads_ == 0) {
            stage_copy(dst, src, bytes);
            return;
        }
        Job job;
        job.dst = static_cast<uint8_t *>(dst);
        job.src = static_cast<const uint8_t *>(src);
        job.bytes = bytes;
        job.slice = kSlice;
        job.nslices = (bytes + kSlice - 1) / kSlice;
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        work_on(job);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return job.done.load() == job.nslices; });
        // (finished jobs are unlinked by whoever takes the last slice or by us here)
        for (auto it = jobs_.begin(); it != jobs_.end(); ++it) {
            if (*it == &job) {
                jobs_.erase(it);
                break;
            }
        }
    }

    unsigned threads() const { return nthreads_ + 1; }

private:
    struct Job {
        uint8_t *dst;
        const uint8_t *src;
        uint64_t bytes, slice, nslices;
        std::atomic<uint64_t> next{0}, done{0};
    };

    CopyPool()
    {
        unsigned hc = std::thread::hardware_concurrency();
        const char *env = getenv("HUF_B200_COPY_THREADS");
        // default: all cores but two (the stage threads of a call need some), at most 16
        unsigned want = env ? (unsigned)atoi(env) : (hc > 3 ? (hc - 2 > 16 ? 16u : hc - 2) : 1u);
        if (hc && want > hc) want = hc;
        if (want < 1) want = 1;
        nthreads_ = want - 1;  // the calling thread is one of the workers
        for (unsigned i = 0; i < nthreads_; i++) workers_.emplace_back([this] { run(); });
    }

    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }

    void work_on(Job &job)
    {
        for (;;) {
            const uint64_t i = job.next.fetch_add(1);
            if (i >= job.nslices) return;
            const uint64_t at = i * job.slice;
            const uint64_t len = job.bytes - at < job.slice ? job.bytes - at : job.slice;
            stage_copy(job.dst + at, job.src + at, len);
            if (job.done.fetch_add(1) + 1 == job.nslices) {
                std::lock_guard<std::mutex> lk(mu_);
                done_cv_.notify_all();
            }
        }
    }

    void run()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            Job *job = nullptr;
            for (Job *j : jobs_) {
                if (j->next.load() < j->nslices) {
                    job = j;
                    break;
                }
            }
            if (!job) {
                if (stop_) return;
                cv_.wait(lk);
                continue;
            }
            // the job stays alive while slices are open: its owner waits for done == nslices,
            // and a slice is only counted done after its memcpy
            const uint64_t i = job->next.fetch_add(1);
            if (i >= job->nslices) continue;
            lk.unlock();
            const uint64_t at = i * job->slice;
            const uint64_t len = job->bytes - at < job->slice ? job->bytes - at : job->slice;
            stage_copy(job->dst + at, job->src + at, len);
            const bool last = job->done.fetch_add(1) + 1 == job->nslices;
            lk.lock();
            if (last) done_cv_.notify_all();
        }
    }

    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Job *> jobs_;
    std::vector<std::thread> workers_;
    unsigned nthreads_ = 0;
    bool stop_ = false;
};

// ---- a small blocking queue -------------------------------------------------------------------

template <typename T>
class Chan {
public:
    void push(T v)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(v);
        }
        cv_.notify_one();
    }
    T pop()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !q_.empty(); });
        T v = q_.front();
        q_.pop_front();
        return v;
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

// ---- buffers cached across calls (one set per device) -------------------------------------------

constexpr int kSlots = 6;   // spans in flight at most (fill, H2D, kernels, D2H, deliver overlap)
constexpr int kEnd = -1;

struct Buf {
    uint8_t *p = nullptr;
    uint64_t cap = 0;
};

inline huf_error_t reserve_pinned(Buf &b, uint64_t want)
{
    if (want <= b.cap) return HUF_ERROR_SUCCESS;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    want = (want + (want >> 3) + 4095) & ~uint64_t(4095);
    if (cudaHostAlloc(reinterpret_cast<void **>(&b.p), want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return HUF_ERROR_MEMORY_ALLOCATION;
    }
    b.cap = want;
    return HUF_ERROR_SUCCESS;
}

inline huf_error_t reserve_device(Buf &b, uint64_t want, bool keep = false, uint64_t keep_bytes = 0)
{
    if (want <= b.cap) return HUF_ERROR_SUCCESS;
    want = (want + (want >> 3) + 4095) & ~uint64_t(4095);
    uint8_t *fresh = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&fresh), want) != cudaSuccess) {
        cudaGetLastError();
        return HUF_ERROR_MEMORY_ALLOCATION;
    }
    if (keep && b.p && keep_bytes) cudaMemcpy(fresh, b.p, keep_bytes, cudaMemcpyDeviceToDevice);
    if (b.p) cudaFree(b.p);
    b.p = fresh;
    b.cap = want;
    return HUF_ERROR_SUCCESS;
}

struct PipeState {
    bool ready = false;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[kSlots], ev_d2h[kSlots];
    Buf pin_in[kSlots], pin_out[kSlots], d_in[kSlots], d_out[kSlots];
    Buf d_stream;  // decode: the whole compressed stream

    huf_error_t init()
    {
        if (ready) return HUF_ERROR_SUCCESS;
        cudaError_t e = cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking);
        for (int i = 0; i < kSlots && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&ev_h2d[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_d2h[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return HUF_ERROR_FATAL;
        }
        ready = true;
        return HUF_ERROR_SUCCESS;
    }

    void release()
    {
        for (int i = 0; i < kSlots; i++) {
            if (pin_in[i].p) cudaFreeHost(pin_in[i].p);
            if (pin_out[i].p) cudaFreeHost(pin_out[i].p);
            if (d_in[i].p) cudaFree(d_in[i].p);
            if (d_out[i].p) cudaFree(d_out[i].p);
            pin_in[i] = pin_out[i] = d_in[i] = d_out[i] = Buf();
        }
        if (d_stream.p) cudaFree(d_stream.p);
        d_stream = Buf();
        if (ready) {
            for (int i = 0; i < kSlots; i++) {
                cudaEventDestroy(ev_h2d[i]);
                cudaEventDestroy(ev_d2h[i]);
            }
            cudaStreamDestroy(s_h2d);
            cudaStreamDestroy(s_d2h);
        }
        ready = false;
    }
};

// Bytes per span: HUF_B200_SPAN_MIB (default 32), or HUF_B200_SPAN_BYTES for tests that want
// many spans out of a small input.  Read at every call (cheap) so a process can change it.
// Slots used by a call: HUF_B200_SLOTS (2..kSlots, default 5).
inline int slot_count()
{
    const char *env = getenv("HUF_B200_SLOTS");
    int n = env ? atoi(env) : 5;
    if (n < 2) n = 2;
    if (n > kSlots) n = kSlots;
    return n;
}

inline uint64_t span_bytes()
{
    if (const char *env = getenv("HUF_B200_SPAN_BYTES")) {
        const uint64_t v = (uint64_t)atoll(env);
        if (v >= 64) return v;
    }
    const char *env = getenv("HUF_B200_SPAN_MIB");
    uint64_t mib = env ? (uint64_t)atoll(env) : 32;
    if (mib < 1) mib = 1;
    if (mib > 4096) mib = 4096;
    return mib << 20;
}

// Fill `dst` with up to `want` bytes of the source, starting at source offset `at` (data mode)
// or simply the next bytes (pull mode).
inline huf_error_t source_fill(const huf_b200_source_t &src, uint64_t at, uint8_t *dst, uint64_t want,
                               uint64_t *got)
{
    *got = 0;
    if (src.data) {
        const uint64_t left = src.size > at ? src.size - at : 0;
        const uint64_t n = want < left ? want : left;
        if (n) CopyPool::get().copy(dst, static_cast<const uint8_t *>(src.data) + at, n);
        *got = n;
        return HUF_ERROR_SUCCESS;
    }
    if (!src.pull) return HUF_ERROR_INVALID_ARGUMENT;
    return src.pull(src.arg, dst, want, got);
}

// Deliver `count` bytes at `p` (pinned) to the sink.
inline huf_error_t sink_deliver(const huf_b200_sink_t &sink, const uint8_t *p, uint64_t count)
{
    if (!count) return HUF_ERROR_SUCCESS;
    if (sink.reserve) {
        void *dst = nullptr;
        huf_error_t e = sink.reserve(sink.arg, count, &dst);
        if (e != HUF_ERROR_SUCCESS) return e;
        if (dst) {
            CopyPool::get().copy(dst, p, count);
            return sink.commit ? sink.commit(sink.arg, count) : HUF_ERROR_SUCCESS;
        }
    }
    if (!sink.push) return HUF_ERROR_INVALID_ARGUMENT;
    return sink.push(sink.arg, p, count);
}

struct StageTimes {
    double fill = 0, h2d_wait = 0, kern = 0, d2h_wait = 0, deliver = 0;
};

}  // namespace pipe
}  // namespace hufb200
