// probe_host.cu — host<->device staging options measured on the GPU box (decides the e2e design).
// Build: nvcc -O3 -o scripts/_bin/probe_host scripts/probe_host.cu -lpthread
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void par_memcpy(void *d, const void *s, size_t n, int nt)
{
    std::vector<std::thread> th;
    size_t sl = ((n + nt - 1) / nt + 4095) & ~size_t(4095);
    for (int t = 0; t < nt; t++) {
        size_t at = (size_t)t * sl;
        if (at >= n) break;
        size_t len = n - at < sl ? n - at : sl;
        th.emplace_back([=] { memcpy((char *)d + at, (const char *)s + at, len); });
    }
    for (auto &x : th) x.join();
}

int main()
{
    const size_t N = 1ull << 30;
    void *d; cudaMalloc(&d, N);
    cudaStream_t st; cudaStreamCreate(&st);
    printf("cores %u\n", std::thread::hardware_concurrency());

    // pinned baseline
    void *pin; cudaHostAlloc(&pin, N, cudaHostAllocDefault);
    memset(pin, 1, N);
    for (int r = 0; r < 2; r++) {
        double t0 = now(); cudaMemcpyAsync(d, pin, N, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
        double t1 = now(); cudaMemcpyAsync(pin, d, N, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        double t2 = now();
        printf("pinned 1 GiB: H2D %.1f GB/s  D2H %.1f GB/s\n", N / (t1 - t0) / 1e9, N / (t2 - t1) / 1e9);
    }
    // full duplex
    {
        void *d2; cudaMalloc(&d2, N); void *pin2; cudaHostAlloc(&pin2, N, cudaHostAllocDefault); memset(pin2, 2, N);
        cudaStream_t s2; cudaStreamCreate(&s2);
        double t0 = now();
        cudaMemcpyAsync(d, pin, N, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(pin2, d2, N, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(st); cudaStreamSynchronize(s2);
        double t1 = now();
        printf("pinned duplex 1+1 GiB: %.1f GB/s each direction\n", N / (t1 - t0) / 1e9);
        cudaFree(d2); cudaFreeHost(pin2);
    }
    // calloc'd buffer (touched), pageable cudaMemcpy
    char *pg = (char *)calloc(N, 1);
    memset(pg, 3, N);
    {
        double t0 = now(); cudaMemcpy(d, pg, N, cudaMemcpyHostToDevice);
        double t1 = now(); cudaMemcpy(pg, d, N, cudaMemcpyDeviceToHost);
        double t2 = now();
        printf("pageable cudaMemcpy: H2D %.1f GB/s  D2H %.1f GB/s\n", N / (t1 - t0) / 1e9, N / (t2 - t1) / 1e9);
    }
    // cudaHostRegister of a touched buffer
    for (int r = 0; r < 2; r++) {
        double t0 = now(); cudaError_t e = cudaHostRegister(pg, N, cudaHostRegisterDefault);
        double t1 = now();
        cudaMemcpyAsync(d, pg, N, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st);
        double t2 = now(); cudaHostUnregister(pg);
        double t3 = now();
        printf("register touched 1 GiB: %s %.1f ms (%.1f GB/s), H2D %.1f GB/s, unregister %.1f ms\n", cudaGetErrorName(e),
               (t1 - t0) * 1e3, N / (t1 - t0) / 1e9, N / (t2 - t1) / 1e9, (t3 - t2) * 1e3);
    }
    // register in 64 MiB pieces (pipelinable)
    {
        const size_t P = 64ull << 20;
        double t0 = now();
        for (size_t at = 0; at < N; at += P) cudaHostRegister(pg + at, P, cudaHostRegisterDefault);
        double t1 = now();
        for (size_t at = 0; at < N; at += P) cudaHostUnregister(pg + at);
        double t2 = now();
        printf("register in 64 MiB pieces: %.1f ms total (%.1f GB/s), unregister %.1f ms\n", (t1 - t0) * 1e3, N / (t1 - t0) / 1e9, (t2 - t1) * 1e3);
    }
    // register a FRESH calloc (untouched pages) then D2H into it
    {
        char *fr = (char *)calloc(N, 1);
        double t0 = now(); cudaError_t e = cudaHostRegister(fr, N, cudaHostRegisterDefault);
        double t1 = now(); cudaMemcpyAsync(fr, d, N, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        double t2 = now(); cudaHostUnregister(fr);
        printf("register fresh calloc 1 GiB: %s %.1f ms (%.1f GB/s), D2H %.1f GB/s\n", cudaGetErrorName(e), (t1 - t0) * 1e3,
               N / (t1 - t0) / 1e9, N / (t2 - t1) / 1e9);
        free(fr);
    }
    // fresh calloc with MADV_HUGEPAGE + first touch by threads
    for (int nt : {1, 4, 8, 16}) {
        char *fr = (char *)aligned_alloc(2 << 20, N);
        madvise(fr, N, MADV_HUGEPAGE);
        double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([=] { memset(fr + (N / nt) * t, 0, N / nt); });
        for (auto &x : th) x.join();
        double t1 = now();
        printf("first touch (hugepage hint) 1 GiB with %d threads: %.1f ms (%.1f GB/s)\n", nt, (t1 - t0) * 1e3, N / (t1 - t0) / 1e9);
        double t2 = now(); cudaError_t e = cudaHostRegister(fr, N, cudaHostRegisterDefault); double t3 = now();
        printf("   then register: %s %.1f ms\n", cudaGetErrorName(e), (t3 - t2) * 1e3);
        cudaHostUnregister(fr);
        free(fr);
    }
    // threaded memcpy pageable -> pinned
    for (int nt : {1, 2, 4, 8, 16}) {
        double t0 = now(); par_memcpy(pin, pg, N, nt); double t1 = now();
        par_memcpy(pg, pin, N, nt); double t2 = now();
        printf("memcpy %2d threads: pageable->pinned %.1f GB/s, pinned->pageable %.1f GB/s\n", nt, N / (t1 - t0) / 1e9, N / (t2 - t1) / 1e9);
    }
    // memcpy into FRESH pages
    for (int nt : {8, 16}) {
        char *fr = (char *)calloc(N, 1);
        double t0 = now(); par_memcpy(fr, pin, N, nt); double t1 = now();
        printf("memcpy %2d threads pinned->fresh calloc: %.1f GB/s\n", nt, N / (t1 - t0) / 1e9);
        free(fr);
    }
    // zero-copy: kernel-free check of mapped pinned read bandwidth via cudaMemcpy D2D from mapped pointer
    {
        void *dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, pin, 0) == cudaSuccess) {
            double t0 = now(); cudaMemcpyAsync(d, dp, N, cudaMemcpyDefault, st); cudaStreamSynchronize(st); double t1 = now();
            printf("mapped pinned -> device via device pointer: %.1f GB/s\n", N / (t1 - t0) / 1e9);
        }
    }
    return 0;
}
