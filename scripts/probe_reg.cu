// probe_reg.cu — can cudaHostRegister be parallelised over threads? (decides zero-copy staging)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    const size_t N = 1ull << 30;
    cudaFree(0);
    char *pg = (char *)aligned_alloc(2 << 20, N);
    memset(pg, 3, N);
    for (size_t P : {size_t(16) << 20, size_t(64) << 20}) {
        for (int nt : {1, 4, 8, 16}) {
            const size_t np = N / P;
            double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++)
                th.emplace_back([=] { for (size_t i = t; i < np; i += nt) cudaHostRegister(pg + i * P, P, cudaHostRegisterDefault); });
            for (auto &x : th) x.join();
            double t1 = now();
            th.clear();
            for (int t = 0; t < nt; t++)
                th.emplace_back([=] { for (size_t i = t; i < np; i += nt) cudaHostUnregister(pg + i * P); });
            for (auto &x : th) x.join();
            double t2 = now();
            printf("pieces of %zu MiB, %2d threads: register %.1f ms (%.1f GB/s), unregister %.1f ms (%.1f GB/s)\n", P >> 20, nt,
                   (t1 - t0) * 1e3, N / (t1 - t0) / 1e9, (t2 - t1) * 1e3, N / (t2 - t1) / 1e9);
        }
    }
    return 0;
}
