// Launch-rate probe: T host threads, each launching an empty kernel into its own stream for ~1 s.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
__global__ void k_empty(int* p) { if (p && threadIdx.x == 999) *p = 1; }
int main(int argc, char** argv) {
    for (int T : {1, 4, 8, 16, 32}) {
        std::atomic<long long> total{0};
        std::vector<std::thread> th;
        auto t0 = std::chrono::steady_clock::now();
        for (int t = 0; t < T; t++) th.emplace_back([&] {
            cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
            long long n = 0;
            while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 1.0) {
                for (int i = 0; i < 20; i++) k_empty<<<1, 32, 0, s>>>(nullptr);
                cudaStreamSynchronize(s);
                n += 20;
            }
            total += n; cudaStreamDestroy(s);
        });
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %2d: %.0f launches/s\n", T, total / dt);
    }
}
