// cuinit_probe — where a one-shot CUDA process spends its start-up time on this box (scripts/gpu_cli_timing.sh).
// Usage: cuinit_probe [device_count_to_touch]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k_nop() {}
int main(int argc, char **argv)
{
    const int touch = argc > 1 ? atoi(argv[1]) : 1;
    double t0 = now();
    int n = 0;
    cudaGetDeviceCount(&n);
    double t1 = now();
    printf("cudaGetDeviceCount -> %d: %.1f ms\n", n, t1 - t0);
    for (int d = 0; d < touch && d < n; d++)
    {
        double a = now();
        cudaSetDevice(d);
        cudaFree(0);
        double b = now();
        k_nop<<<1, 1>>>();
        cudaDeviceSynchronize();
        double c = now();
        void *p = nullptr, *h = nullptr;
        cudaMalloc(&p, 768ull << 20);
        double e = now();
        cudaMallocHost(&h, 96ull << 20);
        double f = now();
        printf("device %d: context %.1f ms, first kernel %.1f ms, cudaMalloc 768 MiB %.1f ms, cudaMallocHost 96 MiB %.1f ms\n", d, b - a,
               c - b, e - c, f - e);
    }
    double t2 = now();
    printf("total %.1f ms\n", t2 - t0);
    return 0;
}
