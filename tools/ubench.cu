// Scratch micro-benchmarks: fp64 FMA throughput and streaming-store bandwidth on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double *out, int iters)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void fill(double *p, long n)
{
    long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p + i), "d"(1.0) : "memory");
}
int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s SMs=%d clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    double *out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        dfma<<<148, warps * 32>>>(out, 100);
        cudaEventRecord(e0);
        dfma<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * 148.0 * warps * 32;
        printf("dfma warps/SM=%d: %.2f TFLOP/s (%.1f FMA/clk/SM at 1.9 GHz)\n", warps, fl / ms / 1e9,
               fl / 2 / (ms * 1e-3) / 148 / 1.9e9);
    }
    long n = 1L << 30; double *big; cudaMalloc(&big, n * 8);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        fill<<<148 * 16, 256>>>(big, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("streaming store 8 GiB: %.1f GB/s\n", n * 8.0 / ms / 1e6);
    }
    return 0;
}
