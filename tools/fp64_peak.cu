// FP64 FMA and exp() throughput probe for the roofline notes (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp64_peak.cu -o tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fma_kernel(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_kernel(double* out, int iters) {
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = -(threadIdx.x * 1e-2 + i);
    double s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s += exp(a[i]);
            a[i] -= 1e-6;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* d;
    cudaMalloc(&d, (size_t)blocks * threads * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * iters * (double)blocks * threads;
        printf("fp64 fma: %.2f TFLOP/s (%.3f ms)\n", flops / ms * 1e-9, ms);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        exp_kernel<<<blocks, threads>>>(d, iters / 10);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double n = 4.0 * (iters / 10) * (double)blocks * threads;
        printf("fp64 exp: %.2f Gexp/s (%.3f ms)\n", n / ms * 1e-6, ms);
    }
    printf("SMs %d clock %d kHz L2 %d MB\n", p.multiProcessorCount, p.clockRate, p.l2CacheSize >> 20);
    return 0;
}
