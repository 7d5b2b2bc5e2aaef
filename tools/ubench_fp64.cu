// micro-benchmark: FP64 FMA vs DMMA (mma.sync m8n8k4 f64) issue rates on this GPU (dev tool)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
    double a[8]; double x = threadIdx.x * 1e-9, y = 1.0000001;
    for (int i = 0; i < 8; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
    double c[4][2]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000; float ms;
        k_dfma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 148 * warps * 32 * 8.0 * iters;
        printf("DFMA  warps/SM=%2d : %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        k_dmma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 148 * warps * 4.0 * 256 * iters;
        printf("DMMA  warps/SM=%2d : %.2f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    return 0;
}
