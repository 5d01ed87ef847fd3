// Micro-benchmarks for the fp64 sweep building blocks on B200 (scratch tool, not part of the library).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void dfma_lat(double *out, long long *cyc, double m) {
    double y = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) y = fma(m, y, 1.0);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void dfma_tp(double *out, long long *cyc, double m) {  // 8 independent chains per thread
    double y[8];
    for (int k = 0; k < 8; ++k) y[k] = threadIdx.x + k;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = fma(m, y[k], 1.0);
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_lat(double *out, long long *cyc) {
    double y = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) y = __shfl_up_sync(0xffffffffu, y, 1);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void scan_stage_lat(double *out, long long *cyc, double m) {  // shfl + dfma dependent
    double y = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) y = fma(m, __shfl_up_sync(0xffffffffu, y, 1), y);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void bar_lat(double *out, long long *cyc) {
    __shared__ double s[64];
    double y = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
        if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = y;
        __syncthreads();
        y += s[(threadIdx.x >> 5) ^ 1];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 8 * 4096);
    long long h[4096];
    auto rep = [&](const char *name, int blocks, int threads, double per) {
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-28s blocks=%4d threads=%4d  %8.2f cyc/iter (max over blocks) err=%s\n", name, blocks, threads, mx / per, cudaGetErrorString(cudaGetLastError()));
    };
    for (int th : {32, 128, 256, 512, 1024}) { dfma_lat<<<1, th>>>(out, cyc, 0.5); rep("dfma dependent chain", 1, th, N); }
    for (int th : {32, 128, 256, 512, 1024}) { dfma_tp<<<1, th>>>(out, cyc, 0.5); rep("dfma 8 chains (per 8 dfma)", 1, th, N); }
    for (int th : {32, 128, 512}) { shfl_lat<<<1, th>>>(out, cyc); rep("shfl(double) dependent", 1, th, N); }
    for (int th : {32, 128, 512}) { scan_stage_lat<<<1, th>>>(out, cyc, 0.5); rep("shfl+dfma stage", 1, th, N); }
    for (int th : {64, 128, 256, 512}) { bar_lat<<<1, th>>>(out, cyc); rep("sts+bar+lds+dadd", 1, th, N); }
    return 0;
}
