// Dependent-issue latencies on B200 that bound one time step of the sweep kernels (single warp, clock64):
//   DFMA chain, FFMA chain, SHFL.UP of a double (2 x SHFL.32) + DFMA, STS + BAR.SYNC + LDS round trip (4 / 8 / 16 warps)
// Scratch tool: nvcc -O3 -gencode arch=compute_100a,code=sm_100a scratch/ubench_lat.cu -o scratch/ubench_lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, long long *cyc, double m) {
    double y = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) y = fma(m, y, 1.0);
    long long t1 = clock64();
    out[threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float *out, long long *cyc, float m) {
    float y = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; ++i) y = fmaf(m, y, 1.0f);
    long long t1 = clock64();
    out[threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc, double m) {
    double y = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) y = fma(m, __shfl_up_sync(0xffffffffu, y, 1), y);
    long long t1 = clock64();
    out[threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double *out, long long *cyc) {
    __shared__ double s[64];
    double y = out[threadIdx.x];
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 1024; ++i) {
        if ((threadIdx.x & 31) == 31) s[(i & 1) * 32 + w] = y;
        __syncthreads();
        y += s[(i & 1) * 32 + (w + 1) % nw];
    }
    long long t1 = clock64();
    out[threadIdx.x] = y;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *d; float *f; long long *c, h;
    cudaMalloc(&d, 8192); cudaMalloc(&f, 8192); cudaMalloc(&c, 64);
    cudaMemset(d, 0, 8192); cudaMemset(f, 0, 8192);
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma<<<1, 32>>>(d, c, 0.5); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA dependent chain:             %.1f cycles per op\n", h / 4096.0);
        k_ffma<<<1, 32>>>(f, c, 0.5f); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("FFMA dependent chain:             %.1f cycles per op\n", h / 4096.0);
        k_shfl<<<1, 32>>>(d, c, 0.5); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("SHFL.UP(double) + DFMA stage:     %.1f cycles per stage\n", h / 1024.0);
        for (int nw : {4, 8, 16}) {
            k_bar<<<1, 32 * nw>>>(d, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("STS + BAR.SYNC + LDS + DADD, %2d warps: %.1f cycles per round\n", nw, h / 1024.0);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
