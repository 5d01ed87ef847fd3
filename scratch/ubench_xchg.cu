// Exchange of a double with lane-2^s: two SHFL vs STS.64 + LDS.64 through a per-warp patch (scratch tool).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int MODE>
__global__ void xchg(double *out, long long *cyc, double m) {
    __shared__ double buf[2][32 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double y0 = threadIdx.x, y1 = threadIdx.x + 1.0, y2 = y0 + 2, y3 = y0 + 3;
    double *mine0 = &buf[0][warp * 32 + lane], *mine1 = &buf[1][warp * 32 + lane];
    const int src = (lane + 31) & 31;
    double *nb0 = &buf[0][warp * 32 + src], *nb1 = &buf[1][warp * 32 + src];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) {  // dependent chain of scan stages via shuffle
            y0 = fma(m, __shfl_up_sync(0xffffffffu, y0, 1), y0);
        } else if (MODE == 1) {  // dependent chain via shared memory
            *mine0 = y0; __syncwarp(); y0 = fma(m, *nb0, y0); __syncwarp();
        } else if (MODE == 2) {  // 4 independent chains via shuffle (throughput)
            y0 = fma(m, __shfl_up_sync(0xffffffffu, y0, 1), y0); y1 = fma(m, __shfl_up_sync(0xffffffffu, y1, 2), y1);
            y2 = fma(m, __shfl_up_sync(0xffffffffu, y2, 4), y2); y3 = fma(m, __shfl_up_sync(0xffffffffu, y3, 8), y3);
        } else {  // 4 independent exchanges via shared memory
            mine0[0] = y0; mine1[0] = y1; __syncwarp();
            double a = *nb0, b = *nb1; __syncwarp();
            mine0[0] = y2; mine1[0] = y3; __syncwarp();
            double c = *nb0, d = *nb1; __syncwarp();
            y0 = fma(m, a, y0); y1 = fma(m, b, y1); y2 = fma(m, c, y2); y3 = fma(m, d, y3);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = y0 + y1 + y2 + y3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 8 * 64);
    long long h;
    auto rep = [&](const char *name, int th, double per) {
        cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s threads=%4d  %7.2f cyc per exchange+fma  err=%s\n", name, th, h / per, cudaGetErrorString(cudaGetLastError()));
    };
    for (int th : {32, 128, 512, 1024}) { xchg<0><<<1, th>>>(out, cyc, 0.5); rep("shfl chain (latency)", th, N); }
    for (int th : {32, 128, 512, 1024}) { xchg<1><<<1, th>>>(out, cyc, 0.5); rep("sts+lds chain (latency)", th, N); }
    for (int th : {128, 512, 1024}) { xchg<2><<<1, th>>>(out, cyc, 0.5); rep("shfl x4 independent (per exchange)", th, 4.0 * N); }
    for (int th : {128, 512, 1024}) { xchg<3><<<1, th>>>(out, cyc, 0.5); rep("sts+lds x4 independent (per exchange)", th, 4.0 * N); }
    return 0;
}
