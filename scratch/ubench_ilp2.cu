// Would two members per team, interleaved in one instruction stream (ILP 2 x TLP 2), beat four independent
// teams per SM (TLP 4)?  Synthetic K3 step (full chains, K-table carry), U members per thread.  Scratch tool.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../fiberis_b200/csrc/fbr_sweep.cuh"
using namespace fbr;

template <int W, int U, int MINB>
__global__ void __launch_bounds__(32 * W, MINB) k3u_kernel(const double *in, double *out, int n_steps) {
    __shared__ __align__(16) double s_K[U][2][W][W];
    __shared__ __align__(16) double s_E[U][2][W];
    __shared__ double s_dummy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double l[U][8], r[U][8], g[U][8], x[U][8];
    ScanMultipliers sm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double *p = in + ((size_t)((blockIdx.x * U + u) % 64) * 32 * W + tid) * 32;
        for (int j = 0; j < 8; ++j) { l[u][j] = p[j]; r[u][j] = p[8 + j]; g[u][j] = p[16 + j]; x[u][j] = p[24 + j]; }
        sm[u] = scan_multipliers<double>(chunk_product<double, 8>(l[u]), chunk_product<double, 8>(g[u]), lane);
        if (tid < W) for (int v = 0; v < W; ++v) { s_K[u][0][tid][v] = v < tid ? 0.01 : 0.0; s_K[u][1][tid][v] = v > tid ? 0.01 : 0.0; }
    }
    __syncthreads();
    unsigned st_f[U], st_b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        st_f[u] = smem_addr(lane == 31 ? &s_E[u][0][warp] : &s_dummy);
        st_b[u] = smem_addr(lane == 0 ? &s_E[u][1][warp] : &s_dummy);
    }
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;
#pragma unroll 1
    for (int n = 0; n < n_steps; ++n) {
        {
            double y[U], c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) y[u] = chunk_end_fwd_plain<double, 8>(x[u], l[u]);
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int u = 0; u < U; ++u) y[u] = fma(sm[u].f[s], __shfl_up_sync(kAllLanes, y[u], 1 << s), y[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) { c[u] = and_mask(__shfl_up_sync(kAllLanes, y[u], 1), keep_f); sts_f64(st_f[u], y[u]); }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double cw = 0.0;
#pragma unroll
                for (int v = 0; v < W - 1; ++v) cw = fma(s_K[u][0][warp][v], s_E[u][0][v], cw);
                c[u] = fma(sm[u].ex_f, cw, c[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) chunk_apply_fwd_plain<double, 8>(x[u], l[u], r[u], c[u]);
        }
        {
            double z[U], c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) z[u] = chunk_end_bwd_plain<double, 8>(x[u], g[u]);
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int u = 0; u < U; ++u) z[u] = fma(sm[u].b[s], __shfl_down_sync(kAllLanes, z[u], 1 << s), z[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) { c[u] = and_mask(__shfl_down_sync(kAllLanes, z[u], 1), keep_b); sts_f64(st_b[u], z[u]); }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double cw = 0.0;
#pragma unroll
                for (int v = 1; v < W; ++v) cw = fma(s_K[u][1][warp][v], s_E[u][1][v], cw);
                c[u] = fma(sm[u].ex_b, cw, c[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) chunk_apply_bwd_plain<double, 8>(x[u], g[u], c[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        double *q = out + ((size_t)(blockIdx.x * U + u) * 32 * W + tid) * 8;
        for (int j = 0; j < 8; ++j) q[j] = x[u][j];
    }
}

template <int W, int U, int MINB>
void run(const char *name, const double *in, double *out, int ctas, int n_steps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k3u_kernel<W, U, MINB><<<ctas, 32 * W>>>(in, out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k3u_kernel<W, U, MINB><<<ctas, 32 * W>>>(in, out, n_steps);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3u_kernel<W, U, MINB>, 32 * W, 0);
    printf("%-40s W=%d U=%d ctas=%4d occ=%d  %7.3f us/step  -> C3 (4096 members x 5000 steps) %.1f ms  err=%s\n", name, W, U, ctas, occ,
           ms * 1e3 / n_steps, ms / n_steps * 5000.0 * 4096.0 / (ctas * U), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t)64 * 512 * 32;
    double *in, *out; cudaMalloc(&in, n * 8); cudaMalloc(&out, (size_t)1184 * 512 * 8 * 8);
    double *h = (double *)malloc(n * 8);
    for (size_t i = 0; i < n; ++i) h[i] = 0.1 + 0.3 * ((i * 2654435761u) % 1000) / 1000.0;
    cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice);
    const int S = 2000;
    run<4, 1, 4>("one member per team, 4 teams/SM", in, out, 592, S);
    run<4, 2, 2>("two members per team, 2 teams/SM", in, out, 296, S);
    run<4, 2, 3>("two members per team, 3 teams/SM (spills)", in, out, 444, S);
    run<2, 2, 4>("W=2 (512 cells), two members, 4 teams/SM", in, out, 592, S);
    run<2, 1, 8>("W=2 (512 cells), one member, 8 teams/SM", in, out, 1184, S);
    return 0;
}
