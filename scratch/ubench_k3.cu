// Which ingredient of the ensemble time loop costs what?  Synthetic experiments (scratch tool).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../fiberis_b200/csrc/fbr_sweep.cuh"
using namespace fbr;

// FLAGS bit0: exact K-table carry (else neighbour only)   bit1: override by FSEL chain   bit2: override by switch
//       bit3: observation by select tree + prefetch       bit4: observation by switch    bit5: old full chains (no half split)
template <int W, int FLAGS>
__global__ void __launch_bounds__(32 * W, 16 / W) k3_kernel(const double *in, double *out, const double *tab, int n_steps, double *mis) {
    __shared__ __align__(16) double s_K[2][W][W];
    __shared__ __align__(16) double s_E[2][W + 1];
    __shared__ double s_dummy;
    __shared__ double s_obs[32 * 16];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_obs[i] = 0.25;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double l[8], r[8], g[8], x[8];
    const double *p = in + ((size_t)(blockIdx.x % 64) * 32 * W + tid) * 32;
    for (int j = 0; j < 8; ++j) { l[j] = p[j]; r[j] = p[8 + j]; g[j] = p[16 + j]; x[j] = p[24 + j]; }
    const HalfProducts hf = half_products_fwd(l), hb = half_products_bwd(g);
    const ScanMultipliers sm = scan_multipliers(hf.a * hf.b, hb.a * hb.b, lane);
    if (tid < W) for (int v = 0; v < W; ++v) { s_K[0][tid][v] = v < tid ? 0.01 : 0.0; s_K[1][tid][v] = v > tid ? 0.01 : 0.0; }
    if (tid == 0) { s_E[0][0] = 0; s_E[1][W] = 0; }
    __syncthreads();
    const unsigned st_f = smem_addr(lane == 31 ? &s_E[0][warp + ((FLAGS & 1) ? 0 : 1)] : &s_dummy), ld_f = smem_addr(&s_E[0][warp]);
    const unsigned st_b = smem_addr(lane == 0 ? &s_E[1][warp] : &s_dummy), ld_b = smem_addr(&s_E[1][warp + 1]);
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;
    const int ov_j = (tid % 43 == 0) ? (tid % 8) : -1;   // ~3 threads of 128 own an override
    const int ob_j = (tid % 8 == 3) ? ((tid / 8) % 8) : -1;  // every 8th thread observes
    const double *ovp = tab + (tid % 4), *obp = tab + 1024 + (tid % 16);
    double ov_val = 0.5, ob_val = 0.25, acc = 0.0;
    for (int n = 0; n < n_steps; ++n) {
        if (FLAGS & 2) {
            if (ov_j >= 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = (k == ov_j) ? ov_val : x[k];
            }
            ovp += 4; ov_val = __ldg(ovp);
        }
        if (FLAGS & 4) {
            if (ov_j >= 0) {
                switch (ov_j) { case 0: x[0] = ov_val; break; case 1: x[1] = ov_val; break; case 2: x[2] = ov_val; break; case 3: x[3] = ov_val; break;
                                case 4: x[4] = ov_val; break; case 5: x[5] = ov_val; break; case 6: x[6] = ov_val; break; default: x[7] = ov_val; break; }
                ovp += 4; ov_val = __ldg(ovp);
            }
        }
        {
            double ya, y;
            if (FLAGS & 32) { y = 0.0; for (int j = 0; j < 8; ++j) y = fma(-l[j], y, x[j]); ya = 0; }
            else y = chunk_end_fwd(x, l, hf.b, ya);
#pragma unroll
            for (int s = 0; s < 5; ++s) y = fma(sm.f[s], __shfl_up_sync(kAllLanes, y, 1 << s), y);
            double c = and_mask(__shfl_up_sync(kAllLanes, y, 1), keep_f);
            sts_f64(st_f, y);
            __syncthreads();
            double cw = 0.0;
            if (FLAGS & 1) {
#pragma unroll
                for (int v = 0; v < W - 1; ++v) cw = fma(s_K[0][warp][v], s_E[0][v], cw);
            } else cw = lds_f64(ld_f);
            c = fma(sm.ex_f, cw, c);
            if (FLAGS & 32) { y = c; for (int j = 0; j < 8; ++j) { y = fma(-l[j], y, x[j]); x[j] = r[j] * y; } }
            else chunk_apply_fwd(x, l, r, hf.a, ya, c);
        }
        {
            double za, z;
            if (FLAGS & 32) { z = 0.0; for (int j = 7; j >= 0; --j) z = fma(-g[j], z, x[j]); za = 0; }
            else z = chunk_end_bwd(x, g, hb.b, za);
#pragma unroll
            for (int s = 0; s < 5; ++s) z = fma(sm.b[s], __shfl_down_sync(kAllLanes, z, 1 << s), z);
            double c = and_mask(__shfl_down_sync(kAllLanes, z, 1), keep_b);
            sts_f64(st_b, z);
            __syncthreads();
            double cw = 0.0;
            if (FLAGS & 1) {
#pragma unroll
                for (int v = 1; v < W; ++v) cw = fma(s_K[1][warp][v], s_E[1][v], cw);
            } else cw = lds_f64(ld_b);
            c = fma(sm.ex_b, cw, c);
            if (FLAGS & 32) { z = c; for (int j = 7; j >= 0; --j) { z = fma(-g[j], z, x[j]); x[j] = z; } }
            else chunk_apply_bwd(x, g, hb.a, za, c);
        }
        if (FLAGS & 8) {
            if (ob_j >= 0) {
                const bool b0 = ob_j & 1, b1 = ob_j & 2, b2 = ob_j & 4;
                const double a0 = b0 ? x[1] : x[0], a1 = b0 ? x[3] : x[2], a2 = b0 ? x[5] : x[4], a3 = b0 ? x[7] : x[6];
                const double c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
                const double res = (b2 ? c1 : c0) - ob_val;
                acc = fma(res, res, acc);
            }
            obp += 16; ob_val = __ldg(obp);
        }
        if (FLAGS & 64) {   // tree executed by everybody, masked accumulate
            const bool b0 = ob_j & 1, b1 = ob_j & 2, b2 = ob_j & 4;
            const double a0 = b0 ? x[1] : x[0], a1 = b0 ? x[3] : x[2], a2 = b0 ? x[5] : x[4], a3 = b0 ? x[7] : x[6];
            const double c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
            const double res = (b2 ? c1 : c0) - ob_val;
            acc = fma(res, res, acc);
            if (!(FLAGS & 128)) { obp += 16; ob_val = __ldg(obp); }
        }
        if (FLAGS & 256) { obp += 16; ob_val = __ldg(obp); acc += ob_val; }
        if (FLAGS & 512) {  // branch, tree, no load
            if (ob_j >= 0) {
                const bool b0 = ob_j & 1, b1 = ob_j & 2, b2 = ob_j & 4;
                const double a0 = b0 ? x[1] : x[0], a1 = b0 ? x[3] : x[2], a2 = b0 ? x[5] : x[4], a3 = b0 ? x[7] : x[6];
                const double c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
                const double res = (b2 ? c1 : c0) - ob_val;
                acc = fma(res, res, acc);
            }
        }
        if (FLAGS & 1024) {  // observe through shared memory: owner stores x[ob_j]... everybody stores cell (tid & 7)?  here: fixed cell 3
            const double res = x[3] - ob_val;
            acc = fma(res, res, acc);
            obp += 16; ob_val = __ldg(obp);
        }
        if (FLAGS & 2048) {  // tree by everybody, value from shared memory (staged block of rows)
            const bool b0 = ob_j & 1, b1 = ob_j & 2, b2 = ob_j & 4;
            const double a0 = b0 ? x[1] : x[0], a1 = b0 ? x[3] : x[2], a2 = b0 ? x[5] : x[4], a3 = b0 ? x[7] : x[6];
            const double c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
            const double res = (b2 ? c1 : c0) - s_obs[(n & 31) * 16 + (tid & 15)];
            acc = fma(res, res, acc);
        }
        if (FLAGS & 4096) {  // fixed cell, immediate consumption of a global load
            const double res = x[3] - __ldg(obp + 16 * n);
            acc = fma(res, res, acc);
        }
        if (FLAGS & 8192) {  // branch + tree, value from shared memory
            if (ob_j >= 0) {
                const bool b0 = ob_j & 1, b1 = ob_j & 2, b2 = ob_j & 4;
                const double a0 = b0 ? x[1] : x[0], a1 = b0 ? x[3] : x[2], a2 = b0 ? x[5] : x[4], a3 = b0 ? x[7] : x[6];
                const double c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2;
                const double res = (b2 ? c1 : c0) - s_obs[(n & 31) * 16 + (tid & 15)];
                acc = fma(res, res, acc);
            }
        }
        if (FLAGS & 16) {
            if (ob_j >= 0) {
                double v;
                switch (ob_j) { case 0: v = x[0]; break; case 1: v = x[1]; break; case 2: v = x[2]; break; case 3: v = x[3]; break;
                                case 4: v = x[4]; break; case 5: v = x[5]; break; case 6: v = x[6]; break; default: v = x[7]; break; }
                const double res = v - ob_val;
                acc = fma(res, res, acc);
                obp += 16; ob_val = __ldg(obp);
            }
        }
    }
    double *q = out + ((size_t)blockIdx.x * 32 * W + tid) * 8;
    for (int j = 0; j < 8; ++j) q[j] = x[j];
    if (acc != 0.123) mis[blockIdx.x * 32 * W + tid] = acc;
}

template <int W, int FLAGS>
void run(const char *name, const double *in, double *out, const double *tab, double *mis, int ctas, int n_steps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k3_kernel<W, FLAGS><<<ctas, 32 * W>>>(in, out, tab, 10, mis);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k3_kernel<W, FLAGS><<<ctas, 32 * W>>>(in, out, tab, n_steps, mis);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-52s W=%2d ctas=%4d  %7.3f us/step -> C3 %.1f ms  err=%s\n", name, W, ctas, ms * 1e3 / n_steps,
           ms / n_steps * 5000.0 * 4096.0 / ctas, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = (size_t)64 * 512 * 32;
    double *in, *out, *tab, *mis; cudaMalloc(&in, n * 8); cudaMalloc(&out, (size_t)592 * 512 * 8 * 8); cudaMalloc(&tab, 64 << 20); cudaMalloc(&mis, 592 * 512 * 8);
    double *h = (double *)malloc(n * 8);
    for (size_t i = 0; i < n; ++i) h[i] = 0.1 + 0.3 * ((i * 2654435761u) % 1000) / 1000.0;
    cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice); cudaMemset(tab, 0, 64 << 20);
    const int S = 2000;
    run<4, 1 | 2048>("K, obs tree all threads, value from smem", in, out, tab, mis, 592, S);
    run<4, 1 | 8192>("K, branch + tree, value from smem", in, out, tab, mis, 592, S);
    run<4, 1 | 4096>("K, fixed cell, immediate ldg", in, out, tab, mis, 592, S);
    run<4, 1 | 64>("K, obs tree all threads + ldg", in, out, tab, mis, 592, S);
    run<4, 1 | 64 | 128>("K, obs tree all threads, no ldg", in, out, tab, mis, 592, S);
    run<4, 1 | 256>("K, ldg only", in, out, tab, mis, 592, S);
    run<4, 1 | 512>("K, branch + tree, no ldg", in, out, tab, mis, 592, S);
    run<4, 1 | 1024>("K, fixed cell obs + ldg", in, out, tab, mis, 592, S);
    run<4, 0>("half chains, neighbour carry", in, out, tab, mis, 592, S);
    run<4, 32>("full chains, neighbour carry", in, out, tab, mis, 592, S);
    run<4, 1>("half chains, K-table carry", in, out, tab, mis, 592, S);
    run<4, 33>("full chains, K-table carry", in, out, tab, mis, 592, S);
    run<4, 1 | 2>("half, K, override FSEL", in, out, tab, mis, 592, S);
    run<4, 1 | 4>("half, K, override switch", in, out, tab, mis, 592, S);
    run<4, 1 | 8>("half, K, obs select tree", in, out, tab, mis, 592, S);
    run<4, 1 | 16>("half, K, obs switch", in, out, tab, mis, 592, S);
    run<4, 1 | 2 | 8>("half, K, override FSEL + obs tree", in, out, tab, mis, 592, S);
    run<4, 1 | 4 | 16>("half, K, override switch + obs switch", in, out, tab, mis, 592, S);
    run<4, 33 | 2 | 8>("full, K, override FSEL + obs tree (old K3)", in, out, tab, mis, 592, S);
    return 0;
}
