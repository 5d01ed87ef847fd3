// Synthetic time-loop experiments for the near-mode sweep step (scratch tool).
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../fiberis_b200/csrc/fbr_sweep.cuh"
using namespace fbr;

// MODE 0: CTA barrier (current), 1: no sync at all (timing only), 2: pairwise flagged messages in smem
template <int W, int NS, int MODE>
__global__ void __launch_bounds__(32 * W, W >= 16 ? 1 : (W >= 8 ? 2 : 4)) step_kernel(const double *in, double *out, int n_steps) {
    __shared__ __align__(16) double s_Ef[W + 1];
    __shared__ __align__(16) double s_Eb[W + 1];
    __shared__ double s_dummy;
    __shared__ __align__(16) unsigned long long s_msg[2][W + 1][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double l[8], r[8], g[8], x[8];
    const double *p = in + ((size_t)blockIdx.x * 32 * W + tid) * 32;
    for (int j = 0; j < 8; ++j) { l[j] = p[j]; r[j] = p[8 + j]; g[j] = p[16 + j]; x[j] = p[24 + j]; }
    const HalfProducts hf = half_products_fwd(l), hb = half_products_bwd(g);
    const ScanMultipliers sm = scan_multipliers(hf.a * hf.b, hb.a * hb.b, lane);
    double Pf[NS], Pb[NS];
    for (int s = 0; s < NS; ++s) { Pf[s] = sm.f[s]; Pb[s] = sm.b[s]; }
    if (tid == 0) { s_Ef[0] = 0.0; s_Eb[W] = 0.0; }
    if (tid < 2 * (W + 1) * 2) ((unsigned long long *)s_msg)[tid] = 0ull;
    __syncthreads();
    const unsigned st_f = smem_addr(lane == 31 ? &s_Ef[warp + 1] : &s_dummy), ld_f = smem_addr(&s_Ef[warp]);
    const unsigned st_b = smem_addr(lane == 0 ? &s_Eb[warp] : &s_dummy), ld_b = smem_addr(&s_Eb[warp + 1]);
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;
    const unsigned mst_f = smem_addr(&s_msg[0][warp + 1][0]), mld_f = smem_addr(&s_msg[0][warp][0]);
    const unsigned mst_b = smem_addr(&s_msg[1][warp][0]), mld_b = smem_addr(&s_msg[1][warp + 1][0]);
    for (unsigned step = 1; step <= (unsigned)n_steps; ++step) {
        {
            double ya;
            double y = chunk_end_fwd(x, l, hf.b, ya);
#pragma unroll
            for (int s = 0; s < NS; ++s) y = fma(Pf[s], __shfl_up_sync(kAllLanes, y, 1 << s), y);
            const double carry = and_mask(__shfl_up_sync(kAllLanes, y, 1), keep_f);
            double c;
            if (MODE == 0) { sts_f64(st_f, y); __syncthreads(); c = lds_f64(ld_f); }
            else if (MODE == 1) { sts_f64(st_f, y); c = lds_f64(ld_f); }
            else {
                if (lane == 31) {
                    const unsigned long long b = (unsigned long long)__double_as_longlong(y);
                    const unsigned long long lo = (b & 0xffffffffull) | ((unsigned long long)step << 32), hi = (b >> 32) | ((unsigned long long)step << 32);
                    asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(mst_f), "l"(lo), "l"(hi) : "memory");
                }
                c = 0.0;
                if (warp > 0) {
                    unsigned long long lo, hi;
                    do {
                        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(mld_f) : "memory");
                    } while ((unsigned)(lo >> 32) != step || (unsigned)(hi >> 32) != step);
                    c = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
                }
            }
            chunk_apply_fwd(x, l, r, hf.a, ya, fma(sm.ex_f, c, carry));
        }
        {
            double za;
            double z = chunk_end_bwd(x, g, hb.b, za);
#pragma unroll
            for (int s = 0; s < NS; ++s) z = fma(Pb[s], __shfl_down_sync(kAllLanes, z, 1 << s), z);
            const double carry = and_mask(__shfl_down_sync(kAllLanes, z, 1), keep_b);
            double c;
            if (MODE == 0) { sts_f64(st_b, z); __syncthreads(); c = lds_f64(ld_b); }
            else if (MODE == 1) { sts_f64(st_b, z); c = lds_f64(ld_b); }
            else {
                if (lane == 0) {
                    const unsigned long long b = (unsigned long long)__double_as_longlong(z);
                    const unsigned long long lo = (b & 0xffffffffull) | ((unsigned long long)step << 32), hi = (b >> 32) | ((unsigned long long)step << 32);
                    asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(mst_b), "l"(lo), "l"(hi) : "memory");
                }
                c = 0.0;
                if (warp < W - 1) {
                    unsigned long long lo, hi;
                    do {
                        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(mld_b) : "memory");
                    } while ((unsigned)(lo >> 32) != step || (unsigned)(hi >> 32) != step);
                    c = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
                }
            }
            chunk_apply_bwd(x, g, hb.a, za, fma(sm.ex_b, c, carry));
        }
    }
    double *q = out + ((size_t)blockIdx.x * 32 * W + tid) * 8;
    for (int j = 0; j < 8; ++j) q[j] = x[j];
}

template <int W, int NS, int MODE>
void run(const char *name, const double *in, double *out, int ctas, int n_steps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    step_kernel<W, NS, MODE><<<ctas, 32 * W>>>(in, out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    step_kernel<W, NS, MODE><<<ctas, 32 * W>>>(in, out, n_steps);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cells = (double)ctas * 32 * W * 8;
    printf("%-34s W=%2d NS=%d ctas=%4d  %7.3f us/step  %7.1f Gcu/s  err=%s\n", name, W, NS, ctas, ms * 1e3 / n_steps,
           cells * n_steps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int W, int NS>
void check(const double *in, double *out, int ctas, size_t n) {
    double *o2; cudaMalloc(&o2, n * 8 / 4);
    step_kernel<W, NS, 0><<<ctas, 32 * W>>>(in, out, 50);
    step_kernel<W, NS, 2><<<ctas, 32 * W>>>(in, o2, 50);
    cudaDeviceSynchronize();
    size_t cnt = (size_t)ctas * 32 * W * 8;
    double *a = (double *)malloc(cnt * 8), *b = (double *)malloc(cnt * 8);
    cudaMemcpy(a, out, cnt * 8, cudaMemcpyDeviceToHost); cudaMemcpy(b, o2, cnt * 8, cudaMemcpyDeviceToHost);
    double mx = 0, ref = 0; size_t first = cnt;
    for (size_t i = 0; i < cnt; ++i) { double d = fabs(a[i] - b[i]); if (d > mx) mx = d; if (fabs(a[i]) > ref) ref = fabs(a[i]); if (d > 1e-12 && first == cnt) first = i; }
    printf("check W=%d NS=%d: max |barrier - flags| = %.3e (ref %.3e) first bad %zu\n", W, NS, mx, ref, first);
}
int main() {
    const size_t n = (size_t)148 * 4 * 512 * 32;
    double *in, *out; cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8);
    double *h = (double *)malloc(n * 8);
    for (size_t i = 0; i < n; ++i) h[i] = 0.1 + 0.3 * ((i * 2654435761u) % 1000) / 1000.0;
    cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice);
    check<16, 3>(in, out, 148, n); check<8, 5>(in, out, 4, n); check<4, 5>(in, out, 4, n);
    const int S = 2000;
    run<16, 3, 0>("barrier", in, out, 148, S);
    run<16, 3, 1>("no sync (timing only)", in, out, 148, S);
    run<16, 3, 2>("pairwise flags", in, out, 148, S);
    run<16, 5, 0>("barrier", in, out, 148, S);
    run<16, 5, 2>("pairwise flags", in, out, 148, S);
    run<8, 3, 0>("barrier", in, out, 296, S);
    run<8, 3, 1>("no sync (timing only)", in, out, 296, S);
    run<8, 3, 2>("pairwise flags", in, out, 296, S);
    run<4, 3, 0>("barrier", in, out, 592, S);
    run<4, 3, 2>("pairwise flags", in, out, 592, S);
    run<4, 5, 0>("barrier", in, out, 592, S);
    run<4, 5, 2>("pairwise flags", in, out, 592, S);
    run<16, 3, 0>("barrier, 1 CTA only", in, out, 1, S);
    run<16, 3, 2>("pairwise, 1 CTA only", in, out, 1, S);
    return 0;
}
