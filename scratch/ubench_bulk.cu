// Round-2 experiment ("no TMA anywhere"): staging a window / tile of four fp64 rows in shared memory —
//   A  16-byte cp.async (LDGSTS) by every thread into a layout that is conflict-free for the 64-bytes-per-thread reads
//      (what K3W / K3P do: piece j of a thread's 64 bytes at [(j * 256 + tid) * 16]),
//   B  cp.async.bulk (1-D TMA, UBLKCP) + mbarrier: ONE thread issues four 16 KB copies; the rows land linearly, so a
//      thread's 64 bytes are read with 16-byte loads 64 bytes apart (4-way bank conflicts per quarter warp).
// Both double-buffered over a long stream of tiles (2048 cells x 4 rows = 64 KB per tile); reports achieved GB/s.
// Scratch tool: ./scratch/ubench_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kT = 256, kCells = 2048;

__device__ __forceinline__ unsigned saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kT, 2) stage_ldgsts(const double *__restrict__ src, double *out, int64_t tiles_per_cta) {
    extern __shared__ __align__(128) double s[];  // 2 x 4 x 2048 doubles
    const int tid = threadIdx.x;
    const int64_t first = blockIdx.x * tiles_per_cta;
    auto issue = [&](int64_t t, int buf) {
        const double *g = src + ((first + t) % (gridDim.x * tiles_per_cta)) * 4 * kCells + (int64_t)tid * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned d = saddr(s) + buf * 65536 + 16u * (kT * (q * 4 + j) + tid);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + q * kCells + 2 * j) : "memory");
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc = 0.0;
    issue(0, 0);
    for (int64_t t = 0; t < tiles_per_cta; ++t) {
        if (t + 1 < tiles_per_cta) issue(t + 1, (t + 1) & 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // my own copies of tile t: no barrier needed
        const int buf = t & 1;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(s) + buf * 65536 + 16 * (kT * (q * 4 + j) + tid));
                acc += v.x * 1.0000001 + v.y;
            }
    }
    out[blockIdx.x * kT + tid] = acc;
}

__global__ void __launch_bounds__(kT, 2) stage_bulk(const double *__restrict__ src, double *out, int64_t tiles_per_cta) {
    extern __shared__ __align__(128) double s[];
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x;
    const int64_t first = blockIdx.x * tiles_per_cta;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int64_t t, int buf) {  // thread 0 only
        const double *g = src + ((first + t) % (gridDim.x * tiles_per_cta)) * 4 * kCells;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar[buf])), "r"(65536u) : "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             saddr(s) + buf * 65536 + q * 16384),
                         "l"(g + q * kCells), "r"(16384u), "r"(saddr(&bar[buf]))
                         : "memory");
    };
    double acc = 0.0;
    if (tid == 0) issue(0, 0);
    for (int64_t t = 0; t < tiles_per_cta; ++t) {
        const int buf = t & 1;
        if (t + 1 < tiles_per_cta) {
            __syncthreads();  // everybody has read buffer (t + 1) & 1 (tile t - 1) before the async proxy overwrites it
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t + 1, (t + 1) & 1);
            }
        }
        const unsigned parity = (unsigned)((t >> 1) & 1);
        unsigned ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(saddr(&bar[buf])), "r"(parity) : "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // linear rows: thread tid owns bytes [tid * 64, tid * 64 + 64) of every row
                const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(s) + buf * 65536 + q * 16384 + tid * 64 + j * 16);
                acc += v.x * 1.0000001 + v.y;
            }
    }
    out[blockIdx.x * kT + tid] = acc;
}

int main() {
    const int ctas = 296;
    const int64_t tiles_per_cta = 64;  // 296 x 64 x 64 KB = 1.2 GB streamed once: far beyond L2
    const size_t words = (size_t)ctas * tiles_per_cta * 4 * kCells;
    double *src, *out;
    cudaMalloc(&src, words * 8);
    cudaMalloc(&out, (size_t)ctas * kT * 8 * 2);
    cudaMemset(src, 0, words * 8);
    cudaFuncSetAttribute(stage_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaFuncSetAttribute(stage_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            if (which == 0) stage_ldgsts<<<ctas, kT, 131072>>>(src, out, tiles_per_cta);
            else stage_bulk<<<ctas, kT, 131072>>>(src, out + ctas * kT, tiles_per_cta);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (rep && ms < best) best = ms;
        }
        printf("%-58s %8.3f ms  %7.1f GB/s  (%s)\n", which == 0 ? "A: 16-byte cp.async, interleaved conflict-free layout" : "B: cp.async.bulk + mbarrier, linear rows (conflicting reads)",
               best, words * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
