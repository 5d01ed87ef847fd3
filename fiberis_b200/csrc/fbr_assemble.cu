// K1 fused coefficient assembly, right-hand side, K5 error norms (K1b lives in fbr_factorize.cu).
// All bandwidth-trivial one-off kernels (they run once per dt, the time loop is in fbr_run_onchip.cu).
#include <stdarg.h>

#include <dlfcn.h>

#include "fbr_common.cuh"

namespace fbr {

std::string &last_error_ref() {
    static thread_local std::string s;
    return s;
}
int &launch_count_ref() {
    static thread_local int n = 0;
    return n;
}

int device_facts(DeviceFacts *out) {
    static thread_local DeviceFacts cached;
    int dev = 0;
    FBR_CUDA(cudaGetDevice(&dev));
    if (cached.device != dev) {
        cudaDeviceProp p;
        FBR_CUDA(cudaGetDeviceProperties(&p, dev));
        cached.device = dev;
        cached.sm_count = p.multiProcessorCount;
        cached.cc_major = p.major;
        cached.cc_minor = p.minor;
        cached.smem_optin = (int64_t)p.sharedMemPerBlockOptin;
        cached.l2_bytes = (int64_t)p.l2CacheSize;
        cached.hbm_bytes = (int64_t)p.totalGlobalMem;
    }
    *out = cached;
    return FBR_OK;
}

// ------------------------------------------------------------------------------------------------
// K1: one thread per (model, cell).  The arithmetic follows the NumPy expression order of
// matbuilder.py:21-28 with round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an
// FMA: the three diagonals are bit-identical to the reference's dense A.
// ------------------------------------------------------------------------------------------------
struct DiffusivityDense {
    const double *D;
    int64_t stride;
    __device__ double at(int64_t m, int64_t i) const { return D[m * stride + i]; }
};
struct DiffusivityZonal {
    const double *zone_value;  // (M, n_zones) = 10 ** theta, expanded by the host
    const int32_t *zone_of_cell;
    int n_zones;
    __device__ double at(int64_t m, int64_t i) const { return zone_value[m * n_zones + zone_of_cell[i]]; }
};

// PML ramp of matbuilder.py:170-196 for one node (pml > 0).
__device__ __forceinline__ double pml_sigma_at(const double *__restrict__ x, int64_t i, int64_t nx, double pml,
                                               double sigma_max) {
    const double dist_left = __dsub_rn(x[i], x[0]), dist_right = __dsub_rn(x[nx - 1], x[i]);
    double sigma = 0.0;
    if (dist_left < pml) sigma = __dmul_rn(sigma_max, __ddiv_rn(__dsub_rn(pml, dist_left), pml));
    if (dist_right < pml) sigma = fmax(sigma, __dmul_rn(sigma_max, __ddiv_rn(__dsub_rn(pml, dist_right), pml)));
    return sigma;
}

__global__ void __launch_bounds__(256) pml_sigma_kernel(const double *__restrict__ mesh, int64_t nx, double pml,
                                                        double sigma_max, double *__restrict__ sigma) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nx; i += (int64_t)gridDim.x * blockDim.x)
        sigma[i] = pml != 0.0 ? pml_sigma_at(mesh, i, nx, pml, sigma_max) : 0.0;
}

template <class Diff>
__global__ void __launch_bounds__(256) assemble_kernel(const double *__restrict__ mesh, int64_t mesh_stride,
                                                       Diff diff, int64_t M, int64_t nx, double dt, int lbc,
                                                       int rbc, const int64_t *__restrict__ src_idx, int n_src,
                                                       double pml, double sigma_max, double *__restrict__ dl,
                                                       double *__restrict__ d, double *__restrict__ du) {
    const int64_t total = M * nx;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / nx, i = e - m * nx;
        const double *x = mesh + m * mesh_stride;
        double a = 0.0, b = 0.0, c = 0.0;
        if (i > 0 && i < nx - 1) {
            const double xm = x[i - 1], x0 = x[i], xp = x[i + 1];
            const double Dm = diff.at(m, i - 1), D0 = diff.at(m, i), Dp = diff.at(m, i + 1);
            const double dxm = __dsub_rn(x0, xm), dxp = __dsub_rn(xp, x0);
            const double dem = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, Dm), D0), __dadd_rn(Dm, D0));
            const double dep = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, D0), Dp), __dadd_rn(D0, Dp));
            const double sum = __dadd_rn(dxm, dxp);
            const double al = __ddiv_rn(__dmul_rn(dem, dt), __ddiv_rn(__dmul_rn(dxm, sum), 2.0));
            const double ar = __ddiv_rn(__dmul_rn(dep, dt), __ddiv_rn(__dmul_rn(dxp, sum), 2.0));
            a = -al;
            b = __dadd_rn(__dadd_rn(1.0, al), ar);
            c = -ar;
        }
        if (i == 0) {  // matbuilder.py:49-57
            if (lbc == FBR_BC_DIRICHLET) b = 1.0;
            else if (lbc == FBR_BC_NEUMANN) { b = -1.0; c = 1.0; }
        }
        if (i == nx - 1) {  // matbuilder.py:60-69
            if (rbc == FBR_BC_DIRICHLET) b = 1.0;
            else if (rbc == FBR_BC_NEUMANN) { b = -1.0; a = 1.0; }
        }
        for (int k = 0; k < n_src; ++k)  // matbuilder.py:73-76 / 156-161
            if (src_idx[k] == i) { a = 0.0; b = 1.0; c = 0.0; }
        if (pml != 0.0)  // matbuilder.py:79-81
            b = __dadd_rn(b, __dmul_rn(dt, pml_sigma_at(x, i, nx, pml, sigma_max)));
        if (i == 0) a = 0.0;
        if (i == nx - 1) c = 0.0;
        dl[e] = a;
        d[e] = b;
        du[e] = c;
    }
}

__global__ void __launch_bounds__(256) rhs_kernel(const double *__restrict__ p, int64_t M, int64_t nx, int lbc,
                                                  int rbc, const int64_t *__restrict__ src_idx, int n_src,
                                                  const double *__restrict__ src_val, double *__restrict__ b) {
    const int64_t total = M * nx;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % nx;
        double v = p[e];
        if (i == 0 && lbc != FBR_BC_NONE) v = 0.0;
        if (i == nx - 1 && rbc != FBR_BC_NONE) v = 0.0;
        for (int k = 0; k < n_src; ++k)
            if (src_idx[k] == i) v = src_val[k];
        b[e] = v;
    }
}

// K5: one CTA per model; fixed-order tree reduction (deterministic).
__global__ void __launch_bounds__(256) error_norms_kernel(const double *__restrict__ u1,
                                                          const double *__restrict__ u2, int64_t nx,
                                                          double *__restrict__ err) {
    __shared__ double s_num[8], s_den[8];
    const int64_t o = blockIdx.x * nx;
    double num = 0.0, den = 0.0;
    for (int64_t i = threadIdx.x; i < nx; i += blockDim.x) {
        const double a = u1[o + i], df = a - u2[o + i];
        num += df * df;
        den += a * a;
    }
    for (int s = 16; s > 0; s >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, s);
        den += __shfl_xor_sync(0xffffffffu, den, s);
    }
    if ((threadIdx.x & 31) == 0) { s_num[threadIdx.x >> 5] = num; s_den[threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double n2 = 0.0, d2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { n2 += s_num[w]; d2 += s_den[w]; }
        err[blockIdx.x] = sqrt(n2) / sqrt(d2);  // 0/0 -> nan, as numpy (tso.py:18)
    }
}

static inline int grid_for(int64_t total, int block, int sm_count) {
    int64_t g = (total + block - 1) / block;
    const int64_t cap = (int64_t)sm_count * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fbr

using namespace fbr;

static int check_common(int64_t M, int64_t nx, int lbc, int rbc, int n_src, const void *src_idx) {
    FBR_REQUIRE(M >= 1 && nx >= 2, "M (%lld) must be >= 1 and nx (%lld) >= 2", (long long)M, (long long)nx);
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "boundary codes must be 0 (None), 1 (Dirichlet) or 2 (Neumann)");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || src_idx), "n_src=%d but src_idx is NULL", n_src);
    return FBR_OK;
}

extern "C" int fbr_assemble_f64(const double *mesh, int64_t mesh_stride, const double *diffusivity,
                                int64_t diff_stride, int64_t M, int64_t nx, double dt, int lbc, int rbc,
                                const int64_t *src_idx, int n_src, double pml_thickness, double sigma_max,
                                double *dl, double *d, double *du, void *stream) {
    if (int rc = check_common(M, nx, lbc, rbc, n_src, src_idx)) return rc;
    FBR_REQUIRE(mesh && diffusivity && dl && d && du, "NULL array argument");
    FBR_REQUIRE((mesh_stride == 0 || mesh_stride >= nx) && (diff_stride == 0 || diff_stride >= nx), "bad stride");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    DiffusivityDense diff{diffusivity, diff_stride};
    assemble_kernel<<<grid_for(M * nx, 256, f.sm_count), 256, 0, as_stream(stream)>>>(
        mesh, mesh_stride, diff, M, nx, dt, lbc, rbc, src_idx, n_src, pml_thickness, sigma_max, dl, d, du);
    return check_launch("fbr_assemble_f64");
}

extern "C" int fbr_assemble_zonal_f64(const double *mesh, const double *zone_value, int n_zones,
                                      const int32_t *zone_of_cell, int64_t M, int64_t nx, double dt, int lbc,
                                      int rbc, const int64_t *src_idx, int n_src, double pml_thickness,
                                      double sigma_max, double *dl, double *d, double *du, void *stream) {
    if (int rc = check_common(M, nx, lbc, rbc, n_src, src_idx)) return rc;
    FBR_REQUIRE(mesh && zone_value && zone_of_cell && dl && d && du, "NULL array argument");
    FBR_REQUIRE(n_zones >= 1, "n_zones must be >= 1");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    DiffusivityZonal diff{zone_value, zone_of_cell, n_zones};
    assemble_kernel<<<grid_for(M * nx, 256, f.sm_count), 256, 0, as_stream(stream)>>>(
        mesh, 0, diff, M, nx, dt, lbc, rbc, src_idx, n_src, pml_thickness, sigma_max, dl, d, du);
    return check_launch("fbr_assemble_zonal_f64");
}

extern "C" int fbr_pml_sigma_f64(const double *mesh, int64_t nx, double pml_thickness, double sigma_max,
                                 double *sigma, void *stream) {
    FBR_REQUIRE(mesh && sigma && nx >= 1, "bad argument");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    pml_sigma_kernel<<<grid_for(nx, 256, f.sm_count), 256, 0, as_stream(stream)>>>(mesh, nx, pml_thickness, sigma_max,
                                                                                 sigma);
    return check_launch("fbr_pml_sigma_f64");
}

extern "C" int fbr_rhs_f64(const double *p, int64_t M, int64_t nx, int lbc, int rbc, const int64_t *src_idx,
                           int n_src, const double *src_val, double *b, void *stream) {
    if (int rc = check_common(M, nx, lbc, rbc, n_src, src_idx)) return rc;
    FBR_REQUIRE(p && b && (n_src == 0 || src_val), "NULL array argument");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    rhs_kernel<<<grid_for(M * nx, 256, f.sm_count), 256, 0, as_stream(stream)>>>(p, M, nx, lbc, rbc, src_idx, n_src,
                                                                              src_val, b);
    return check_launch("fbr_rhs_f64");
}

extern "C" int fbr_error_norms_f64(const double *u1, const double *u2, int64_t M, int64_t nx, double *err,
                                   void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 1 && u1 && u2 && err, "bad argument");
    error_norms_kernel<<<(unsigned)M, 256, 0, as_stream(stream)>>>(u1, u2, nx, err);
    return check_launch("fbr_error_norms_f64");
}

// Misfit of stored rows against observations (the fused form lives in K3; this one serves every other run kernel):
// misfit[r] = sum_{t >= first_row} sum_k w_k (snap[r, t, obs_idx[k]] - obs[t, k])^2, one CTA per model, fixed summation order.
__global__ void __launch_bounds__(256) misfit_kernel(const double *__restrict__ snap, int64_t n_rows, int64_t row_stride,
                                                     int64_t model_stride, const int64_t *__restrict__ obs_idx, int n_obs,
                                                     const double *__restrict__ obs, const double *__restrict__ obs_w,
                                                     int64_t first_row, double *__restrict__ misfit) {
    __shared__ double s_part[8];
    const double *base = snap + blockIdx.x * model_stride;
    double acc = 0.0;
    const int64_t total = (n_rows - first_row) * n_obs;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int64_t t = first_row + e / n_obs;
        const int k = (int)(e % n_obs);
        const double r = base[t * row_stride + obs_idx[k]] - obs[t * n_obs + k];
        acc = fma(obs_w ? obs_w[k] * r : r, r, acc);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_part[w];
        misfit[blockIdx.x] = v;
    }
}

extern "C" int fbr_misfit_f64(const double *snap, int64_t n_models, int64_t n_rows, int64_t nx, int64_t snap_row_stride,
                              int64_t snap_model_stride, const int64_t *obs_idx, int n_obs, const double *obs,
                              const double *obs_w, int64_t first_row, double *misfit, void *stream) {
    FBR_REQUIRE(snap && obs_idx && obs && misfit, "NULL array argument");
    FBR_REQUIRE(n_models >= 1 && n_rows >= 1 && nx >= 1 && n_obs >= 1, "n_models, n_rows, nx and n_obs must be >= 1");
    FBR_REQUIRE(first_row >= 0 && first_row <= n_rows, "first_row outside [0, n_rows]");
    FBR_REQUIRE(snap_row_stride >= nx && snap_model_stride >= nx, "snapshot strides must be >= nx");
    misfit_kernel<<<(unsigned)n_models, 256, 0, as_stream(stream)>>>(snap, n_rows, snap_row_stride, snap_model_stride, obs_idx,
                                                                      n_obs, obs, obs_w, first_row, misfit);
    return check_launch("fbr_misfit_f64");
}

// One front door for the scratch sizes ("library allocates nothing persistent").
extern "C" int64_t fbr_workspace_bytes(int what, int64_t M, int64_t nx) {
    switch (what) {
        case FBR_WS_FACTORIZE: return fbr_factorize_work_bytes(M, nx);
        case FBR_WS_SOLVE: return fbr_solve_tridiag_work_bytes(M, nx, FBR_SOLVER_AUTO);
        case FBR_WS_SOLVE_THOMAS: return fbr_solve_tridiag_work_bytes(M, nx, FBR_SOLVER_THOMAS);
        case FBR_WS_RUN: return 0;
        case FBR_WS_RUN_LONG: return fbr_run_long_work_bytes(nx);
        case FBR_WS_RUN_STREAM: return fbr_run_stream_work_bytes(nx);
        case FBR_WS_RUN_WINDOW: return fbr_run_window_work_bytes(nx);
        case FBR_WS_RUN_PASS: return fbr_run_pass_work_bytes(nx);
        default: fail(FBR_ERR_BAD_ARG, "unknown workspace kind %d", what); return -1;
    }
}

extern "C" int fbr_version(void) { return FBR_VERSION; }
extern "C" const char *fbr_last_error(void) { return last_error_ref().c_str(); }
extern "C" int fbr_last_launch_count(void) { return launch_count_ref(); }

// (rows, cols) -> (cols, rows), 32 x 32 tiles through shared memory (padded: conflict-free), both sides coalesced.
// Turns the (time, cell) rows the long-mesh kernels write into the (cell, time) layout of Data2D / pack_result
// (src/fiberis/simulator/core/pds.py:415) on the device: 1e7 cells x 11 rows in 0.4 ms instead of a host-side transpose.
__global__ void __launch_bounds__(256) transpose_kernel(const double *__restrict__ src, int64_t rows, int64_t cols,
                                                        double *__restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    for (int64_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
#pragma unroll
        for (int k = 0; k < 32; k += 8)
            if (r0 + ty + k < rows && c0 + tx < cols) tile[ty + k][tx] = src[(r0 + ty + k) * cols + c0 + tx];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; k += 8)
            if (c0 + ty + k < cols && r0 + tx < rows) dst[(c0 + ty + k) * rows + r0 + tx] = tile[tx][ty + k];
        __syncthreads();
    }
}

extern "C" int fbr_transpose_f64(const double *src, int64_t rows, int64_t cols, double *dst, void *stream) {
    FBR_REQUIRE(src && dst && rows >= 1 && cols >= 1, "NULL pointer or empty matrix");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    const int64_t tiles = ((cols + 31) / 32) * ((rows + 31) / 32);
    const int64_t cap = (int64_t)f.sm_count * 16;
    transpose_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, as_stream(stream)>>>(src, rows, cols, dst);
    return check_launch("fbr_transpose_f64");
}

// fp64 issue peak of this device, measured: every thread runs 8 independent DFMA chains, 16 warps per SM (the
// residency of K3), so the fp64 pipe — not latency — bounds the loop.  The denominator of bench.py's compute roofline.
__global__ void __launch_bounds__(512, 1) fp64_peak_kernel(double *sink, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;  // keeps the chains alive
}

extern "C" int fbr_probe_fp64_peak(double *h_dfma_per_s, double *d_sink, void *stream_v) {
    FBR_REQUIRE(h_dfma_per_s && d_sink, "NULL output pointer");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    cudaStream_t stream = as_stream(stream_v);
    cudaEvent_t e0, e1;
    FBR_CUDA(cudaEventCreate(&e0));
    FBR_CUDA(cudaEventCreate(&e1));
    const int iters = 4096;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition warms up
        FBR_CUDA(cudaEventRecord(e0, stream));
        fp64_peak_kernel<<<f.sm_count, 512, 0, stream>>>(d_sink, iters, 1.0);
        FBR_CUDA(cudaEventRecord(e1, stream));
        FBR_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FBR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (int rc = check_launch("fbr_probe_fp64_peak")) return rc;
    *h_dfma_per_s = (double)f.sm_count * 512.0 * iters * 64.0 / (best * 1e-3);  // lane-level DFMA per second
    return FBR_OK;
}

// The one collective of the path — the all-gather of per-member misfits over the ranks' contiguous shards —
// for callers that bring their own communicator: the library does not link NCCL, it resolves ncclAllGather from the NCCL
// the process already has loaded (or can load) at first use.
extern "C" int fbr_nccl_allgather_f64(const double *send, double *recv, int64_t count, void *nccl_comm, void *stream) {
    FBR_REQUIRE(send && recv && count >= 0 && nccl_comm, "NULL buffer / communicator");
    typedef int (*allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
    static allgather_fn fn = nullptr;
    if (!fn) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(FBR_ERR_UNSUPPORTED, "NCCL is not available in this process (%s)", dlerror());
        fn = reinterpret_cast<allgather_fn>(dlsym(h, "ncclAllGather"));
        if (!fn) return fail(FBR_ERR_UNSUPPORTED, "ncclAllGather not found in the loaded NCCL");
    }
    const int kNcclFloat64 = 8;  // ncclDataType_t: ncclFloat64 / ncclDouble
    const int rc = fn(send, recv, (size_t)count, kNcclFloat64, nccl_comm, as_stream(stream));
    if (rc != 0) return fail(FBR_ERR_CUDA, "ncclAllGather failed with ncclResult_t %d", rc);
    return FBR_OK;
}

extern "C" int fbr_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *smem_per_block_optin,
                               int64_t *l2_bytes, int64_t *hbm_bytes) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    if (sm_count) *sm_count = f.sm_count;
    if (cc_major) *cc_major = f.cc_major;
    if (cc_minor) *cc_minor = f.cc_minor;
    if (smem_per_block_optin) *smem_per_block_optin = f.smem_optin;
    if (l2_bytes) *l2_bytes = f.l2_bytes;
    if (hbm_bytes) *hbm_bytes = f.hbm_bytes;
    return FBR_OK;
}
