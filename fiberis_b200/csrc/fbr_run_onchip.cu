// K3  persistent on-chip time integration (the hot kernel).
//
// One CTA ("team") owns one model at a time and integrates ALL n_steps of it before taking the
// next model: pressure and the three time-invariant factor arrays live in REGISTERS for the whole
// run, so the only per-step memory traffic is what the caller asked to be written (snapshots).
// Models are independent, so nothing forces them to advance in lock step; time
// blocking per model removes the 16 B/cell-update state stream and the 24 B/cell-update factor
// stream that a step-at-a-time batched solver would pull from HBM/L2.
//
// Per step the team solves  L U x = b  with the factors of fbr_factorize_f64:
//     forward   y_i = b_i - l_i y_{i-1}
//     backward  x_i = r_i y_i - g_i x_{i+1}
// Both are first-order linear recurrences.  Each thread owns CPT consecutive cells; a sweep is
//   (1) a local pass with zero carry-in (CPT dependent FMAs),
//   (2) a Kogge-Stone scan of the chunk-end values across the warp (5 shuffle+FMA stages) — the
//       multipliers of every stage are products of l (or g) only, i.e. time-invariant, so they are
//       computed once per model and kept in registers,
//   (3) a carry across the warps of the team through shared memory (W-1 FMAs),
//   (4) a second local pass with the true carry-in.
// No divisions, no HBM traffic and two __syncthreads per step (none for single-warp teams).
#include <cuda_pipeline.h>
#include <stdarg.h>

#include "fbr_common.cuh"

namespace fbr {

constexpr int kTileSteps = 16;    // steps of source / observation values staged per smem tile
constexpr int kMaxThreads = 512;  // per team
constexpr int kMaxSrc = 64;
constexpr int kMaxObs = 64;
constexpr unsigned kFull = 0xffffffffu;

struct RunArgs {
    const double *fl, *fr, *fg;  // (M, nx) factors
    const double *p0;
    int64_t p0_stride;
    int64_t M, nx, n_steps;
    int lbc, rbc;
    const int64_t *src_idx;
    int n_src;
    const double *src_table;  // (n_steps, n_src)
    int64_t snapshot_every;
    const int64_t *snap_models;
    int64_t n_snap_models;
    void *snap;     // ((n_steps / snapshot_every), n_snap_models, nx) of T
    void *p_final;  // (M, nx) of T
    const int64_t *obs_idx;
    int n_obs;
    const double *obs;    // (n_steps + 1, n_obs)
    const double *obs_w;  // (n_obs) or NULL
    double *misfit;       // (M)
};

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// slot codes packed one byte per owned cell
constexpr unsigned kSlotNone = 0xFF, kSlotZero = 0xFE;

template <typename T, int CPT>
__global__ void __launch_bounds__(kMaxThreads, 1) run_onchip_kernel(const RunArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = blockDim.x >> 5;
    // smem carve-up: [Tf W][Tb W][Ef W][Eb W] of T, misfit scratch W doubles, snap slot, tiles
    T *Tf = reinterpret_cast<T *>(smem_raw);
    T *Tb = Tf + W;
    T *Ef = Tb + W;
    T *Eb = Ef + W;
    double *red = reinterpret_cast<double *>(smem_raw + ((4 * W * sizeof(T) + 15) / 16) * 16);
    long long *snap_slot_s = reinterpret_cast<long long *>(red + W);
    const int tile_row = a.n_src + a.n_obs;  // doubles per staged step
    double *tile = reinterpret_cast<double *>(snap_slot_s + 2);  // [2][kTileSteps][tile_row]
    double *wobs = tile + 2 * kTileSteps * tile_row;             // [n_obs] observation weights

    const int64_t nx = a.nx;
    const int64_t cell0 = (int64_t)tid * CPT;

    for (int64_t m = blockIdx.x; m < a.M; m += gridDim.x) {
        // ---------------- per-model setup ----------------
        T l[CPT], r[CPT], g[CPT], x[CPT];
        uint64_t ov = ~0ull, ob = ~0ull;  // override / observation slot per owned cell
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int64_t i = cell0 + j;
            if (i < nx) {
                l[j] = (T)a.fl[m * nx + i];
                r[j] = (T)a.fr[m * nx + i];
                g[j] = (T)a.fg[m * nx + i];
                x[j] = (T)a.p0[m * a.p0_stride + i];
                unsigned slot = kSlotNone;
                if ((i == 0 && a.lbc != FBR_BC_NONE) || (i == nx - 1 && a.rbc != FBR_BC_NONE)) slot = kSlotZero;
                for (int k = 0; k < a.n_src; ++k)
                    if (a.src_idx[k] == i) slot = (unsigned)k;
                ov = (ov & ~(0xFFull << (8 * j))) | ((uint64_t)slot << (8 * j));
                unsigned oslot = kSlotNone;
                for (int k = 0; k < a.n_obs; ++k)
                    if (a.obs_idx[k] == i) oslot = (unsigned)k;
                ob = (ob & ~(0xFFull << (8 * j))) | ((uint64_t)oslot << (8 * j));
            } else {  // padding cell: identity row, decoupled
                l[j] = (T)0; r[j] = (T)1; g[j] = (T)0; x[j] = (T)0;
            }
        }
        const bool has_ov = ov != ~0ull, has_ob = ob != ~0ull;

        // scan multipliers (time-invariant): stage products of the chunk multipliers
        T PfL[5], PbL[5], PfEx, PbEx;
        {
            T pf = (T)1, pb = (T)1;
#pragma unroll
            for (int j = 0; j < CPT; ++j) { pf *= -l[j]; pb *= -g[j]; }
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                PfL[s] = pf;
                PbL[s] = pb;
                const T upf = __shfl_up_sync(kFull, pf, 1 << s);
                const T dnb = __shfl_down_sync(kFull, pb, 1 << s);
                if (lane >= (1 << s)) pf *= upf;
                if (lane + (1 << s) < 32) pb *= dnb;
            }
            // pf = product over lanes 0..lane, pb = product over lanes lane..31
            PfEx = __shfl_up_sync(kFull, pf, 1);
            PbEx = __shfl_down_sync(kFull, pb, 1);
            if (lane == 0) PfEx = (T)1;
            if (lane == 31) PbEx = (T)1;
            if (W > 1) {
                __syncthreads();  // previous model's readers are done with Tf/Tb
                if (lane == 31) Tf[warp] = pf;
                if (lane == 0) Tb[warp] = pb;
            }
        }
        // snapshot slot of this model (-1: not recorded)
        if (tid == 0) {
            long long slot = -1;
            if (a.snap) {
                if (!a.snap_models) slot = m;
                else
                    for (int64_t q = 0; q < a.n_snap_models; ++q)
                        if (a.snap_models[q] == m) slot = q;
            }
            snap_slot_s[0] = slot;
        }
        // first tile of staged per-step values
        auto stage_tile = [&](int64_t tile_idx) {
            if (tile_row == 0) return;
            double *dst = tile + (tile_idx & 1) * (kTileSteps * tile_row);
            const int64_t n0 = tile_idx * kTileSteps;
            for (int e = tid; e < kTileSteps * tile_row; e += blockDim.x) {
                const int s = e / tile_row, c = e - s * tile_row;
                const int64_t n = n0 + s;
                if (n < a.n_steps) {
                    const double *src = c < a.n_src ? a.src_table + n * a.n_src + c
                                                    : a.obs + (n + 1) * a.n_obs + (c - a.n_src);
                    __pipeline_memcpy_async(dst + e, src, sizeof(double));
                }
            }
            __pipeline_commit();
        };
        stage_tile(0);
        for (int k = tid; k < a.n_obs; k += blockDim.x) wobs[k] = a.obs_w ? a.obs_w[k] : 1.0;
        __syncthreads();
        const long long snap_slot = snap_slot_s[0];
        double acc = 0.0;
        int64_t snap_countdown = a.snapshot_every, snap_row = 0;
        int in_tile = 0;
        int64_t tile_idx = 0;

        // ---------------- time loop ----------------
        for (int64_t n = 0; n < a.n_steps; ++n) {
            if (in_tile == kTileSteps) { in_tile = 0; ++tile_idx; }
            if (in_tile == 0) {
                __pipeline_wait_prior(0);
                __syncthreads();  // tile tile_idx is visible; the other buffer is free
                stage_tile(tile_idx + 1);
            }
            const double *row = tile + (tile_idx & 1) * (kTileSteps * tile_row) + in_tile * tile_row;
            ++in_tile;

            // right-hand side: b = p^n with boundary / source overrides (matbuilder.py:45-76)
            if (has_ov) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const unsigned slot = (unsigned)(ov >> (8 * j)) & 0xFF;
                    if (slot == kSlotZero) x[j] = (T)0;
                    else if (slot != kSlotNone) x[j] = (T)row[slot];
                }
            }
            // ---- forward sweep ----
            T y = (T)0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) y = fma_t<T>(-l[j], y, x[j]);
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const T up = __shfl_up_sync(kFull, y, 1 << s);
                if (lane >= (1 << s)) y = fma_t<T>(PfL[s], up, y);
            }
            T carry = __shfl_up_sync(kFull, y, 1);
            if (lane == 0) carry = (T)0;
            if (W > 1) {
                if (lane == 31) Ef[warp] = y;
                __syncthreads();
                T c = (T)0;
                for (int v = 0; v < warp; ++v) c = fma_t<T>(Tf[v], c, Ef[v]);
                carry = fma_t<T>(PfEx, c, carry);
            }
            y = carry;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                y = fma_t<T>(-l[j], y, x[j]);
                x[j] = r[j] * y;  // t_j = r_j y_j
            }
            // ---- backward sweep ----
            T z = (T)0;
#pragma unroll
            for (int j = CPT - 1; j >= 0; --j) z = fma_t<T>(-g[j], z, x[j]);
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const T dn = __shfl_down_sync(kFull, z, 1 << s);
                if (lane + (1 << s) < 32) z = fma_t<T>(PbL[s], dn, z);
            }
            carry = __shfl_down_sync(kFull, z, 1);
            if (lane == 31) carry = (T)0;
            if (W > 1) {
                if (lane == 0) Eb[warp] = z;
                __syncthreads();
                T c = (T)0;
                for (int v = W - 1; v > warp; --v) c = fma_t<T>(Tb[v], c, Eb[v]);
                carry = fma_t<T>(PbEx, c, carry);
            }
            z = carry;
#pragma unroll
            for (int j = CPT - 1; j >= 0; --j) {
                z = fma_t<T>(-g[j], z, x[j]);
                x[j] = z;
            }
            // ---- fused objective ----
            if (has_ob) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const unsigned slot = (unsigned)(ob >> (8 * j)) & 0xFF;
                    if (slot != kSlotNone) {
                        const double res = (double)x[j] - row[a.n_src + slot];
                        acc = fma(wobs[slot] * res, res, acc);
                    }
                }
            }
            // ---- sampled snapshot ----
            if (snap_slot >= 0 && --snap_countdown == 0) {
                snap_countdown = a.snapshot_every;
                T *dst = reinterpret_cast<T *>(a.snap) + (snap_row * a.n_snap_models + snap_slot) * nx + cell0;
                ++snap_row;
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    if (cell0 + j < nx) dst[j] = x[j];
            }
        }
        __pipeline_wait_prior(0);

        // ---------------- per-model epilogue ----------------
        if (a.p_final) {
            T *dst = reinterpret_cast<T *>(a.p_final) + m * nx + cell0;
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (cell0 + j < nx) dst[j] = x[j];
        }
        if (a.misfit) {
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
            if (lane == 0) red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < W; ++w) tot += red[w];
                a.misfit[m] = tot;
            }
        }
        __syncthreads();  // smem (tiles, red, snap slot) is reused by the next model
    }
}

template <typename T, int CPT>
static int launch_onchip(const RunArgs &a, int threads, cudaStream_t stream) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    const int W = threads / 32;
    size_t smem = ((4 * W * sizeof(T) + 15) / 16) * 16 + W * sizeof(double) + 2 * sizeof(long long) +
                  2 * (size_t)kTileSteps * (a.n_src + a.n_obs) * sizeof(double) + (size_t)a.n_obs * sizeof(double);
    auto kern = run_onchip_kernel<T, CPT>;
    if (smem > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "run kernel does not fit on an SM (threads=%d smem=%zu)", threads, smem);
    int64_t grid = (int64_t)occ * f.sm_count;
    if (grid > a.M) grid = a.M;
    kern<<<(unsigned)grid, threads, smem, stream>>>(a);
    return check_launch("fbr_run (on-chip)");
}

template <typename T>
static int run_dispatch(const RunArgs &a, cudaStream_t stream) {
    FBR_REQUIRE(a.M >= 1 && a.nx >= 1 && a.n_steps >= 0, "M, nx must be >= 1 and n_steps >= 0");
    FBR_REQUIRE(a.fl && a.fr && a.fg && a.p0, "NULL factor / initial-state pointer");
    FBR_REQUIRE(a.p0_stride == 0 || a.p0_stride >= a.nx, "bad p0_stride");
    FBR_REQUIRE(a.lbc >= 0 && a.lbc <= 2 && a.rbc >= 0 && a.rbc <= 2, "bad boundary code");
    FBR_REQUIRE(a.n_src >= 0 && a.n_src <= kMaxSrc, "n_src=%d outside [0, %d]", a.n_src, kMaxSrc);
    FBR_REQUIRE(a.n_src == 0 || (a.src_idx && a.src_table), "sources given without src_idx / src_table");
    FBR_REQUIRE(a.n_obs >= 0 && a.n_obs <= kMaxObs, "n_obs=%d outside [0, %d]", a.n_obs, kMaxObs);
    FBR_REQUIRE(a.n_obs == 0 || (a.obs_idx && a.obs && a.misfit), "observations given without obs / misfit");
    FBR_REQUIRE(!a.snap || (a.snapshot_every >= 1 && a.n_snap_models >= 1), "snapshot_every / n_snap_models must be >= 1");
    FBR_REQUIRE(!a.snap || a.snap_models || a.n_snap_models == a.M, "snap_models == NULL needs n_snap_models == M");
    if (a.nx > fbr_run_max_nx())
        return fail(FBR_ERR_UNSUPPORTED, "nx=%lld exceeds the on-chip run kernel limit %lld", (long long)a.nx,
                    (long long)fbr_run_max_nx());
    int cpt = a.nx >= 256 ? 8 : a.nx >= 128 ? 4 : a.nx >= 64 ? 2 : 1;
    if (const char *e = getenv("FBR_CPT")) {  // tuning aid
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4 || v == 8) && (a.nx + v - 1) / v <= kMaxThreads) cpt = v;
    }
    const int threads = (int)(((a.nx + cpt - 1) / cpt + 31) / 32) * 32;
    switch (cpt) {
        case 1: return launch_onchip<T, 1>(a, threads, stream);
        case 2: return launch_onchip<T, 2>(a, threads, stream);
        case 4: return launch_onchip<T, 4>(a, threads, stream);
        default: return launch_onchip<T, 8>(a, threads, stream);
    }
}

}  // namespace fbr

using namespace fbr;

extern "C" int64_t fbr_run_max_nx(void) { return 8 * (int64_t)kMaxThreads; }

#define FBR_FILL_ARGS()                                                                                   \
    RunArgs a;                                                                                            \
    a.fl = fl; a.fr = fr; a.fg = fg; a.p0 = p0; a.p0_stride = p0_stride; a.M = M; a.nx = nx;              \
    a.n_steps = n_steps; a.lbc = lbc; a.rbc = rbc; a.src_idx = src_idx; a.n_src = n_src;                  \
    a.src_table = src_table; a.snapshot_every = snapshot_every; a.snap_models = snap_models;              \
    a.n_snap_models = n_snap_models; a.snap = snap; a.p_final = p_final; a.obs_idx = obs_idx;             \
    a.n_obs = obs_idx ? n_obs : 0; a.obs = obs; a.obs_w = obs_w; a.misfit = misfit;

extern "C" int fbr_run_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                           int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                           const int64_t *src_idx, int n_src, const double *src_table, int64_t snapshot_every,
                           const int64_t *snap_models, int64_t n_snap_models, double *snap, double *p_final,
                           const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                           double *misfit, void *stream) {
    FBR_FILL_ARGS();
    return run_dispatch<double>(a, as_stream(stream));
}

extern "C" int fbr_run_f32(const double *fl, const double *fr, const double *fg, const double *p0,
                           int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                           const int64_t *src_idx, int n_src, const double *src_table, int64_t snapshot_every,
                           const int64_t *snap_models, int64_t n_snap_models, float *snap, float *p_final,
                           const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                           double *misfit, void *stream) {
    FBR_FILL_ARGS();
    return run_dispatch<float>(a, as_stream(stream));
}
