// K3S  streaming time integration for meshes too long to stay on chip (BASELINE.json configs[4]:
// 1e7 cells).  State and factors live in HBM; every step streams them through a hierarchy of
// first-order linear recurrences  y_i = s_i + m_i y_{i-1}  (forward: m = -l, s = b; backward, on
// reversed indices: m = -g, s = r y):
//     level 0: one THREAD per chunk of kLc consecutive cells folds its chunk with zero carry-in,
//     level 1..: the chunk-end values form the same kind of recurrence, kLc times shorter; it is
//                folded again until a single CTA can scan it in shared memory,
//     then the carries are applied on the way down and every thread re-runs its chunk with the
//     true carry-in.
// Chunk multipliers (products of m) are time-invariant and computed once.  Arrays are stored
// CHUNK-INTERLEAVED (element j of chunk c at j * n_chunks + c) so that a warp of 32 chunks reads 32
// consecutive doubles per cell index: every global access of the hot kernels is fully coalesced.
// HBM-bound; algorithmic floor of this formulation: 88 B per cell-update.
#include "fbr_common.cuh"

namespace fbr {

constexpr int kLc = 32;            // cells (or sub-chunks) per thread
constexpr int kTopMax = 1024;      // a level this short is scanned by one CTA
constexpr int kMaxLevels = 8;

// address of element j of chunk c (nC chunks); REV walks the array from its far end
template <bool REV>
__device__ __forceinline__ int64_t at(int64_t c, int j, int64_t nC) {
    return REV ? (int64_t)(kLc - 1 - j) * nC + (nC - 1 - c) : (int64_t)j * nC + c;
}
// where chunk c's end value goes in the next (interleaved) level
__device__ __forceinline__ int64_t next_slot(int64_t c, int64_t nC_next) { return (c % kLc) * nC_next + c / kLc; }

// natural order -> interleaved (padding cells get `pad`), and back
__global__ void __launch_bounds__(256) interleave_kernel(const double *__restrict__ src, int64_t n, int64_t nC,
                                                         double pad, double scale, double *__restrict__ dst) {
    const int64_t total = nC * kLc;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = e / nC, c = e - j * nC, i = c * kLc + j;  // coalesced writes
        dst[e] = i < n ? scale * src[i] : pad;
    }
}
__global__ void __launch_bounds__(256) deinterleave_kernel(const double *__restrict__ src, int64_t n, int64_t nC,
                                                           double *__restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[(i % kLc) * nC + i / kLc];  // coalesced writes
}

// chunk products of the multipliers (setup)
template <bool REV>
__global__ void __launch_bounds__(128) level_prod_kernel(const double *__restrict__ mult, int64_t nC, int64_t nC_next,
                                                         double *__restrict__ prod_next) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nC) return;
    double p = 1.0;
#pragma unroll 8
    for (int j = 0; j < kLc; ++j) p *= mult[at<REV>(c, j, nC)];
    prod_next[next_slot(c, nC_next)] = p;
}
__global__ void pad_kernel(double *a, int64_t from, int64_t to, double v) {
    for (int64_t i = from + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < to; i += (int64_t)gridDim.x * blockDim.x) a[i] = v;
}

// up-sweep: end value of every chunk with zero carry-in.  NEG: the stored array is l (or g), m = -array.
template <bool REV, bool NEG>
__global__ void __launch_bounds__(128) level_reduce_kernel(const double *__restrict__ mult, const double *__restrict__ val,
                                                           int64_t nC, int64_t nC_next, double *__restrict__ end_next) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nC) return;
    double y = 0.0;
#pragma unroll 8
    for (int j = 0; j < kLc; ++j) {
        const int64_t a = at<REV>(c, j, nC);
        const double m = mult[a];
        y = NEG ? fma(-m, y, val[a]) : fma(m, y, val[a]);
    }
    end_next[next_slot(c, nC_next)] = y;
}

// down-sweep: re-run the chunk with the true carry-in (the solved value of the previous chunk at the
// next level).  SCALE: out = scale[i] * y (the forward sweep hands t = r y to the backward sweep).
template <bool REV, bool NEG, bool SCALE>
__global__ void __launch_bounds__(128) level_apply_kernel(const double *__restrict__ mult, const double *val,
                                                          const double *__restrict__ solved_next,
                                                          const double *__restrict__ scale, int64_t nC, int64_t nC_next,
                                                          double *out) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nC) return;
    double y = c > 0 ? solved_next[next_slot(c - 1, nC_next)] : 0.0;
#pragma unroll 8
    for (int j = 0; j < kLc; ++j) {
        const int64_t a = at<REV>(c, j, nC);
        const double m = mult[a];
        y = NEG ? fma(-m, y, val[a]) : fma(m, y, val[a]);
        out[a] = SCALE ? scale[a] * y : y;
    }
}

// top of the hierarchy: n <= kLc * kTopMax values (interleaved, nC chunks) solved by one CTA:
// thread per chunk folds, Hillis-Steele scan of the (multiplier, value) pairs in shared memory, re-run.
__global__ void __launch_bounds__(kTopMax) level_top_kernel(const double *__restrict__ mult, const double *val,
                                                            int64_t nC, double *out) {
    __shared__ double s_p[kTopMax], s_v[kTopMax];
    const int c = threadIdx.x;
    double p = 1.0, y = 0.0;
    if (c < nC) {
        for (int j = 0; j < kLc; ++j) {
            const int64_t a = (int64_t)j * nC + c;
            const double m = mult[a];
            y = fma(m, y, val[a]);
            p *= m;
        }
    }
    s_p[c] = p;
    s_v[c] = y;
    __syncthreads();
    for (int d = 1; d < kTopMax; d <<= 1) {  // inclusive scan of affine maps
        double pp = 1.0, vv = 0.0;
        const bool on = c >= d;
        if (on) { pp = s_p[c - d]; vv = s_v[c - d]; }
        __syncthreads();
        if (on) { s_v[c] = fma(s_p[c], vv, s_v[c]); s_p[c] *= pp; }
        __syncthreads();
    }
    if (c < nC) {
        double yy = c > 0 ? s_v[c - 1] : 0.0;
        for (int j = 0; j < kLc; ++j) {
            const int64_t a = (int64_t)j * nC + c;
            yy = fma(mult[a], yy, val[a]);
            out[a] = yy;
        }
    }
}

// right-hand-side overrides applied in place to the interleaved state (a handful of cells per step)
__global__ void stream_override_kernel(double *state, int64_t nx, int64_t nC, int lbc, int rbc,
                                       const int64_t *__restrict__ src_idx, int n_src, const double *__restrict__ src_row) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // sequential: later duplicates win, sources beat boundary rows
        if (lbc != FBR_BC_NONE) state[0] = 0.0;
        if (rbc != FBR_BC_NONE) state[((nx - 1) % kLc) * nC + (nx - 1) / kLc] = 0.0;
        for (int k = 0; k < n_src; ++k) {
            const int64_t i = src_idx[k];
            if (i >= 0 && i < nx) state[(i % kLc) * nC + i / kLc] = src_row[k];
        }
    }
}

struct StreamPlan {
    int levels = 0;             // number of chunked levels (level 0 = cells)
    int64_t nC[kMaxLevels + 1]; // chunks at level k (= elements of level k + 1)
    int64_t words = 0;          // doubles of workspace
};

static StreamPlan plan_stream(int64_t nx) {
    StreamPlan p;
    int64_t n = nx;
    for (;;) {
        const int64_t c = (n + kLc - 1) / kLc;
        p.nC[p.levels++] = c;
        if (c <= kTopMax || p.levels == kMaxLevels) break;
        n = c;
    }
    // level 0: l, r, g, state, t  (5 arrays of nC0*kLc); per higher level and direction: mult, val, solved
    p.words = 5 * p.nC[0] * kLc;
    for (int k = 1; k < p.levels; ++k) p.words += 2 * 3 * p.nC[k] * kLc;
    p.words += 2 * 3 * (int64_t)kTopMax * kLc;  // top level arrays are at most this long
    return p;
}

}  // namespace fbr

using namespace fbr;

extern "C" int64_t fbr_run_stream_work_bytes(int64_t nx) { return plan_stream(nx).words * (int64_t)sizeof(double); }

#define LAUNCH_OK(what) do { if (int rc__ = check_launch(what)) return rc__; } while (0)

extern "C" int fbr_run_stream_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                  int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                  const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                                  double *snap, double *p_final, double *work, void *stream) {
    FBR_REQUIRE(nx >= 1 && n_steps >= 0, "bad nx / n_steps");
    FBR_REQUIRE(fl && fr && fg && p0 && work, "NULL factor / initial-state / work pointer");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_table)), "sources given without src_idx / src_table");
    FBR_REQUIRE(!snap || (snapshot_every >= 1 && snap_row_stride >= nx), "bad snapshot layout");
    const StreamPlan pl = plan_stream(nx);
    if (pl.nC[pl.levels - 1] > kTopMax) return fail(FBR_ERR_UNSUPPORTED, "mesh too long for %d levels", kMaxLevels);
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    cudaStream_t st = as_stream(stream);
    const int L = pl.levels;
    const int64_t n0 = pl.nC[0] * kLc;
    // ---- carve the workspace ----
    double *w = work;
    double *il = w; w += n0;
    double *ir = w; w += n0;
    double *ig = w; w += n0;
    double *state = w; w += n0;
    double *tt = w; w += n0;
    // level k >= 1 arrays hold nC[k-1] elements padded to nC[k] * kLc
    double *mult[2][kMaxLevels + 1], *val[2][kMaxLevels + 1], *sol[2][kMaxLevels + 1];
    for (int k = 1; k <= L - 1 + 0; ++k) {
        const int64_t len = pl.nC[k] * kLc;
        for (int dir = 0; dir < 2; ++dir) {
            mult[dir][k] = w; w += len;
            val[dir][k] = w; w += len;
            sol[dir][k] = w; w += len;
        }
    }
    // NOTE: level indices: level k has elements = chunks of level k-1; top level is L-1 with nC[L-1] chunks
    const int gs = f.sm_count * 8;
    auto blocks = [](int64_t n, int b) { return (unsigned)((n + b - 1) / b); };
    // ---- setup: interleave factors and the initial state; chunk products for every level ----
    interleave_kernel<<<gs, 256, 0, st>>>(fl, nx, pl.nC[0], 0.0, 1.0, il); LAUNCH_OK("stream: interleave l");
    interleave_kernel<<<gs, 256, 0, st>>>(fr, nx, pl.nC[0], 1.0, 1.0, ir); LAUNCH_OK("stream: interleave r");
    interleave_kernel<<<gs, 256, 0, st>>>(fg, nx, pl.nC[0], 0.0, 1.0, ig); LAUNCH_OK("stream: interleave g");
    interleave_kernel<<<gs, 256, 0, st>>>(p0, nx, pl.nC[0], 0.0, 1.0, state); LAUNCH_OK("stream: interleave p0");
    if (L >= 2) {
        // level-1 multipliers = products of (-l) / (-g) over each level-0 chunk
        for (int dir = 0; dir < 2; ++dir) {
            pad_kernel<<<64, 256, 0, st>>>(mult[dir][1], 0, pl.nC[1] * kLc, 0.0); LAUNCH_OK("stream: pad");
            pad_kernel<<<64, 256, 0, st>>>(val[dir][1], 0, pl.nC[1] * kLc, 0.0); LAUNCH_OK("stream: pad");
        }
        // kLc is even, so the product of -m equals the product of m
        level_prod_kernel<false><<<blocks(pl.nC[0], 128), 128, 0, st>>>(il, pl.nC[0], pl.nC[1], mult[0][1]); LAUNCH_OK("stream: prod");
        level_prod_kernel<true><<<blocks(pl.nC[0], 128), 128, 0, st>>>(ig, pl.nC[0], pl.nC[1], mult[1][1]); LAUNCH_OK("stream: prod");
        for (int k = 2; k <= L - 1; ++k)
            for (int dir = 0; dir < 2; ++dir) {
                pad_kernel<<<64, 256, 0, st>>>(mult[dir][k], 0, pl.nC[k] * kLc, 0.0); LAUNCH_OK("stream: pad");
                pad_kernel<<<64, 256, 0, st>>>(val[dir][k], 0, pl.nC[k] * kLc, 0.0); LAUNCH_OK("stream: pad");
                level_prod_kernel<false><<<blocks(pl.nC[k - 1], 128), 128, 0, st>>>(mult[dir][k - 1], pl.nC[k - 1], pl.nC[k], mult[dir][k]);
                LAUNCH_OK("stream: prod");
            }
    }
    // ---- one sweep = up, top, down ----
    auto sweep = [&](int dir) -> int {
        double *in_val = dir == 0 ? state : tt;     // forward reads b, backward reads t
        double *out = dir == 0 ? tt : state;        // forward writes t = r y, backward writes x
        if (L == 1) {  // the whole mesh fits the top kernel (only for tiny nx; kept for completeness)
            return fail(FBR_ERR_UNSUPPORTED, "streaming path is meant for nx > %d", kTopMax * kLc);
        }
        // up
        if (dir == 0) level_reduce_kernel<false, true><<<blocks(pl.nC[0], 128), 128, 0, st>>>(il, in_val, pl.nC[0], pl.nC[1], val[0][1]);
        else level_reduce_kernel<true, true><<<blocks(pl.nC[0], 128), 128, 0, st>>>(ig, in_val, pl.nC[0], pl.nC[1], val[1][1]);
        LAUNCH_OK("stream: reduce0");
        for (int k = 1; k <= L - 2; ++k) {
            level_reduce_kernel<false, false><<<blocks(pl.nC[k], 128), 128, 0, st>>>(mult[dir][k], val[dir][k], pl.nC[k], pl.nC[k + 1], val[dir][k + 1]);
            LAUNCH_OK("stream: reduce");
        }
        // top
        level_top_kernel<<<1, kTopMax, 0, st>>>(mult[dir][L - 1], val[dir][L - 1], pl.nC[L - 1], sol[dir][L - 1]);
        LAUNCH_OK("stream: top");
        // down
        for (int k = L - 2; k >= 1; --k) {
            level_apply_kernel<false, false, false><<<blocks(pl.nC[k], 128), 128, 0, st>>>(mult[dir][k], val[dir][k], sol[dir][k + 1], nullptr, pl.nC[k], pl.nC[k + 1], sol[dir][k]);
            LAUNCH_OK("stream: apply");
        }
        if (dir == 0) level_apply_kernel<false, true, true><<<blocks(pl.nC[0], 128), 128, 0, st>>>(il, in_val, sol[0][1], ir, pl.nC[0], pl.nC[1], out);
        else level_apply_kernel<true, true, false><<<blocks(pl.nC[0], 128), 128, 0, st>>>(ig, in_val, sol[1][1], nullptr, pl.nC[0], pl.nC[1], out);
        LAUNCH_OK("stream: apply0");
        return FBR_OK;
    };
    int64_t snap_row = 0;
    for (int64_t n = 0; n < n_steps; ++n) {
        stream_override_kernel<<<1, 32, 0, st>>>(state, nx, pl.nC[0], lbc, rbc, src_idx, n_src, src_table + n * n_src);
        LAUNCH_OK("stream: overrides");
        if (int rc = sweep(0)) return rc;
        if (int rc = sweep(1)) return rc;
        if (snap && (n + 1) % snapshot_every == 0) {
            deinterleave_kernel<<<gs, 256, 0, st>>>(state, nx, pl.nC[0], snap + snap_row * snap_row_stride);
            LAUNCH_OK("stream: snapshot");
            ++snap_row;
        }
    }
    if (p_final) {
        deinterleave_kernel<<<gs, 256, 0, st>>>(state, nx, pl.nC[0], p_final);
        LAUNCH_OK("stream: final");
    }
    return FBR_OK;
}
