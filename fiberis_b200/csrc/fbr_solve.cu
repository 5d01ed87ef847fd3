// K2  bare batched tridiagonal solves (matrix changes every call: adaptive time stepping) and the
// Jacobi "explicit" mode.  Three variants, chosen by shape (north_star: "variant chosen by mesh length"):
//   THOMAS  one thread per system, sequential elimination      -> systems longer than 4096 cells
//   PCR     one CTA per system, shared-memory parallel cyclic reduction, log2(nx) levels
//   REG     register-resident Moebius-scan pivots + scan sweeps (fbr_solve_reg.cu) -> AUTO's choice up to 4096 cells
#include "fbr_common.cuh"

namespace fbr {

constexpr int kPcrMaxNx = 4096;
constexpr int kJacobiMaxNx = 8192;

// ---- THOMAS: thread per system; work holds the modified super-diagonal ----
__global__ void __launch_bounds__(128) thomas_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                     const double *__restrict__ du, const double *b, double *x,
                                                     int64_t M, int64_t nx, double *__restrict__ work,
                                                     int32_t *__restrict__ singular) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t o = m * nx;
    double cp = 0.0, xp = 0.0;
    int32_t bad = 0;
    for (int64_t i = 0; i < nx; ++i) {
        const double a = dl[o + i];
        const double den = d[o + i] - a * cp;
        if (!bad && (den == 0.0 || !isfinite(den))) bad = (int32_t)(i + 1);
        cp = du[o + i] / den;
        xp = (b[o + i] - a * xp) / den;
        work[o + i] = cp;
        x[o + i] = xp;
    }
    for (int64_t i = nx - 2; i >= 0; --i) {
        xp = x[o + i] - work[o + i] * xp;
        x[o + i] = xp;
    }
    if (singular) singular[m] = bad;
}

// ---- THOMAS through transposing tiles: the "interleaved layout" is made on the fly ----
// A warp owns 32 systems and walks the mesh in tiles of 32 cells: tiles are read / written with
// coalesced 256-byte requests (lane = cell) and processed with lane = system out of shared memory
// ([cell][33] doubles: conflict-free both ways).  Forward sweep stores the modified super-diagonal in
// `work` and the intermediate solution in x; the backward sweep walks the tiles in reverse.
// 72 B of traffic per unknown (40 B in, 16 B scratch out and in again, 8 B + 8 B for x).
constexpr int kThCells = 32, kThWarps = 2;

template <int NARR>
__device__ __forceinline__ void tile_load(double (*const (&dst)[NARR])[33], const double *const (&src)[NARR], int64_t m0,
                                          int rows, int64_t nx, int64_t c0, int cells, int lane) {
    for (int r0 = 0; r0 < rows; r0 += 8) {  // 8 * NARR independent loads in flight per thread
        double v[NARR][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool ok = r0 + k < rows && lane < cells;
            const int64_t o = (m0 + (ok ? r0 + k : 0)) * nx + c0 + (ok ? lane : 0);
#pragma unroll
            for (int q = 0; q < NARR; ++q) v[q][k] = src[q][o];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (r0 + k < rows && lane < cells) {
#pragma unroll
                for (int q = 0; q < NARR; ++q) dst[q][lane][r0 + k] = v[q][k];
            }
    }
}

__global__ void __launch_bounds__(32 * kThWarps) thomas_tiled_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                                     const double *__restrict__ du, const double *b, double *x,
                                                                     int64_t M, int64_t nx, double *__restrict__ work,
                                                                     int32_t *__restrict__ singular) {
    extern __shared__ double s_tile[];  // [warp][4][cell][33]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double(*s0)[33] = reinterpret_cast<double(*)[33]>(s_tile + (size_t)warp * 4 * kThCells * 33);
    double(*s1)[33] = s0 + kThCells;
    double(*s2)[33] = s1 + kThCells;
    double(*s3)[33] = s2 + kThCells;
    const int64_t m0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    if (m0 >= M) return;
    const int rows = M - m0 < 32 ? (int)(M - m0) : 32;
    double cp = 0.0, xp = 0.0;
    int32_t bad = 0;
    // ---- forward ----
    for (int64_t c0 = 0; c0 < nx; c0 += kThCells) {
        const int cells = nx - c0 < kThCells ? (int)(nx - c0) : kThCells;
        __syncwarp();
        {
            double(*const dst[4])[33] = {s0, s1, s2, s3};
            const double *const src[4] = {dl, d, du, b};
            tile_load<4>(dst, src, m0, rows, nx, c0, cells, lane);
        }
        __syncwarp();
        if (lane < rows) {
            for (int c = 0; c < cells; ++c) {
                const double a = s0[c][lane];
                const double den = s1[c][lane] - a * cp;
                if (!bad && (den == 0.0 || !isfinite(den))) bad = (int32_t)(c0 + c + 1);
                const double r = 1.0 / den;
                cp = s2[c][lane] * r;
                xp = (s3[c][lane] - a * xp) * r;
                s0[c][lane] = cp;
                s1[c][lane] = xp;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; ++r) {
            const int64_t o = (m0 + r) * nx + c0 + lane;
            if (lane < cells) { work[o] = s0[lane][r]; x[o] = s1[lane][r]; }
        }
    }
    // ---- backward ----
    const int64_t last0 = (nx - 1) / kThCells * kThCells;
    for (int64_t c0 = last0; c0 >= 0; c0 -= kThCells) {
        const int cells = nx - c0 < kThCells ? (int)(nx - c0) : kThCells;
        __syncwarp();
        {
            double(*const dst[2])[33] = {s0, s1};
            const double *const src[2] = {work, x};
            tile_load<2>(dst, src, m0, rows, nx, c0, cells, lane);
        }
        __syncwarp();
        if (lane < rows) {
            for (int c = cells - 1; c >= 0; --c) {
                if (c0 + c < nx - 1) xp = s1[c][lane] - s0[c][lane] * xp;
                else xp = s1[c][lane];
                s1[c][lane] = xp;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; ++r) {
            const int64_t o = (m0 + r) * nx + c0 + lane;
            if (lane < cells) x[o] = s1[lane][r];
        }
    }
    if (singular && lane < rows) singular[m0 + lane] = bad;
}

// ---- PCR: CTA per system; every level halves the coupling distance ----
template <int CELLS>
__global__ void __launch_bounds__(1024) pcr_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                   const double *__restrict__ du, const double *b, double *x,
                                                   int64_t nx64, int32_t *__restrict__ singular) {
    extern __shared__ double sm[];
    const int n = (int)nx64;
    double *sa = sm, *sb = sm + n, *sc = sm + 2 * n, *sr = sm + 3 * n;
    __shared__ int s_bad;
    const int64_t o = blockIdx.x * nx64;
    if (threadIdx.x == 0) s_bad = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sa[i] = dl[o + i];
        sb[i] = d[o + i];
        sc[i] = du[o + i];
        sr[i] = b[o + i];
    }
    __syncthreads();
    for (int s = 1; s < n; s <<= 1) {
        double na[CELLS], nb[CELLS], nc[CELLS], nr[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < n) {
                const int im = i - s, ip = i + s;
                double a = sa[i], bb = sb[i], c = sc[i], r = sr[i];
                double xa = 0.0, xc = 0.0;
                if (im >= 0) {
                    const double piv = sb[im];
                    if (piv == 0.0) s_bad = im + 1;
                    const double k1 = a / piv;
                    xa = -sa[im] * k1;
                    bb -= sc[im] * k1;
                    r -= sr[im] * k1;
                }
                if (ip < n) {
                    const double piv = sb[ip];
                    if (piv == 0.0) s_bad = ip + 1;
                    const double k2 = c / piv;
                    xc = -sc[ip] * k2;
                    bb -= sa[ip] * k2;
                    r -= sr[ip] * k2;
                }
                na[k] = xa; nb[k] = bb; nc[k] = xc; nr[k] = r;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const int i = threadIdx.x + k * blockDim.x;
            if (i < n) { sa[i] = na[k]; sb[i] = nb[k]; sc[i] = nc[k]; sr[i] = nr[k]; }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double piv = sb[i];
        if (piv == 0.0 || !isfinite(piv)) s_bad = i + 1;
        x[o + i] = sr[i] / piv;
    }
    __syncthreads();
    if (singular && threadIdx.x == 0) singular[blockIdx.x] = s_bad;
}

// ---- Jacobi (PDESolver_EXP.py:5-39): CTA per system, iterate in shared memory ----
__global__ void __launch_bounds__(1024) jacobi_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                      const double *__restrict__ du, const double *__restrict__ b,
                                                      double *__restrict__ x, int64_t nx64, double tol, int max_iter,
                                                      int32_t *__restrict__ iters) {
    extern __shared__ double sm[];
    const int n = (int)nx64;
    double *cur = sm, *nxt = sm + n;
    __shared__ double s_red[32];
    __shared__ int s_done;
    const int64_t o = blockIdx.x * nx64;
    for (int i = threadIdx.x; i < n; i += blockDim.x) cur[i] = 0.0;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    int used = -max_iter;
    for (int it = 0; it < max_iter; ++it) {
        double dmax = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double di = d[o + i], x0 = cur[i];
            double s = __dmul_rn(di, x0);
            if (i > 0) s = __dadd_rn(__dmul_rn(dl[o + i], cur[i - 1]), s);
            if (i < n - 1) s = __dadd_rn(s, __dmul_rn(du[o + i], cur[i + 1]));
            const double except_i = __dsub_rn(s, __dmul_rn(di, x0));
            const double v = __ddiv_rn(__dsub_rn(b[o + i], except_i), di);
            nxt[i] = v;
            dmax = fmax(dmax, fabs(v - x0));
        }
        for (int s = 16; s > 0; s >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, s));
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dmax;
        __syncthreads();
        if (threadIdx.x == 0) {
            double mx = 0.0;
            for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) mx = fmax(mx, s_red[w]);
            if (mx < tol) s_done = 1;
        }
        __syncthreads();
        double *t = cur; cur = nxt; nxt = t;
        if (s_done) { used = it + 1; break; }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[o + i] = cur[i];
    if (iters && threadIdx.x == 0) iters[blockIdx.x] = used;
}

}  // namespace fbr

using namespace fbr;

extern "C" int64_t fbr_solve_tridiag_work_bytes(int64_t M, int64_t nx, int variant) {
    if (variant == FBR_SOLVER_PIVOT) return 4 * M * nx * (int64_t)sizeof(double);
    if (variant == FBR_SOLVER_PCR || variant == FBR_SOLVER_REG) return 0;
    if (variant == FBR_SOLVER_AUTO && nx <= kSolveRegMaxNx) return 0;
    return M * nx * (int64_t)sizeof(double);
}

// ---- partial pivoting, where the reference has it --------------------------------------------------------------------
// The reference solves with scipy.linalg.solve -> LAPACK (src/fiberis/simulator/solver/PDESolver_IMP.py:27-28): Gaussian
// elimination with PARTIAL PIVOTING.  The simulator's own matrices never need the row swaps, so every hot kernel eliminates
// without them; a regular matrix whose un-pivoted elimination meets a zero pivot (PML on a Neumann row with dt * sigma = 1:
// diagonal -1 + 1 = 0) takes this path instead: the published DGTSV algorithm, one thread per system — row i is swapped
// with row i + 1 when |dl_{i+1}| > |d_i|, the swap's fill-in goes to a second super-diagonal.  work: 4 * M * nx doubles
// (copies of d, du, the fill-in du2 and the right-hand side); singular[m] = 1 + row of an exactly zero pivot.
__global__ void __launch_bounds__(64) gtsv_pivot_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                        const double *__restrict__ du, const double *__restrict__ b,
                                                        double *__restrict__ x, int64_t M, int64_t nx, double *__restrict__ work,
                                                        int32_t *__restrict__ singular) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double *sub = dl + m * nx;  // sub[i] = A[i, i-1]
    double *dd = work + (0 * M + m) * nx, *u1 = work + (1 * M + m) * nx, *u2 = work + (2 * M + m) * nx, *rhs = work + (3 * M + m) * nx;
    for (int64_t i = 0; i < nx; ++i) { dd[i] = d[m * nx + i]; u1[i] = du[m * nx + i]; u2[i] = 0.0; rhs[i] = b[m * nx + i]; }
    int bad = 0;
    double below = nx > 1 ? sub[1] : 0.0;  // the (possibly updated) entry below the current pivot
    for (int64_t i = 0; i + 1 < nx; ++i) {
        const double next_below = i + 2 < nx ? sub[i + 2] : 0.0;
        if (fabs(dd[i]) >= fabs(below)) {  // no row interchange
            if (dd[i] == 0.0) { if (!bad) bad = (int)i + 1; break; }
            const double fact = below / dd[i];
            dd[i + 1] -= fact * u1[i];
            rhs[i + 1] -= fact * rhs[i];
        } else {  // interchange rows i and i + 1
            const double fact = dd[i] / below;
            dd[i] = below;
            const double t = dd[i + 1];
            dd[i + 1] = u1[i] - fact * t;
            if (i + 2 < nx) { u2[i] = u1[i + 1]; u1[i + 1] = -fact * u2[i]; }
            u1[i] = t;
            const double tb = rhs[i];
            rhs[i] = rhs[i + 1];
            rhs[i + 1] = tb - fact * rhs[i + 1];
        }
        below = next_below;
    }
    if (!bad && dd[nx - 1] == 0.0) bad = (int)nx;
    if (singular) singular[m] = bad;
    if (bad) return;
    double *out = x + m * nx;
    out[nx - 1] = rhs[nx - 1] / dd[nx - 1];
    if (nx > 1) out[nx - 2] = (rhs[nx - 2] - u1[nx - 2] * out[nx - 1]) / dd[nx - 2];
    for (int64_t i = nx - 3; i >= 0; --i) out[i] = (rhs[i] - u1[i] * out[i + 1] - u2[i] * out[i + 2]) / dd[i];
}

// A general dense system (API completeness of solver_implicit: the simulator itself never builds one): Gaussian elimination
// with partial pivoting by ONE CTA on a working copy [A | b] of n x (n + 1) doubles in global memory.
__global__ void __launch_bounds__(256) dense_pivot_kernel(double *__restrict__ Ab, int n, double *__restrict__ x, int32_t *singular) {
    __shared__ int s_piv;
    __shared__ double s_red_v[256];
    __shared__ int s_red_i[256];
    const int tid = threadIdx.x, ld = n + 1;
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int row = k;
        for (int i = k + tid; i < n; i += blockDim.x) {
            const double v = fabs(Ab[(int64_t)i * ld + k]);
            if (v > best) { best = v; row = i; }
        }
        s_red_v[tid] = best; s_red_i[tid] = row;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (tid < s && (s_red_v[tid + s] > s_red_v[tid] || (s_red_v[tid + s] == s_red_v[tid] && s_red_i[tid + s] < s_red_i[tid]))) {
                s_red_v[tid] = s_red_v[tid + s]; s_red_i[tid] = s_red_i[tid + s];
            }
            __syncthreads();
        }
        if (tid == 0) s_piv = s_red_i[0];
        __syncthreads();
        if (s_red_v[0] == 0.0) { if (tid == 0 && singular) *singular = k + 1; return; }
        const int p = s_piv;
        if (p != k)
            for (int j = k + tid; j < ld; j += blockDim.x) {
                const double t = Ab[(int64_t)k * ld + j];
                Ab[(int64_t)k * ld + j] = Ab[(int64_t)p * ld + j];
                Ab[(int64_t)p * ld + j] = t;
            }
        __syncthreads();
        const double piv = Ab[(int64_t)k * ld + k];
        for (int i = k + 1 + (tid >> 5); i < n; i += blockDim.x >> 5) {  // a warp per row
            const double f = Ab[(int64_t)i * ld + k] / piv;
            __syncwarp();  // (everybody has read column k of this row before lane 0's columns overwrite it)
            for (int j = k + 1 + (tid & 31); j < ld; j += 32) Ab[(int64_t)i * ld + j] -= f * Ab[(int64_t)k * ld + j];
        }
        __syncthreads();
    }
    for (int i = n - 1; i >= 0; --i) {  // back substitution, one row at a time
        double part = 0.0;
        for (int j = i + 1 + tid; j < n; j += blockDim.x) part += Ab[(int64_t)i * ld + j] * x[j];
        s_red_v[tid] = part;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (tid < s) s_red_v[tid] += s_red_v[tid + s];
            __syncthreads();
        }
        if (tid == 0) x[i] = (Ab[(int64_t)i * ld + n] - s_red_v[0]) / Ab[(int64_t)i * ld + i];
        __syncthreads();
    }
    if (tid == 0 && singular) *singular = 0;
}

extern "C" int fbr_solve_dense_f64(double *Ab, int64_t n, double *x, int32_t *singular, void *stream) {
    FBR_REQUIRE(Ab && x && n >= 1 && n <= 16384, "bad argument (1 <= n <= 16384)");
    dense_pivot_kernel<<<1, 256, 0, as_stream(stream)>>>(Ab, (int)n, x, singular);
    return check_launch("fbr_solve_dense_f64");
}

// Jacobi iteration on a general dense system (API completeness of solver_explicit, PDESolver_EXP.py:5-39: x starts at 0,
// x_new = (b - R x) / diag(A), stop when max |x_new - x| < tol).  ONE CTA of 32 warps: a warp per row (coalesced reads of
// the row, the diagonal term left out of the sum as the reference's R = A - diag(A) does), iterates ping-pong between x
// and work; iters = iterations used, or -max_iter when the tolerance was not met.
__global__ void __launch_bounds__(1024) jacobi_dense_kernel(const double *__restrict__ A, const double *__restrict__ b, double *x,
                                                            double *work, int n, double tol, int max_iter, int *iters) {
    __shared__ double s_max[32];
    __shared__ int s_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
    double *cur = x, *nxt = work;
    for (int i = tid; i < n; i += blockDim.x) cur[i] = 0.0;
    __syncthreads();
    int used = -max_iter;
    for (int it = 0; it < max_iter; ++it) {
        double worst = 0.0;
        for (int i = warp; i < n; i += warps) {
            const double *row = A + (int64_t)i * n;
            double sum = 0.0;
            for (int j = lane; j < n; j += 32)
                if (j != i) sum = fma(row[j], cur[j], sum);
            for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
            const double v = (b[i] - sum) / row[i];
            if (lane == 0) nxt[i] = v;
            const double diff = fabs(v - cur[i]);
            if (!(diff <= worst)) worst = diff;  // NaN propagates: never "converged"
        }
        if (lane == 0) s_max[warp] = worst;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int w = 0; w < warps; ++w)
                if (!(s_max[w] <= m)) m = s_max[w];
            s_done = m < tol;
        }
        __syncthreads();
        double *t = cur; cur = nxt; nxt = t;
        if (s_done) { used = it + 1; break; }
    }
    if (cur != x)
        for (int i = tid; i < n; i += blockDim.x) x[i] = cur[i];
    if (tid == 0 && iters) *iters = used;
}

extern "C" int fbr_jacobi_dense_f64(const double *A, const double *b, double *x, int64_t n, double tol, int max_iter,
                                    int32_t *iters, double *work, void *stream) {
    FBR_REQUIRE(A && b && x && work && n >= 1 && n <= 16384 && max_iter >= 0, "bad argument (1 <= n <= 16384)");
    jacobi_dense_kernel<<<1, 1024, 0, as_stream(stream)>>>(A, b, x, work, (int)n, tol, max_iter, iters);
    return check_launch("fbr_jacobi_dense_f64");
}

extern "C" int fbr_solve_tridiag_f64(const double *dl, const double *d, const double *du, const double *b,
                                     double *x, int64_t M, int64_t nx, int variant, double *work,
                                     int32_t *singular, void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 1 && dl && d && du && b && x, "bad argument");
    FBR_REQUIRE(variant >= FBR_SOLVER_AUTO && variant <= FBR_SOLVER_PIVOT, "unknown solver variant %d", variant);
    if (variant == FBR_SOLVER_PIVOT) {
        FBR_REQUIRE(work, "the PIVOT variant needs a work array of fbr_solve_tridiag_work_bytes()");
        gtsv_pivot_kernel<<<(unsigned)((M + 63) / 64), 64, 0, as_stream(stream)>>>(dl, d, du, b, x, M, nx, work, singular);
        return check_launch("fbr_solve_tridiag_f64 (pivoted)");
    }
    if (variant == FBR_SOLVER_AUTO)
        // measured on B200 (scratch/k2_bench.py, profiles/r01_k2_bare_solve.txt): the register-resident variant
        // wins at every shape it covers; beyond 4096 cells the tile-transposed Thomas walk
        variant = nx <= kSolveRegMaxNx ? FBR_SOLVER_REG : FBR_SOLVER_THOMAS;
    if (variant == FBR_SOLVER_REG) return solve_reg_dispatch(dl, d, du, b, x, M, nx, singular, as_stream(stream));
    if (variant == FBR_SOLVER_PCR) {
        if (nx > kPcrMaxNx)
            return fail(FBR_ERR_UNSUPPORTED, "PCR variant covers nx <= %d (got %lld)", kPcrMaxNx, (long long)nx);
        int threads = (int)((nx + 31) / 32) * 32;
        if (threads > 1024) threads = 1024;
        const int cells = (int)((nx + threads - 1) / threads);
        const size_t smem = 4 * (size_t)nx * sizeof(double);
        auto go = [&](auto kern) -> int {
            if (smem > 48 * 1024)
                FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)M, threads, smem, as_stream(stream)>>>(dl, d, du, b, x, nx, singular);
            return check_launch("fbr_solve_tridiag_f64 (PCR)");
        };
        switch (cells) {
            case 1: return go(pcr_kernel<1>);
            case 2: return go(pcr_kernel<2>);
            default: return go(pcr_kernel<4>);
        }
    }
    FBR_REQUIRE(work, "the THOMAS variant needs a work array of fbr_solve_tridiag_work_bytes()");
    if (M >= 64 && !getenv("FBR_THOMAS_PLAIN")) {  // many systems: coalesced through transposing tiles
        const int64_t groups = (M + 31) / 32;
        const int warps = groups >= 4 * 148 ? kThWarps : 1;
        const size_t smem = (size_t)warps * 4 * kThCells * 33 * sizeof(double);
        FBR_CUDA(cudaFuncSetAttribute(thomas_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)kThWarps * 4 * kThCells * 33 * sizeof(double))));
        thomas_tiled_kernel<<<(unsigned)((groups + warps - 1) / warps), 32 * warps, smem, as_stream(stream)>>>(
            dl, d, du, b, x, M, nx, work, singular);
        return check_launch("fbr_solve_tridiag_f64 (Thomas, tiled)");
    }
    const int block = M >= 128 * 148 ? 128 : 32;
    thomas_kernel<<<(unsigned)((M + block - 1) / block), block, 0, as_stream(stream)>>>(dl, d, du, b, x, M, nx, work,
                                                                                       singular);
    return check_launch("fbr_solve_tridiag_f64 (Thomas)");
}

// One implicit step, batched: p_next = A^-1 rhs(p), rhs = p with boundary / source overrides
// (matbuilder.py:45-76 + PDESolver_IMP.py:28).  Up to 4096 cells the overrides are applied to the registers of the
// REG solve (one kernel, 40 B per cell-update); longer systems go through fbr_rhs_f64 + the chosen variant.
extern "C" int fbr_step_f64(const double *dl, const double *d, const double *du, const double *p, int64_t M, int64_t nx,
                            int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val,
                            int64_t src_val_stride, double *p_next, int variant, double *work, int32_t *singular,
                            void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 2 && dl && d && du && p && p_next, "bad argument");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "boundary codes must be 0 (None), 1 (Dirichlet) or 2 (Neumann)");
    FBR_REQUIRE(n_src >= 0 && n_src < 0xFE && (n_src == 0 || (src_idx && src_val)), "bad source arguments (n_src=%d)", n_src);
    FBR_REQUIRE(src_val_stride == 0 || src_val_stride >= n_src, "src_val_stride must be 0 (shared values) or >= n_src");
    FBR_REQUIRE(variant >= FBR_SOLVER_AUTO && variant <= FBR_SOLVER_REG, "unknown solver variant %d", variant);
    if (variant == FBR_SOLVER_REG || (variant == FBR_SOLVER_AUTO && nx <= kSolveRegMaxNx))
        return step_reg_dispatch(dl, d, du, p, p_next, M, nx, lbc, rbc, src_idx, n_src, src_val, src_val_stride, singular,
                                 as_stream(stream));
    if (src_val_stride != 0 && M > 1)
        return fail(FBR_ERR_UNSUPPORTED, "per-model source values need the REG variant (nx <= %d)", kSolveRegMaxNx);
    if (int rc = fbr_rhs_f64(p, M, nx, lbc, rbc, src_idx, n_src, src_val, p_next, stream)) return rc;
    return fbr_solve_tridiag_f64(dl, d, du, p_next, p_next, M, nx, variant, work, singular, stream);
}

extern "C" int fbr_jacobi_f64(const double *dl, const double *d, const double *du, const double *b, double *x,
                              int64_t M, int64_t nx, double tol, int max_iter, int32_t *iters, void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 1 && dl && d && du && b && x && max_iter >= 0, "bad argument");
    if (nx > kJacobiMaxNx)
        return fail(FBR_ERR_UNSUPPORTED, "Jacobi mode covers nx <= %d (got %lld)", kJacobiMaxNx, (long long)nx);
    int threads = (int)((nx + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = 2 * (size_t)nx * sizeof(double);
    if (smem > 48 * 1024)
        FBR_CUDA(cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_kernel<<<(unsigned)M, threads, smem, as_stream(stream)>>>(dl, d, du, b, x, nx, tol, max_iter, iters);
    return check_launch("fbr_jacobi_f64");
}
