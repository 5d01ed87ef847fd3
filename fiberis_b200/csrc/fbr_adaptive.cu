// Device-side adaptive time stepping: the whole step-doubling loop of PDS1D_*.solve(optimizer=True)
// (src/fiberis/simulator/core/pds.py:312-336) and its controller (src/fiberis/simulator/optimizer/
// tso.py:4-50) in ONE persistent kernel, one CTA per model.  Per iteration the CTA
//   samples the sources at t_n - t0 (np.interp semantics, src/fiberis/analyzer/Data1D/core1D.py:247),
//   assembles A(dt) and A(dt/2) from per-cell coefficients it keeps in registers (K1 arithmetic,
//   bit-identical: alpha = (Deff * dt) / den with IEEE division),
//   solves full / half / half with parallel cyclic reduction in shared memory (K2, no factorisation
//   because A changes every iteration),
//   reduces err = ||full - half||_2 / ||full||_2 (K5), accepts the FULL step when err <= tol and
//   computes the next dt exactly as adjust_dt does (including its nan -> min_dt behaviour).
// Accepted states and times go to a caller-provided ring of `capacity` rows; the kernel stops when
// the ring is full, t >= t_total or max_iters is reached and leaves (t, dt, state) for a relaunch.
#include "fbr_common.cuh"

namespace fbr {

struct AdaptArgs {
    const double *mesh; int64_t mesh_stride;
    const double *diff; int64_t diff_stride;
    double *state;            // (M, nx) in / out
    int64_t M, nx;
    int lbc, rbc;
    const int64_t *src_idx; int n_src;
    const double *src_taxis, *src_data;  // (n_src, src_width)
    const int32_t *src_len; int src_width;
    double t0, t_total;
    double *ctrl;             // (M, 2): t, dt  in / out
    double tol, adj_tol, safety, p_order, max_dt, min_dt;
    int64_t max_iters, capacity;
    double *snap;             // (M, capacity, nx)
    double *taxis_out;        // (M, capacity)
    int64_t *counts;          // (M, 4): rows, iterations, done, singular row + 1
};

// np.interp for one abscissa (numpy/core/src/multiarray/compiled_base.c: arr_interp), no FMA contraction
__device__ double interp1(const double *__restrict__ xp, const double *__restrict__ fp, int n, double x) {
    if (n == 1) return fp[0];
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    if (x != x) return x;
    int lo = 0, hi = n - 1;  // invariant xp[lo] <= x, and x < xp[hi] unless x == xp[n-1]
    if (x == xp[n - 1]) return fp[n - 1];
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    if (xp[lo] == x) return fp[lo];
    const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo])), fp[lo]);
    if (r != r) {
        r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo + 1])), fp[lo + 1]);
        if (r != r && fp[lo] == fp[lo + 1]) r = fp[lo];
    }
    return r;
}

template <int CELLS>
__global__ void __launch_bounds__(1024) adaptive_kernel(const AdaptArgs a) {
    extern __shared__ double sm[];
    const int n = (int)a.nx, tid = threadIdx.x, nt = blockDim.x;
    double *sa = sm, *sb = sm + n, *sc = sm + 2 * n, *sr = sm + 3 * n;
    double *s_src = sm + 4 * n;  // [n_src]
    __shared__ double s_red[2][32];
    __shared__ int s_bad;
    const int64_t m = blockIdx.x;
    const double *x = a.mesh + m * a.mesh_stride, *D = a.diff + m * a.diff_stride;

    // ---- per-cell constants (registers) ----
    double em[CELLS], ep[CELLS], dnl[CELLS], dnr[CELLS], p[CELLS], full[CELLS], half[CELLS];
    int kind[CELLS];  // 0 interior, 1 = identity-type row (diag fixed), 2 Neumann-left, 3 Neumann-right, 4 zero row
    int slot[CELLS];  // source slot or -1; -2 = boundary zero override
#pragma unroll
    for (int k = 0; k < CELLS; ++k) {
        const int i = tid + k * nt;
        em[k] = ep[k] = 0.0; dnl[k] = dnr[k] = 1.0; p[k] = 0.0; kind[k] = 0; slot[k] = -1;
        if (i < n) {
            p[k] = a.state[m * a.nx + i];
            if (i > 0 && i < n - 1) {
                const double xm = x[i - 1], x0 = x[i], xq = x[i + 1], Dm = D[i - 1], D0 = D[i], Dq = D[i + 1];
                const double dxm = __dsub_rn(x0, xm), dxp = __dsub_rn(xq, x0), sum = __dadd_rn(dxm, dxp);
                em[k] = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, Dm), D0), __dadd_rn(Dm, D0));
                ep[k] = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, D0), Dq), __dadd_rn(D0, Dq));
                dnl[k] = __ddiv_rn(__dmul_rn(dxm, sum), 2.0);
                dnr[k] = __ddiv_rn(__dmul_rn(dxp, sum), 2.0);
            } else {
                kind[k] = 4;  // 'None' boundary: zero row unless overridden below
            }
            if (i == 0 && a.lbc == FBR_BC_DIRICHLET) kind[k] = 1;
            if (i == 0 && a.lbc == FBR_BC_NEUMANN) kind[k] = 2;
            if (i == n - 1 && a.rbc == FBR_BC_DIRICHLET) kind[k] = 1;
            if (i == n - 1 && a.rbc == FBR_BC_NEUMANN) kind[k] = 3;
            if ((i == 0 && a.lbc != FBR_BC_NONE) || (i == n - 1 && a.rbc != FBR_BC_NONE)) slot[k] = -2;
            for (int q = 0; q < a.n_src; ++q)
                if (a.src_idx[q] == i) { kind[k] = 1; slot[k] = q; }
        }
    }
    double t = a.ctrl[m * 2 + 0], dt = a.ctrl[m * 2 + 1];
    int64_t rows = 0, iters = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    // PCR solve of A(h) x = rhs; rhs in registers, result in out[]
    auto solve = [&](double h, const double (&rhs)[CELLS], double (&out)[CELLS]) {
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const int i = tid + k * nt;
            if (i < n) {
                double aa = 0.0, bb = 0.0, cc = 0.0;
                if (kind[k] == 0) {
                    const double al = __ddiv_rn(__dmul_rn(em[k], h), dnl[k]), ar = __ddiv_rn(__dmul_rn(ep[k], h), dnr[k]);
                    aa = -al; bb = __dadd_rn(__dadd_rn(1.0, al), ar); cc = -ar;
                } else if (kind[k] == 1) { bb = 1.0; }
                else if (kind[k] == 2) { bb = -1.0; cc = 1.0; }
                else if (kind[k] == 3) { bb = -1.0; aa = 1.0; }
                if (i == 0) aa = 0.0;
                if (i == n - 1) cc = 0.0;
                sa[i] = aa; sb[i] = bb; sc[i] = cc; sr[i] = rhs[k];
            }
        }
        __syncthreads();
        for (int s = 1; s < n; s <<= 1) {
            double na[CELLS], nb[CELLS], nc[CELLS], nr[CELLS];
#pragma unroll
            for (int k = 0; k < CELLS; ++k) {
                const int i = tid + k * nt;
                if (i < n) {
                    const int im = i - s, ip = i + s;
                    double av = sa[i], bv = sb[i], cv = sc[i], rv = sr[i], xa = 0.0, xc = 0.0;
                    if (im >= 0) {
                        const double piv = sb[im];
                        if (piv == 0.0) s_bad = im + 1;
                        const double k1 = av / piv;
                        xa = -sa[im] * k1; bv -= sc[im] * k1; rv -= sr[im] * k1;
                    }
                    if (ip < n) {
                        const double piv = sb[ip];
                        if (piv == 0.0) s_bad = ip + 1;
                        const double k2 = cv / piv;
                        xc = -sc[ip] * k2; bv -= sa[ip] * k2; rv -= sr[ip] * k2;
                    }
                    na[k] = xa; nb[k] = bv; nc[k] = xc; nr[k] = rv;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < CELLS; ++k) {
                const int i = tid + k * nt;
                if (i < n) { sa[i] = na[k]; sb[i] = nb[k]; sc[i] = nc[k]; sr[i] = nr[k]; }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const int i = tid + k * nt;
            out[k] = 0.0;
            if (i < n) {
                const double piv = sb[i];
                if (piv == 0.0 || !isfinite(piv)) s_bad = i + 1;
                out[k] = sr[i] / piv;
            }
        }
        __syncthreads();
    };
    auto override_rhs = [&](const double (&src)[CELLS], double (&dst)[CELLS]) {
#pragma unroll
        for (int k = 0; k < CELLS; ++k) dst[k] = slot[k] == -1 ? src[k] : (slot[k] == -2 ? 0.0 : s_src[slot[k]]);
    };

    int done = !(t < a.t_total);
    while (!done && iters < a.max_iters && rows < a.capacity) {
        ++iters;
        for (int q = tid; q < a.n_src; q += nt)
            s_src[q] = interp1(a.src_taxis + (int64_t)q * a.src_width, a.src_data + (int64_t)q * a.src_width, a.src_len[q],
                               t - a.t0);
        __syncthreads();
        double rhs[CELLS], mid[CELLS];
        override_rhs(p, rhs);
        solve(dt, rhs, full);
        solve(dt / 2, rhs, mid);
        override_rhs(mid, rhs);
        solve(dt / 2, rhs, half);
        if (s_bad) break;
        // err = ||full - half|| / ||full||  (fixed-order reduction)
        double num = 0.0, den = 0.0;
#pragma unroll
        for (int k = 0; k < CELLS; ++k) {
            const double df = full[k] - half[k];
            num += df * df;
            den += full[k] * full[k];
        }
        for (int s = 16; s > 0; s >>= 1) {
            num += __shfl_xor_sync(0xffffffffu, num, s);
            den += __shfl_xor_sync(0xffffffffu, den, s);
        }
        if ((tid & 31) == 0) { s_red[0][tid >> 5] = num; s_red[1][tid >> 5] = den; }
        __syncthreads();
        num = 0.0; den = 0.0;
        for (int w = 0; w < (nt + 31) / 32; ++w) { num += s_red[0][w]; den += s_red[1][w]; }
        __syncthreads();
        const double err = sqrt(num) / sqrt(den);
        // adjust_dt (tso.py:35-50), with Python's min/max semantics for nan
        const double ratio = err < 1e-14 ? 1.0 : pow(a.adj_tol / err, 1.0 / a.p_order);
        double dt_new = dt * a.safety * ratio;
        dt_new = (a.max_dt < dt_new) ? a.max_dt : dt_new;
        dt_new = (dt_new > a.min_dt) ? dt_new : a.min_dt;
        if (err <= a.tol) {  // accept the FULL step (tso.py:24-28)
            t = t + dt;
            double *row = a.snap + (m * a.capacity + rows) * a.nx;
#pragma unroll
            for (int k = 0; k < CELLS; ++k) {
                const int i = tid + k * nt;
                p[k] = full[k];
                if (i < n) row[i] = full[k];
            }
            if (tid == 0) a.taxis_out[m * a.capacity + rows] = t;
            ++rows;
            done = !(t < a.t_total);
        }
        dt = dt_new;
    }
#pragma unroll
    for (int k = 0; k < CELLS; ++k) {
        const int i = tid + k * nt;
        if (i < n) a.state[m * a.nx + i] = p[k];
    }
    if (tid == 0) {
        a.ctrl[m * 2 + 0] = t;
        a.ctrl[m * 2 + 1] = dt;
        a.counts[m * 4 + 0] = rows;
        a.counts[m * 4 + 1] = iters;
        a.counts[m * 4 + 2] = done;
        a.counts[m * 4 + 3] = s_bad;
    }
}

}  // namespace fbr

using namespace fbr;

extern "C" int64_t fbr_run_adaptive_max_nx(void) { return 4096; }

extern "C" int fbr_run_adaptive_f64(const double *mesh, int64_t mesh_stride, const double *diffusivity,
                                    int64_t diff_stride, double *state, int64_t M, int64_t nx, int lbc, int rbc,
                                    const int64_t *src_idx, int n_src, const double *src_taxis, const double *src_data,
                                    const int32_t *src_len, int src_width, double t0, double t_total, double *ctrl,
                                    double tol, double adj_tol, double safety, double p_order, double max_dt,
                                    double min_dt, int64_t max_iters, int64_t capacity, double *snap,
                                    double *taxis_out, int64_t *counts, void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 2 && capacity >= 1 && max_iters >= 0, "bad M / nx / capacity / max_iters");
    FBR_REQUIRE(mesh && diffusivity && state && ctrl && snap && taxis_out && counts, "NULL array argument");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_taxis && src_data && src_len && src_width >= 1)),
                "sources given without their series");
    if (nx > fbr_run_adaptive_max_nx())
        return fail(FBR_ERR_UNSUPPORTED, "device-side adaptive stepping covers nx <= %lld", (long long)fbr_run_adaptive_max_nx());
    AdaptArgs a;
    a.mesh = mesh; a.mesh_stride = mesh_stride; a.diff = diffusivity; a.diff_stride = diff_stride; a.state = state;
    a.M = M; a.nx = nx; a.lbc = lbc; a.rbc = rbc; a.src_idx = src_idx; a.n_src = n_src; a.src_taxis = src_taxis;
    a.src_data = src_data; a.src_len = src_len; a.src_width = src_width; a.t0 = t0; a.t_total = t_total; a.ctrl = ctrl;
    a.tol = tol; a.adj_tol = adj_tol; a.safety = safety; a.p_order = p_order; a.max_dt = max_dt; a.min_dt = min_dt;
    a.max_iters = max_iters; a.capacity = capacity; a.snap = snap; a.taxis_out = taxis_out; a.counts = counts;
    int threads = (int)((nx + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const int cells = (int)((nx + threads - 1) / threads);
    const size_t smem = (4 * (size_t)nx + (size_t)n_src) * sizeof(double);
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)M, threads, smem, as_stream(stream)>>>(a);
        return check_launch("fbr_run_adaptive_f64");
    };
    switch (cells) {
        case 1: return go(adaptive_kernel<1>);
        case 2: return go(adaptive_kernel<2>);
        default: return go(adaptive_kernel<4>);
    }
}
