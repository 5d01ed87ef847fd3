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
#include "fbr_moebius.cuh"
#include "fbr_factor_reg.cuh"
#include "fbr_sweep.cuh"

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
    double *snap;             // (recorded models, capacity, nx)
    double *taxis_out;        // (recorded models, capacity)
    const int64_t *snap_models;  // models whose accepted rows are recorded (NULL: all M)
    int64_t n_snap_models;
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

// slot of model m in the recorded list (-1: its accepted rows are not kept and `capacity` does not stop it)
__device__ __forceinline__ long long record_slot(const AdaptArgs &a, int64_t m) {
    if (!a.snap_models) return (long long)m;
    long long slot = -1;
    for (int64_t q = 0; q < a.n_snap_models; ++q)
        if (a.snap_models[q] == m) slot = q;
    return slot;
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
    const long long rec = record_slot(a, m);
    while (!done && iters < a.max_iters && (rec < 0 || rows < a.capacity)) {
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
            double *row = a.snap + (rec * a.capacity + rows) * a.nx;
#pragma unroll
            for (int k = 0; k < CELLS; ++k) {
                const int i = tid + k * nt;
                p[k] = full[k];
                if (i < n && rec >= 0) row[i] = full[k];
            }
            if (tid == 0 && rec >= 0) a.taxis_out[rec * a.capacity + rows] = t;
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

// ---- register-resident variant -------------------------------------------------------------------
// Same loop, same arithmetic for the matrices (K1's expressions with IEEE division), but the three
// solves of an iteration use the machinery of the run kernels instead of 10 levels of cyclic
// reduction with two divisions per cell and level: a thread owns 8 consecutive cells; A(h) is
// factorised in registers (pivot in front of the chunk from a warp scan of 2x2 Moebius matrices, then
// the plain recurrence with ONE division per cell) and applied with two scan-based sweeps
// (fbr_sweep.cuh); A(dt/2) is factorised once for its two solves.  Per-cell constants and the state
// live in shared memory, the factors and the working vector in registers.
constexpr unsigned kKindInterior = 0, kKindIdentity = 1, kKindNeuL = 2, kKindNeuR = 3, kKindZero = 4;

template <int W>
__global__ void __launch_bounds__(32 * W, (16 / W) > 0 ? (16 / W) : 1) adaptive_reg_kernel(const AdaptArgs a) {
    constexpr int CPT = 8, NT = 32 * W, NXP = NT * CPT;
    extern __shared__ __align__(16) double sm[];
    double *s_em = sm, *s_ep = sm + NXP, *s_dnl = sm + 2 * NXP, *s_dnr = sm + 3 * NXP, *s_p = sm + 4 * NXP;
    double *s_src = sm + 5 * NXP;  // [n_src]
    __shared__ Mat2 s_w[W];
    __shared__ __align__(16) double s_K[2][W][W];
    __shared__ __align__(16) double s_E[2][W];
    __shared__ double s_T[2][W], s_red[2][W], s_clast[W], s_dummy;
    __shared__ int s_bad;
    const int n = (int)a.nx, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cell0 = tid * CPT;
    const int64_t m = blockIdx.x;
    const double *xg = a.mesh + m * a.mesh_stride, *D = a.diff + m * a.diff_stride;
    const unsigned st_f = smem_addr(lane == 31 ? &s_E[0][warp] : &s_dummy);
    const unsigned st_b = smem_addr(lane == 0 ? &s_E[1][warp] : &s_dummy);
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;

    // ---- per-cell constants -> shared memory; row kinds and override slots -> two packed words ----
    unsigned kinds = 0;         // 3 bits per owned cell
    uint64_t slots = ~0ull;     // one byte per owned cell: 0xFF none, 0xFE boundary zero, else source slot
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int i = cell0 + j;
        double em = 0.0, ep = 0.0, dnl = 1.0, dnr = 1.0, pv = 0.0;
        unsigned kind = kKindIdentity;  // padding cells beyond the mesh: identity rows
        if (i < n) {
            pv = a.state[m * a.nx + i];
            if (i > 0 && i < n - 1) {
                kind = kKindInterior;
                const double xm = xg[i - 1], x0 = xg[i], xq = xg[i + 1], Dm = D[i - 1], D0 = D[i], Dq = D[i + 1];
                const double dxm = __dsub_rn(x0, xm), dxp = __dsub_rn(xq, x0), sum = __dadd_rn(dxm, dxp);
                em = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, Dm), D0), __dadd_rn(Dm, D0));
                ep = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, D0), Dq), __dadd_rn(D0, Dq));
                dnl = __ddiv_rn(__dmul_rn(dxm, sum), 2.0);
                dnr = __ddiv_rn(__dmul_rn(dxp, sum), 2.0);
            } else {
                kind = kKindZero;  // 'None' boundary: zero row unless overridden below
            }
            if (i == 0 && a.lbc == FBR_BC_DIRICHLET) kind = kKindIdentity;
            if (i == 0 && a.lbc == FBR_BC_NEUMANN) kind = kKindNeuL;
            if (i == n - 1 && a.rbc == FBR_BC_DIRICHLET) kind = kKindIdentity;
            if (i == n - 1 && a.rbc == FBR_BC_NEUMANN) kind = kKindNeuR;
            if ((i == 0 && a.lbc != FBR_BC_NONE) || (i == n - 1 && a.rbc != FBR_BC_NONE))
                slots = (slots & ~(0xFFull << (8 * j))) | (0xFEull << (8 * j));
            for (int q = 0; q < a.n_src; ++q)
                if (a.src_idx[q] == i) {
                    kind = kKindIdentity;
                    slots = (slots & ~(0xFFull << (8 * j))) | ((uint64_t)q << (8 * j));
                }
        }
        kinds |= kind << (3 * j);
        s_em[i] = em; s_ep[i] = ep; s_dnl[i] = dnl; s_dnr[i] = dnr; s_p[i] = pv;
    }
    double t = a.ctrl[m * 2 + 0], dt = a.ctrl[m * 2 + 1];
    int64_t rows = 0, iters = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    double l[CPT], r[CPT], g[CPT], x[CPT], full[CPT];
    ScanMultipliers smul;

    // A(h) -> factors (l, r, g) in registers + the sweeps' scan multipliers and carry tables
    auto factor = [&](double h) {
        // rows of A(h): a = l[], b = r[], c = g[]  (K1 arithmetic, matbuilder.py:21-76)
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = cell0 + j;
            const unsigned kind = (kinds >> (3 * j)) & 7;
            double aa = 0.0, bb = 0.0, cc = 0.0;
            if (kind == kKindInterior) {
                const double al = __ddiv_rn(__dmul_rn(s_em[i], h), s_dnl[i]), ar = __ddiv_rn(__dmul_rn(s_ep[i], h), s_dnr[i]);
                aa = -al; bb = __dadd_rn(__dadd_rn(1.0, al), ar); cc = -ar;
            } else if (kind == kKindIdentity) { bb = 1.0; }
            else if (kind == kKindNeuL) { bb = -1.0; cc = 1.0; }
            else if (kind == kKindNeuR) { bb = -1.0; aa = 1.0; }
            if (i == 0) aa = 0.0;
            if (i >= n - 1) cc = 0.0;
            l[j] = aa; r[j] = bb; g[j] = cc;
        }
        // pivots from a scan of 2x2 Moebius matrices, independent divisions (fbr_factor_reg.cuh; two CTA barriers)
        const int left = n - cell0;
        const int bad = factor_rows_in_registers<CPT, W>(l, r, g, cell0, left < 0 ? 0 : (left > CPT ? CPT : left), lane, warp,
                                                         s_w, s_clast);
        if (bad) s_bad = bad;
        smul = scan_multipliers<double>(chunk_product<double, CPT>(l), chunk_product<double, CPT>(g), lane);
        if (W > 1) {
            if (lane == 31) s_T[0][warp] = smul.tot_f;
            if (lane == 0) s_T[1][warp] = smul.tot_b;
            __syncthreads();
            if (tid < W) {
                double k = 1.0;
                for (int v = W - 1; v >= 0; --v) {
                    if (v >= tid) s_K[0][tid][v] = 0.0;
                    else { s_K[0][tid][v] = k; k *= s_T[0][v]; }
                }
                k = 1.0;
                for (int v = 0; v < W; ++v) {
                    if (v <= tid) s_K[1][tid][v] = 0.0;
                    else { s_K[1][tid][v] = k; k *= s_T[1][v]; }
                }
            }
            __syncthreads();
        }
    };
    // x <- A^-1 x with the current factors
    auto solve = [&]() {
        {
            double y = chunk_end_fwd_plain<double, CPT>(x, l);
#pragma unroll
            for (int s = 0; s < 5; ++s) y = fma(smul.f[s], __shfl_up_sync(kAll, y, 1 << s), y);
            double c = and_mask(__shfl_up_sync(kAll, y, 1), keep_f);
            if (W > 1) {
                sts_f64(st_f, y);
                __syncthreads();
                double cw = 0.0;
#pragma unroll
                for (int v = 0; v < W - 1; ++v) cw = fma(s_K[0][warp][v], s_E[0][v], cw);
                c = fma(smul.ex_f, cw, c);
            }
            chunk_apply_fwd_plain<double, CPT>(x, l, r, c);
        }
        {
            double z = chunk_end_bwd_plain<double, CPT>(x, g);
#pragma unroll
            for (int s = 0; s < 5; ++s) z = fma(smul.b[s], __shfl_down_sync(kAll, z, 1 << s), z);
            double c = and_mask(__shfl_down_sync(kAll, z, 1), keep_b);
            if (W > 1) {
                sts_f64(st_b, z);
                __syncthreads();
                double cw = 0.0;
#pragma unroll
                for (int v = 1; v < W; ++v) cw = fma(s_K[1][warp][v], s_E[1][v], cw);
                c = fma(smul.ex_b, cw, c);
            }
            chunk_apply_bwd_plain<double, CPT>(x, g, c);
        }
    };
    // right-hand side: b = v with boundary / source overrides (matbuilder.py:45-76)
    auto override_x = [&]() {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const unsigned so = (unsigned)(slots >> (8 * j)) & 0xFF;
            if (so == 0xFE) x[j] = 0.0;
            else if (so != 0xFF) x[j] = s_src[so];
        }
    };

    int done = !(t < a.t_total);
    const long long rec = record_slot(a, m);
    while (!done && iters < a.max_iters && (rec < 0 || rows < a.capacity)) {
        ++iters;
        for (int q = tid; q < a.n_src; q += NT)
            s_src[q] = interp1(a.src_taxis + (int64_t)q * a.src_width, a.src_data + (int64_t)q * a.src_width, a.src_len[q],
                               t - a.t0);
        __syncthreads();
        // pass 0: full = A(dt)^-1 b(p);  pass 1: mid = A(dt/2)^-1 b(p), half = A(dt/2)^-1 b(mid).
        // Written as loops so that the factorisation and the sweeps exist ONCE in the instruction stream
        // (inlined at every use the kernel was 11k instructions and stalled on instruction fetch).
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            factor(pass == 0 ? dt : dt / 2);
#pragma unroll 1
            for (int k = 0; k <= pass; ++k) {
                if (k == 0) {
#pragma unroll
                    for (int j = 0; j < CPT; ++j) x[j] = s_p[cell0 + j];
                }
                override_x();
                solve();
            }
            if (pass == 0) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) full[j] = x[j];
            }
        }
        __syncthreads();
        if (s_bad) break;
        // err = ||full - half|| / ||full||  (fixed-order reduction; padding cells hold zeros)
        double num = 0.0, den = 0.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const double df = full[j] - x[j];
            num += df * df;
            den += full[j] * full[j];
        }
        for (int s = 16; s > 0; s >>= 1) {
            num += __shfl_xor_sync(kAll, num, s);
            den += __shfl_xor_sync(kAll, den, s);
        }
        if (lane == 0) { s_red[0][warp] = num; s_red[1][warp] = den; }
        __syncthreads();
        num = 0.0; den = 0.0;
        for (int w = 0; w < W; ++w) { num += s_red[0][w]; den += s_red[1][w]; }
        const double err = sqrt(num) / sqrt(den);
        // adjust_dt (tso.py:35-50), with Python's min/max semantics for nan
        const double ratio = err < 1e-14 ? 1.0 : pow(a.adj_tol / err, 1.0 / a.p_order);
        double dt_new = dt * a.safety * ratio;
        dt_new = (a.max_dt < dt_new) ? a.max_dt : dt_new;
        dt_new = (dt_new > a.min_dt) ? dt_new : a.min_dt;
        if (err <= a.tol) {  // accept the FULL step (tso.py:24-28)
            t = t + dt;
            double *row = a.snap + (rec * a.capacity + rows) * a.nx;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                s_p[cell0 + j] = full[j];
                if (cell0 + j < n && rec >= 0) row[cell0 + j] = full[j];
            }
            if (tid == 0 && rec >= 0) a.taxis_out[rec * a.capacity + rows] = t;
            ++rows;
            done = !(t < a.t_total);
        }
        dt = dt_new;
        __syncthreads();  // s_red / s_p are reused by the next iteration
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j)
        if (cell0 + j < n) a.state[m * a.nx + cell0 + j] = s_p[cell0 + j];
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
                                    double *taxis_out, const int64_t *snap_models, int64_t n_snap_models,
                                    int64_t *counts, void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 2 && capacity >= 1 && max_iters >= 0, "bad M / nx / capacity / max_iters");
    FBR_REQUIRE(mesh && diffusivity && state && ctrl && counts, "NULL array argument");
    FBR_REQUIRE(snap_models ? n_snap_models >= 0 : n_snap_models == M, "snap_models == NULL needs n_snap_models == M");
    FBR_REQUIRE(n_snap_models == 0 || (snap && taxis_out), "recorded models given without snap / taxis_out");
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
    a.snap_models = snap_models; a.n_snap_models = n_snap_models;
    if (n_src <= 253 && !getenv("FBR_ADAPTIVE_PCR")) {  // register-resident variant (slot bytes hold the source index)
        const int64_t need = (nx + 255) / 256;
        const int W = need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : need <= 8 ? 8 : 16;
        const size_t smem_r = (5 * (size_t)(32 * W * 8) + (size_t)n_src) * sizeof(double);
        auto go_r = [&](auto kern) -> int {
            if (smem_r > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
            kern<<<(unsigned)M, 32 * W, smem_r, as_stream(stream)>>>(a);
            return check_launch("fbr_run_adaptive_f64 (register-resident)");
        };
        switch (W) {
            case 1: return go_r(adaptive_reg_kernel<1>);
            case 2: return go_r(adaptive_reg_kernel<2>);
            case 4: return go_r(adaptive_reg_kernel<4>);
            case 8: return go_r(adaptive_reg_kernel<8>);
            default: return go_r(adaptive_reg_kernel<16>);
        }
    }
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
