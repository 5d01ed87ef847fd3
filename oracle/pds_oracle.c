/*
 * CPU oracle (plain C) for the fibeRIS 1D pressure-diffusion hot path.
 *
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Built by oracle/Makefile into oracle/_build/ and loaded
 * (ctypes) only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs, as the checker and as the timed CPU baseline.  fiberis_b200 never links or loads it.
 *
 * What it restates (paths relative to /root/reference/):
 *   fbo_assemble ........ src/fiberis/simulator/solver/matbuilder.py:21-28,39-81,103-165,170-196
 *                         (same operation order as the NumPy expressions; build with
 *                         -ffp-contract=off so the diagonals are bit-identical)
 *   fbo_gttrf/fbo_gttrs . the linear solve of src/fiberis/simulator/solver/PDESolver_IMP.py:27-28:
 *                         scipy.linalg.solve -> LAPACK.  SciPy (pinned 1.13.1, poetry.lock:898-899;
 *                         declared >=1.8,<2.0, pyproject.toml:46) is not vendored in the reference,
 *                         so this is a restatement of the published LAPACK DGTTRF/DGTTRS algorithm
 *                         (Gaussian elimination with partial pivoting on a tridiagonal matrix, the
 *                         routine pair behind DGTSV that SciPy >= 1.15 dispatches to).
 *   fbo_run ............. the fixed-dt loop of src/fiberis/simulator/core/pds.py:287-310 with the
 *                         right-hand side of matbuilder.py:45-76; A is factorised once because it
 *                         does not change while dt is fixed.
 *   fbo_run_ensemble .... fbo_run over independent members (OpenMP), with the misfit
 *                         sum_n sum_k w_k (p^n[idx_k] - obs[n,k])^2 (parity unpinned: the reference
 *                         defines no misfit in fiberis.simulator).
 * Pinned by tests/test_oracle_c.py against oracle/pds_oracle.py, which in turn is pinned against
 * the reference's golden vectors and reference-generated fixtures (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { BC_NONE = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2 };

int fbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double pml_sigma(const double *x, int64_t i, int64_t nx, double L, double smax) {
    if (L == 0.0) return 0.0;
    double dl = x[i] - x[0], dr = x[nx - 1] - x[i], s = 0.0;
    if (dl < L) s = smax * ((L - dl) / L);
    if (dr < L) {
        double t = smax * ((L - dr) / L);
        if (t > s) s = t;
    }
    return s;
}

/* dl, d, du: length nx each; dl[0] = du[nx-1] = 0 */
void fbo_assemble(const double *x, const double *D, int64_t nx, double dt, int lbc, int rbc,
                  const int64_t *src_idx, int n_src, double pml, double smax, double *dl, double *d, double *du) {
    for (int64_t i = 0; i < nx; ++i) {
        double a = 0, b = 0, c = 0;
        if (i > 0 && i < nx - 1) {
            double dxm = x[i] - x[i - 1], dxp = x[i + 1] - x[i];
            double dem = 2 * D[i - 1] * D[i] / (D[i - 1] + D[i]);
            double dep = 2 * D[i] * D[i + 1] / (D[i] + D[i + 1]);
            double al = dem * dt / (dxm * (dxm + dxp) / 2);
            double ar = dep * dt / (dxp * (dxm + dxp) / 2);
            a = -al;
            b = 1 + al + ar;
            c = -ar;
        }
        if (i == 0) {
            if (lbc == BC_DIRICHLET) b = 1;
            else if (lbc == BC_NEUMANN) { b = -1; c = 1; }
        }
        if (i == nx - 1) {
            if (rbc == BC_DIRICHLET) b = 1;
            else if (rbc == BC_NEUMANN) { b = -1; a = 1; }
        }
        for (int k = 0; k < n_src; ++k)
            if (src_idx[k] == i) { a = 0; b = 1; c = 0; }
        b = b + dt * pml_sigma(x, i, nx, pml, smax);
        dl[i] = i == 0 ? 0 : a;
        d[i] = b;
        du[i] = i == nx - 1 ? 0 : c;
    }
}

/* LAPACK-style storage: dl[0..n-2] sub, d[0..n-1], du[0..n-2] super, du2[0..n-3], ipiv[0..n-1].
 * Returns 0 or 1 + index of the first zero pivot. */
int64_t fbo_gttrf(int64_t n, double *dl, double *d, double *du, double *du2, int64_t *ipiv) {
    for (int64_t i = 0; i < n; ++i) ipiv[i] = i;
    for (int64_t i = 0; i + 2 < n; ++i) du2[i] = 0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        if (fabs(d[i]) >= fabs(dl[i])) {
            if (d[i] != 0) {
                double f = dl[i] / d[i];
                dl[i] = f;
                d[i + 1] -= f * du[i];
            }
        } else {
            double f = d[i] / dl[i];
            d[i] = dl[i];
            dl[i] = f;
            double t = du[i];
            du[i] = d[i + 1];
            d[i + 1] = t - f * d[i + 1];
            if (i + 2 < n) {
                du2[i] = du[i + 1];
                du[i + 1] = -f * du[i + 1];
            }
            ipiv[i] = i + 1;
        }
    }
    for (int64_t i = 0; i < n; ++i)
        if (d[i] == 0) return i + 1;
    return 0;
}

void fbo_gttrs(int64_t n, const double *dl, const double *d, const double *du, const double *du2,
               const int64_t *ipiv, double *b) {
    for (int64_t i = 0; i + 1 < n; ++i) {
        if (ipiv[i] == i) {
            b[i + 1] -= dl[i] * b[i];
        } else {
            double t = b[i];
            b[i] = b[i + 1];
            b[i + 1] = t - dl[i] * b[i];
        }
    }
    b[n - 1] /= d[n - 1];
    if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
    for (int64_t i = n - 3; i >= 0; --i) b[i] = (b[i] - du[i] * b[i + 1] - du2[i] * b[i + 2]) / d[i];
}

/* One model, fixed dt.  snap: NULL or (n_steps / every, nx); p_final: NULL or (nx); misfit: NULL or 1.
 * Returns 0, or 1 + zero-pivot row. */
int64_t fbo_run(const double *x, const double *D, const double *p0, int64_t nx, int64_t n_steps, double dt, int lbc,
                int rbc, const int64_t *src_idx, int n_src, const double *src_table, int64_t every, double *snap,
                double *p_final, const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                double *misfit) {
    double *buf = (double *)malloc(sizeof(double) * (size_t)nx * 8);
    int64_t *ipiv = (int64_t *)malloc(sizeof(int64_t) * (size_t)nx);
    double *dl = buf, *d = buf + nx, *du = buf + 2 * nx, *du2 = buf + 3 * nx, *p = buf + 4 * nx;
    fbo_assemble(x, D, nx, dt, lbc, rbc, src_idx, n_src, 0.0, 0.0, dl, d, du);
    memmove(dl, dl + 1, sizeof(double) * (size_t)(nx - 1)); /* to LAPACK storage */
    int64_t info = fbo_gttrf(nx, dl, d, du, du2, ipiv);
    double acc = 0.0;
    if (!info) {
        memcpy(p, p0, sizeof(double) * (size_t)nx);
        for (int64_t n = 0; n < n_steps; ++n) {
            if (lbc != BC_NONE) p[0] = 0;
            if (rbc != BC_NONE) p[nx - 1] = 0;
            for (int k = 0; k < n_src; ++k) p[src_idx[k]] = src_table[n * n_src + k];
            fbo_gttrs(nx, dl, d, du, du2, ipiv, p);
            for (int k = 0; k < n_obs; ++k) {
                double r = p[obs_idx[k]] - obs[(n + 1) * n_obs + k];
                acc += (obs_w ? obs_w[k] : 1.0) * r * r;
            }
            if (snap && every > 0 && (n + 1) % every == 0)
                memcpy(snap + ((n + 1) / every - 1) * nx, p, sizeof(double) * (size_t)nx);
        }
        if (p_final) memcpy(p_final, p, sizeof(double) * (size_t)nx);
        if (misfit) *misfit = acc;
    }
    free(buf);
    free(ipiv);
    return info;
}

/* M members sharing mesh/sources/BCs.  D: (M, nx) dense when zone_of_cell == NULL, else zone values (M, n_zones).
 * p0: shared (nx).  misfit: (M).  final: NULL or (M, nx).  Returns the number of singular members. */
int64_t fbo_run_ensemble(const double *x, const double *D, int n_zones, const int32_t *zone_of_cell, const double *p0,
                         int64_t M, int64_t nx, int64_t n_steps, double dt, int lbc, int rbc,
                         const int64_t *src_idx, int n_src, const double *src_table, const int64_t *obs_idx,
                         int n_obs, const double *obs, const double *obs_w, double *misfit, double *final,
                         int threads) {
    int64_t bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int64_t m = 0; m < M; ++m) {
        double *Dm = (double *)malloc(sizeof(double) * (size_t)nx);
        if (zone_of_cell)
            for (int64_t i = 0; i < nx; ++i) Dm[i] = D[m * n_zones + zone_of_cell[i]];
        else
            memcpy(Dm, D + m * nx, sizeof(double) * (size_t)nx);
        double mf = 0.0;
        int64_t info = fbo_run(x, Dm, p0, nx, n_steps, dt, lbc, rbc, src_idx, n_src, src_table, 0, NULL,
                               final ? final + m * nx : NULL, obs_idx, n_obs, obs, obs_w, &mf);
        if (misfit) misfit[m] = mf;
        bad += info != 0;
        free(Dm);
    }
    return bad;
}
