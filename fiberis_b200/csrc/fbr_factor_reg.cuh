// Register-resident factorisation of a tridiagonal matrix whose rows are spread over a team of W warps,
// CPT consecutive rows per thread (shared by the bare solve K2 REG, fbr_solve_reg.cu, and the run kernel's
// on-chip assembly mode, fbr_run_onchip.cu).  See fbr_solve_reg.cu for the derivation: pivots from the
// leading-minor recurrence p_i = d_i p_{i-1} + e_i p_{i-2}, e_i = -dl_i du_{i-1}, scanned over the warp as
// 2x2 projective matrices, re-run locally with FMAs only on the chain and INDEPENDENT divisions.
#pragma once
#include "fbr_moebius.cuh"

namespace fbr {

// scale a 2x2 matrix / a pair by the power of two that brings its largest entry to [1, 2): integer
// work on the exponent fields only (the library scalbn / ilogb pair costs ~10x as much)
__device__ __forceinline__ double pow2_scale_for(int hi_max_exp_field) {
    // hi_max_exp_field = (hi word & 0x7ff00000) of the largest magnitude;  2^(1023 - E) has hi word 0x7fe00000 - field
    int s = max(0x7fe00000 - hi_max_exp_field, 0x00100000);  // never below the smallest normal
    if (hi_max_exp_field >= 0x7ff00000) s = 0x3ff00000;      // inf / nan: leave alone
    return __hiloint2double(s, 0);
}
__device__ __forceinline__ int exp_field(double v) { return __double2hiint(v) & 0x7ff00000; }
__device__ __forceinline__ Mat2 rescale_fast(Mat2 m) {
    const double s = pow2_scale_for(max(max(exp_field(m.a), exp_field(m.b)), max(exp_field(m.c), exp_field(m.d))));
    m.a *= s; m.b *= s; m.c *= s; m.d *= s;
    return m;
}
__device__ __forceinline__ Mat2 compose_raw(const Mat2 &later, const Mat2 &earlier) {
    Mat2 r;
    r.a = fma(later.a, earlier.a, later.b * earlier.c);
    r.b = fma(later.a, earlier.b, later.b * earlier.d);
    r.c = fma(later.c, earlier.a, later.d * earlier.c);
    r.d = fma(later.c, earlier.b, later.d * earlier.d);
    return r;
}

// num / den without the library's special-case path: den is a rescaled pivot numerator (normal range, or an
// exact zero -> non-finite quotient, which the caller flags).  Hardware seed, two Newton steps, one correction.
__device__ __forceinline__ double div_fast(double num, double den) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(den));
    double e = fma(-den, x, 1.0);
    x = fma(x, e, x);
    e = fma(-den, x, 1.0);
    x = fma(x, e, x);
    const double q = num * x;
    return fma(fma(-den, q, num), x, q);
}


// In: rows (dl, d, du) of my CPT cells in (l, r, g), identity rows beyond the mesh, dl of row 0 and du of the last row
// already zero.  Out: the un-pivoted LU factors the run kernels use — l_i = dl_i / u_{i-1}, r_i = 1 / u_i,
// g_i = du_i / u_i — and the return value = 1 + first row among my first `nvalid` with a zero / non-finite pivot (0: none).
// cell0 = my first row within the team's tile, first_row = global row of the tile's first cell (0 for a whole system).
// A tile in the middle of a longer system passes TileLink: the super-diagonal entry in front of the tile and the
// (p, q) = (p_{i-1}, p_{i-2}) pair in front of it (from a scan over the tiles); `total`, when non-NULL, receives the
// tile's composite matrix and the function returns right after (first pass of a two-pass factorisation).
// s_w: W matrices, s_c: W doubles of shared scratch; contains two CTA barriers when W > 1 (CTA-uniform call).
struct TileLink {
    double c_front = 0.0, p = 1.0, q = 1.0;
    int64_t first_row = 0;
    Mat2 *total = nullptr;
};
template <int CPT, int W>
__device__ __forceinline__ int factor_rows_in_registers(double (&l)[CPT], double (&r)[CPT], double (&g)[CPT], int cell0,
                                                        int nvalid, int lane, int warp, Mat2 *s_w, double *s_c,
                                                        const TileLink &link = TileLink()) {
    // super-diagonal entry of the row in front of my chunk
    double c_prev = __shfl_up_sync(kAll, g[CPT - 1], 1);
    if (W > 1) {
        if (lane == 31) s_c[warp] = g[CPT - 1];
        __syncthreads();
        if (lane == 0) c_prev = warp > 0 ? s_c[warp - 1] : link.c_front;
    } else if (lane == 0) {
        c_prev = link.c_front;
    }
    double ee[CPT];
    {
        double cp = c_prev;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            ee[j] = -l[j] * cp;
            cp = g[j];
        }
    }
    Mat2 G{r[0], ee[0], 1.0, 0.0};
#pragma unroll
    for (int j = 1; j < CPT; ++j) {
        const Mat2 nm{fma(r[j], G.a, ee[j] * G.c), fma(r[j], G.b, ee[j] * G.d), G.a, G.b};
        G = (j == CPT / 2 - 1) ? rescale_fast(nm) : nm;
    }
    Mat2 inc = rescale_fast(G);
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const Mat2 up = shfl_up_mat(inc, s);
        if (lane >= s) inc = compose_raw(inc, up);
        if (s == 4) inc = rescale_fast(inc);
    }
    inc = rescale_fast(inc);
    Mat2 ex = shfl_up_mat(inc, 1);
    if (lane == 0) ex = Mat2{1.0, 0.0, 0.0, 1.0};
    if (W > 1) {
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        Mat2 before{1.0, 0.0, 0.0, 1.0};
        if (link.total) {  // first pass: only the composite of the whole tile is wanted
            if (warp == 0 && lane == 0) {
                for (int q = 0; q < W; ++q) before = rescale_fast(compose_raw(s_w[q], before));
                *link.total = before;
            }
            return 0;
        }
        for (int q = 0; q < warp; ++q) before = rescale_fast(compose_raw(s_w[q], before));
        ex = compose_raw(ex, before);
    } else if (link.total) {
        if (lane == 31) *link.total = inc;
        return 0;
    }
    double p1 = fma(ex.a, link.p, ex.b * link.q), p2 = fma(ex.c, link.p, ex.d * link.q);  // p_{i-1}, p_{i-2} in front of my chunk
    {
        const double s = pow2_scale_for(max(exp_field(p1), exp_field(p2)));
        p1 *= s; p2 *= s;
    }
    double r_prev = div_fast(p2, p1);  // 1 / u of the row in front of my chunk (unused by row 0: dl_0 = 0)
    if (cell0 == 0 && link.first_row == 0) r_prev = 0.0;
    double nanacc = 0.0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const double p0 = fma(r[j], p1, ee[j] * p2);
        r[j] = div_fast(p1, p0);
        nanacc = fma(r[j], 0.0, fma(p0, 0.0, nanacc));  // NaN <=> a zero pivot (r = inf) or an overflowed minor (p0 = inf)
        p2 = p1; p1 = p0;
        if (j == CPT / 2 - 1) {
            const double s = pow2_scale_for(max(exp_field(p1), exp_field(p2)));
            p1 *= s; p2 *= s;
        }
    }
    int bad = 0;
    if (nanacc != nanacc) {
#pragma unroll
        for (int j = CPT - 1; j >= 0; --j)
            if (j < nvalid && (!isfinite(r[j]) || r[j] == 0.0)) bad = (int)(link.first_row + cell0) + j + 1;
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        l[j] *= r_prev;
        g[j] *= r[j];
        r_prev = r[j];
    }
    return bad;
}

}  // namespace fbr
