// K2 REG  bare batched tridiagonal solve, register-resident: one pass over HBM (40 B per unknown:
// dl, d, du, b in, x out), no scratch array, no shared-memory copy of the system.  (Staging the next system
// of a persistent CTA through shared memory with cp.async was measured and dropped: the kernel's shuffles
// already load the MIO queue, the extra LDGSTS / LDS traffic cost more than the hidden latency gave back —
// 65536 x 512: 0.40 ms staged against 0.31 ms with plain 16-byte loads.  16 cells per thread — half the
// shuffles per cell, but 225 registers — was slower as well: 0.38 ms; prefetch.global.L2 of the CTA's next system: no
// change (0.276 vs 0.274 ms); 20 warps per SM at 96 registers with spills: 0.29 ms.)
//
// A thread owns CPT consecutive cells of one system (64 bytes per array, 16-byte loads), W warps own a
// system.  The elimination is made parallel along the mesh the same way as in the adaptive loop
// (fbr_adaptive.cu) and the run kernels (fbr_sweep.cuh):
//   1. pivots: u_i = d_i - dl_i du_{i-1} / u_{i-1} is a Moebius map of u_{i-1}.  With u_i = p_i / p_{i-1}
//      it is the linear recurrence p_i = d_i p_{i-1} + e_i p_{i-2}, e_i = -dl_i du_{i-1} (the p_i are the
//      leading principal minors up to a power-of-two scale).  A thread composes the 2x2 matrix of its
//      chunk, the warp scans the matrices (fbr_moebius.cuh), warps exchange their totals through shared
//      memory; the thread then runs the two-term recurrence from the true (p, q) in front of its chunk —
//      FMAs only on the dependent chain — and takes r_i = p_{i-1} / p_i with CPT INDEPENDENT divisions
//      (the sequential algorithm has one division per cell ON the chain).
//   2. Thomas in its normalised form: z_i = r_i b_i - (r_i dl_i) z_{i-1};  x_i = z_i - (r_i du_i) x_{i+1}:
//      two first-order linear recurrences = two scan sweeps (local pass, Kogge-Stone over the warp with the
//      multipliers' running products, carry over the warps, second local pass).
// Rows beyond nx are padded with identity rows.  Zero / non-finite pivots are flagged per system
// (1 + first such row among the rows a thread saw; 0 = fine), like the other K2 variants.
#include "fbr_common.cuh"
#include "fbr_moebius.cuh"
#include "fbr_factor_reg.cuh"
#include "fbr_sweep.cuh"

namespace fbr {

template <int CPT>
__device__ __forceinline__ void load_chunk(const double *__restrict__ row, int nvalid, bool vec, double pad, double (&v)[CPT],
                                           bool odd = false) {
    if (vec) {
        const double2 *q = reinterpret_cast<const double2 *>(row);
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j) {
            const double2 t = __ldg(q + j);
            v[2 * j] = t.x;
            v[2 * j + 1] = t.y;
        }
    } else if (odd) {
        // a full chunk that starts on an ODD element (odd nx: every other system's row): one 8-byte load, CPT / 2 - 1
        // 16-byte loads, one 8-byte load instead of CPT scalar ones (65536 x 511: 3.3 TB/s with scalar loads, 5.0 at 512)
        v[0] = __ldg(row);
        const double2 *q = reinterpret_cast<const double2 *>(row + 1);
#pragma unroll
        for (int j = 0; j < CPT / 2 - 1; ++j) {
            const double2 t = __ldg(q + j);
            v[2 * j + 1] = t.x;
            v[2 * j + 2] = t.y;
        }
        v[CPT - 1] = __ldg(row + CPT - 1);
    } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = j < nvalid ? __ldg(row + j) : pad;
    }
}
// the sweeps' stage multipliers for groups of L lanes (fbr_sweep.cuh's scan_multipliers is the L = 32 case)
template <int L>
struct GroupMultipliers {
    static constexpr int NS = L == 32 ? 5 : L == 16 ? 4 : L == 8 ? 3 : L == 4 ? 2 : 1;
    double f[NS], b[NS], ex_f, ex_b, tot_f, tot_b;
};
template <int L>
__device__ __forceinline__ GroupMultipliers<L> group_multipliers(double chunk_f, double chunk_b, int sub) {
    GroupMultipliers<L> m;
    double pf = chunk_f, pb = chunk_b;
#pragma unroll
    for (int s = 0; s < GroupMultipliers<L>::NS; ++s) {
        const bool fa = sub >= (1 << s), ba = sub + (1 << s) < L;
        m.f[s] = fa ? pf : 0.0;
        m.b[s] = ba ? pb : 0.0;
        const double upf = __shfl_up_sync(kAll, pf, 1 << s, L);
        const double dnb = __shfl_down_sync(kAll, pb, 1 << s, L);
        if (fa) pf *= upf;
        if (ba) pb *= dnb;
    }
    m.ex_f = __shfl_up_sync(kAll, pf, 1, L);
    m.ex_b = __shfl_down_sync(kAll, pb, 1, L);
    if (sub == 0) m.ex_f = 1.0;
    if (sub == L - 1) m.ex_b = 1.0;
    m.tot_f = pf;
    m.tot_b = pb;
    return m;
}
template <int L>
__device__ __forceinline__ Mat2 shfl_up_mat_w(const Mat2 &m, int delta) {
    Mat2 r;
    r.a = __shfl_up_sync(kAll, m.a, delta, L); r.b = __shfl_up_sync(kAll, m.b, delta, L);
    r.c = __shfl_up_sync(kAll, m.c, delta, L); r.d = __shfl_up_sync(kAll, m.d, delta, L);
    return r;
}

// fbr_step_f64: the right-hand side is the current pressure with boundary / source overrides
// (matbuilder.py:45-76) applied while it sits in registers; enabled == 0 for the bare solve.
struct StepOverrides {
    int enabled, lbc, rbc, n_src;
    const int64_t *src_idx;
    const double *src_val;   // (n_src), or (M, n_src) with src_val_stride = n_src
    int64_t src_val_stride;
};

// W warps per system, S warps of single-warp systems per CTA (S > 1 only with W == 1: no CTA barrier inside the loop
// then), L lanes per system (L < 32 only with W == 1: 32 / L small systems share a warp, shuffles stay inside a group)
// FACTORS: (dl, d, du) are the time-invariant factors (l, r = 1 / u, g) of fbr_factorize_f64 instead of the diagonals — the
// pivots are not recomputed, the sweeps are K3's (y_i = b_i - l_i y_{i-1}, t_i = r_i y_i, x_i = t_i - g_i x_{i+1}): ONE
// implicit step with reused factors (fbr_run_f64 with n_steps = 1 takes this kernel).
template <int CPT, int W, int S, int L, bool FACTORS = false>
__global__ void __launch_bounds__(32 * W * S, 512 / (32 * W * S)) solve_reg_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                               const double *__restrict__ du, const double *b, double *x,
                                                               int64_t M, int64_t nx, int32_t *__restrict__ singular,
                                                               const StepOverrides ov, const int prefetch) {
    static_assert(S == 1 || W == 1, "several systems per CTA only for single-warp systems");
    static_assert(CPT % 2 == 0, "16-byte accesses and sign-free chunk products need an even chunk");
    static_assert(L == 32 || W == 1, "sub-warp groups only for single-warp systems");
    constexpr int HP = CPT / 2, GPW = 32 / L;  // systems per warp
    __shared__ Mat2 s_w[W];
    __shared__ double s_E[2][W], s_T[2][W];
    __shared__ unsigned s_first[W];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane & (L - 1), grp = lane / L;
    const int wl = W > 1 ? warp : 0, sys = W > 1 ? 0 : warp * GPW + grp;
    const int nxi = (int)nx;
    const int cell0 = (wl * 32 + sub) * CPT;
    const int nvalid = min(max(nxi - cell0, 0), CPT);  // rows of the matrix in my chunk (the rest is identity padding)
    const int jlast = nxi - 1 - cell0;                 // chunk position of the last row, when it is mine
    const int keep_f = sub == 0 ? 0 : -1, keep_b = sub == L - 1 ? 0 : -1;
    const bool ptr_ok = ((reinterpret_cast<uintptr_t>(dl) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(du) |
                          reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * S * GPW;
    // Whole-warp systems of 8-cell threads: while a system is solved, the NEXT one's four rows arrive in shared memory by
    // four bulk copies on an mbarrier and are read with the rotated, conflict-free start piece (as in fbr_run_onchip.cu's
    // step-at-a-time path and fbr_run_pass.cu): the solve lives on the latency of loading its rows.
    constexpr bool kCanPrefetch = CPT == 8 && S == 1 && L == 32;
    extern __shared__ __align__(128) double s_rows[];  // [4][32 W CPT]
    constexpr int kPitch = 32 * W * CPT;
    __shared__ __align__(8) unsigned long long s_rbar;
    [[maybe_unused]] const unsigned rbar = smem_addr(&s_rbar);
    [[maybe_unused]] unsigned n_taken = 0;
    const bool pf = kCanPrefetch && prefetch;
    [[maybe_unused]] auto stage_system = [&](int64_t mm) {  // lane 0 of warps 0 .. min(W, 4) - 1
        if (lane != 0) return;
        const unsigned bytes = (unsigned)(nx * sizeof(double));
        const double *src[4] = {dl + mm * nx, d + mm * nx, du + mm * nx, b + mm * nx};
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (warp == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar), "r"(4u * bytes) : "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (warp == (W >= 4 ? q : W >= 2 ? (q & 1) : 0)) bulk_g2s(smem_addr(s_rows + q * kPitch), src[q], bytes, rbar);
    };
    [[maybe_unused]] auto take_row = [&](int q, double pad, double (&v)[CPT]) {
        const double *row = s_rows + q * kPitch;
        if (nvalid == CPT) {
            const int rot = (tid >> 1) & 3;
            double2 ld[4], mid[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ld[j] = *reinterpret_cast<const double2 *>(row + cell0 + 2 * ((j + rot) & 3));
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                mid[pp].x = (rot & 1) ? ld[(pp + 3) & 3].x : ld[pp].x;
                mid[pp].y = (rot & 1) ? ld[(pp + 3) & 3].y : ld[pp].y;
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                v[(2 * pp) % CPT] = (rot & 2) ? mid[(pp + 2) & 3].x : mid[pp].x;
                v[(2 * pp + 1) % CPT] = (rot & 2) ? mid[(pp + 2) & 3].y : mid[pp].y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] = j < nvalid ? row[cell0 + j] : pad;
        }
    };
    if (kCanPrefetch && pf) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if ((int64_t)blockIdx.x < M) stage_system(blockIdx.x);
    }
    // which of my rows get an overridden right-hand side: one byte per row (0xFF none, 0xFE zero, else source column);
    // later claims win (sources over boundary rows, later duplicates over earlier ones)
    uint64_t slots = ~0ull;
    if (ov.enabled) {
        if (cell0 == 0 && ov.lbc != FBR_BC_NONE) slots = (slots & ~0xFFull) | 0xFEull;
        if ((unsigned)jlast < (unsigned)CPT && ov.rbc != FBR_BC_NONE) slots = (slots & ~(0xFFull << (8 * jlast))) | (0xFEull << (8 * jlast));
        for (int k = 0; k < ov.n_src; ++k) {
            const int64_t jj = ov.src_idx[k] - cell0;
            if (jj >= 0 && jj < nvalid) slots = (slots & ~(0xFFull << (8 * jj))) | ((uint64_t)k << (8 * jj));
        }
    }

    // the loop bound is uniform over a warp (and over the CTA when W > 1); groups past the last system redo system
    // M - 1 and skip the stores
    for (int64_t m_raw = (int64_t)blockIdx.x * S * GPW + sys; m_raw - grp < M; m_raw += stride) {
        const bool live = m_raw < M;
        const int64_t m = live ? m_raw : M - 1;
        const int64_t o = m * nx;
        double l[CPT], dd[CPT], g[CPT], v[CPT];
        const bool vec = nvalid == CPT && ptr_ok && ((o + cell0) & 1) == 0;
        const bool odd = nvalid == CPT && ptr_ok && !vec;
        // super-diagonal entry in front of a warp's first chunk (the first padding row has none)
        double c_first = 0.0;
        if constexpr (kCanPrefetch) {
            if (pf) {  // (S == 1, L == 32: m_raw = blockIdx.x + k gridDim.x, always < M)
                mbar_wait(rbar, n_taken & 1u);
                ++n_taken;
                if constexpr (CPT == 8) { take_row(0, 0.0, l); take_row(1, 1.0, dd); take_row(2, 0.0, g); take_row(3, 0.0, v); }
                if (W > 1 && lane == 0 && cell0 > 0 && cell0 < nxi) c_first = s_rows[2 * kPitch + cell0 - 1];
                __syncthreads();  // everybody has its rows: the buffer is free for the next system's
                if (m_raw + stride < M) stage_system(m_raw + stride);
            }
        }
        if (!pf) {
            load_chunk<CPT>(dl + o + cell0, nvalid, vec, 0.0, l, odd);
            load_chunk<CPT>(d + o + cell0, nvalid, vec, 1.0, dd, odd);
            load_chunk<CPT>(du + o + cell0, nvalid, vec, 0.0, g, odd);
            load_chunk<CPT>(b + o + cell0, nvalid, vec, 0.0, v, odd);
            if (W > 1 && lane == 0 && cell0 > 0 && cell0 < nxi) c_first = __ldg(du + o + cell0 - 1);
        }
        if ((unsigned)jlast < (unsigned)CPT) {
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (j == jlast) g[j] = 0.0;
        }
        if (cell0 == 0) l[0] = 0.0;  // the first sub-diagonal entry is outside the matrix
        if (slots != ~0ull) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const unsigned so = (unsigned)(slots >> (8 * j)) & 0xFFu;
                if (so == 0xFEu) v[j] = 0.0;
                else if (so != 0xFFu) v[j] = __ldg(ov.src_val + m * ov.src_val_stride + so);
            }
        }
        double c_prev = __shfl_up_sync(kAll, g[CPT - 1], 1, L);
        if (sub == 0) c_prev = c_first;

        int bad = 0;
        if constexpr (!FACTORS) {
        // ---- 1. pivots ----
        double ee[CPT];  // e_j = -dl_j du_{j-1}
        {
            double cp = c_prev;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                ee[j] = -l[j] * cp;
                cp = g[j];
            }
        }
        Mat2 G{dd[0], ee[0], 1.0, 0.0};
#pragma unroll
        for (int j = 1; j < CPT; ++j) {
            const Mat2 nm{fma(dd[j], G.a, ee[j] * G.c), fma(dd[j], G.b, ee[j] * G.d), G.a, G.b};
            G = (j == CPT / 2 - 1) ? rescale_fast(nm) : nm;
        }
        Mat2 inc = rescale_fast(G);
#pragma unroll
        for (int s = 1; s < L; s <<= 1) {
            const Mat2 up = shfl_up_mat_w<L>(inc, s);
            if (sub >= s) inc = compose_raw(inc, up);
            if (s == 4 && L > 8) inc = rescale_fast(inc);
        }
        inc = rescale_fast(inc);
        Mat2 ex = shfl_up_mat_w<L>(inc, 1);
        if (sub == 0) ex = Mat2{1.0, 0.0, 0.0, 1.0};
        if (W > 1) {
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();
            Mat2 before{1.0, 0.0, 0.0, 1.0};
            if (W <= 4) {
                for (int q = 0; q < warp; ++q) before = rescale_fast(compose_raw(s_w[q], before));
            } else {  // every warp scans the W warp totals itself (lane = warp index) and picks the prefix in front of it
                Mat2 t = lane < W ? s_w[lane] : Mat2{1.0, 0.0, 0.0, 1.0};
#pragma unroll
                for (int s = 1; s < W; s <<= 1) {
                    const Mat2 up = shfl_up_mat(t, s);
                    if (lane >= s) t = rescale_fast(compose_raw(t, up));
                }
                const int srcl = warp > 0 ? warp - 1 : 0;
                before.a = __shfl_sync(kAll, t.a, srcl); before.b = __shfl_sync(kAll, t.b, srcl);
                before.c = __shfl_sync(kAll, t.c, srcl); before.d = __shfl_sync(kAll, t.d, srcl);
                if (warp == 0) before = Mat2{1.0, 0.0, 0.0, 1.0};
            }
            ex = compose_raw(ex, before);
        }
        // (p_{i-1}, p_{i-2}) in front of my chunk: the map applied to any finite start (e_0 = 0 makes the start irrelevant)
        double p1 = ex.a + ex.b, p2 = ex.c + ex.d;
        {
            const double s = pow2_scale_for(max(exp_field(p1), exp_field(p2)));
            p1 *= s; p2 *= s;
        }
        double r[CPT], nanacc = 0.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const double p0 = fma(dd[j], p1, ee[j] * p2);
            r[j] = div_fast(p1, p0);  // off the chain: the recurrence continues with p0
            nanacc = fma(r[j], 0.0, fma(p0, 0.0, nanacc));  // NaN <=> a zero pivot (r = inf) or an overflowed minor (p0 = inf)
            p2 = p1; p1 = p0;
            if (j == CPT / 2 - 1) {
                const double s = pow2_scale_for(max(exp_field(p1), exp_field(p2)));
                p1 *= s; p2 *= s;
            }
        }
        if (nanacc != nanacc) {  // a zero (or non-finite) pivot makes its quotient non-finite: find the first such row
#pragma unroll
            for (int j = CPT - 1; j >= 0; --j)
                if (j < nvalid && (!isfinite(r[j]) || r[j] == 0.0)) bad = cell0 + j + 1;
        }
        // normalised rows: z_i = r_i b_i - (r_i dl_i) z_{i-1};  x_i = z_i - (r_i du_i) x_{i+1}
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            l[j] *= r[j];
            g[j] *= r[j];
            v[j] *= r[j];
        }
        }
        const GroupMultipliers<L> sm = group_multipliers<L>(chunk_product<double, CPT>(l), chunk_product<double, CPT>(g), sub);
        constexpr int NS = GroupMultipliers<L>::NS;

        // ---- 2. the two sweeps ----
        {
            double y = chunk_end_fwd_plain<double, CPT>(v, l);
#pragma unroll
            for (int s = 0; s < NS; ++s) y = fma(sm.f[s], __shfl_up_sync(kAll, y, 1 << s, L), y);
            double c = and_mask(__shfl_up_sync(kAll, y, 1, L), keep_f);
            if (W > 1) {
                const unsigned first = __reduce_min_sync(kAll, bad ? (unsigned)bad : 0xffffffffu);
                if (lane == 31) { s_E[0][warp] = y; s_T[0][warp] = sm.tot_f; }
                if (lane == 0) { s_T[1][warp] = sm.tot_b; s_first[warp] = first; }
                __syncthreads();
                if (singular && tid == 0) {  // first zero / non-finite pivot of the system (the flags ride on this barrier)
                    unsigned f = 0xffffffffu;
                    for (int q = 0; q < W; ++q) f = min(f, s_first[q]);
                    singular[m] = f == 0xffffffffu ? 0 : (int32_t)f;
                }
                double cw = 0.0;
                for (int q = 0; q < warp; ++q) cw = fma(s_T[0][q], cw, s_E[0][q]);
                c = fma(sm.ex_f, cw, c);
            }
            double yy = c;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                yy = fma(-l[j], yy, v[j]);
                v[j] = FACTORS ? dd[j] * yy : yy;  // (FACTORS: dd holds r = 1 / u)
            }
        }
        {
            double z = chunk_end_bwd_plain<double, CPT>(v, g);
#pragma unroll
            for (int s = 0; s < NS; ++s) z = fma(sm.b[s], __shfl_down_sync(kAll, z, 1 << s, L), z);
            double c = and_mask(__shfl_down_sync(kAll, z, 1, L), keep_b);
            if (W > 1) {
                if (lane == 0) s_E[1][warp] = z;
                __syncthreads();
                double cw = 0.0;
                for (int q = W - 1; q > warp; --q) cw = fma(s_T[1][q], cw, s_E[1][q]);
                c = fma(sm.ex_b, cw, c);
            }
            chunk_apply_bwd_plain<double, CPT>(v, g, c);
        }
        // ---- store ----
        if (live) {
            if (vec) {
                double2 *q = reinterpret_cast<double2 *>(x + o + cell0);
#pragma unroll
                for (int j = 0; j < HP; ++j) q[j] = make_double2(v[2 * j], v[2 * j + 1]);
            } else if (odd) {
                x[o + cell0] = v[0];
                double2 *q = reinterpret_cast<double2 *>(x + o + cell0 + 1);
#pragma unroll
                for (int j = 0; j < HP - 1; ++j) q[j] = make_double2(v[2 * j + 1], v[2 * j + 2]);
                x[o + cell0 + CPT - 1] = v[CPT - 1];
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    if (j < nvalid) x[o + cell0 + j] = v[j];
            }
        }
        if (singular && W == 1) {
            unsigned first = bad ? (unsigned)bad : 0xffffffffu;
            if (L == 32) {
                first = __reduce_min_sync(kAll, first);
            } else {
#pragma unroll
                for (int s = 1; s < L; s <<= 1) first = min(first, __shfl_xor_sync(kAll, first, s, L));
            }
            if (sub == 0 && live) singular[m] = first == 0xffffffffu ? 0 : (int32_t)first;
        }
    }
}

template <int CPT, int W, int S, int L, bool FACTORS = false>
static int launch_reg(const double *dl, const double *d, const double *du, const double *b, double *x, int64_t M, int64_t nx,
                      int32_t *singular, const StepOverrides &ov, cudaStream_t stream) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    auto kern = solve_reg_kernel<CPT, W, S, L, FACTORS>;
    constexpr int NT = 32 * W * S;
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int prefetch = CPT == 8 && S == 1 && L == 32 && nx % 2 == 0 && al16(dl) && al16(d) && al16(du) && al16(b) && M > 1 &&
                         !getenv("FBR_K2_NO_PREFETCH");
    const size_t smem = prefetch ? 4 * (size_t)(32 * W * CPT) * sizeof(double) : 0;
    if (smem > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "REG solve kernel does not fit on an SM (threads=%d)", NT);
    int64_t grid = (int64_t)occ * f.sm_count;
    const int64_t per_cta = S * (32 / L);
    const int64_t groups = (M + per_cta - 1) / per_cta;
    if (grid > groups) grid = groups;
    kern<<<(unsigned)grid, NT, smem, stream>>>(dl, d, du, b, x, M, nx, singular, ov, prefetch);
    return check_launch("fbr_solve_tridiag_f64 (REG)");
}

static int reg_dispatch(const double *dl, const double *d, const double *du, const double *b, double *x, int64_t M, int64_t nx,
                        int32_t *singular, const StepOverrides &ov, cudaStream_t stream) {
    if (nx > kSolveRegMaxNx)
        return fail(FBR_ERR_UNSUPPORTED, "REG variant covers nx <= %d (got %lld)", kSolveRegMaxNx, (long long)nx);
    if (nx <= 16) return launch_reg<2, 1, 4, 8>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 32) return launch_reg<4, 1, 4, 8>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 64) return launch_reg<8, 1, 4, 8>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 128) return launch_reg<8, 1, 4, 16>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 256) return launch_reg<8, 1, 4, 32>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 512) return launch_reg<8, 2, 1, 32>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 1024) return launch_reg<8, 4, 1, 32>(dl, d, du, b, x, M, nx, singular, ov, stream);
    if (nx <= 2048) return launch_reg<8, 8, 1, 32>(dl, d, du, b, x, M, nx, singular, ov, stream);
    return launch_reg<8, 16, 1, 32>(dl, d, du, b, x, M, nx, singular, ov, stream);
}

// one implicit step with reused factors (fl, fr, fg): right-hand side overrides + K3's two sweeps, 129 <= nx <= 4096
int step_factors_dispatch(const double *fl, const double *fr, const double *fg, const double *p, double *p_next, int64_t M, int64_t nx,
                          int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val, cudaStream_t stream) {
    StepOverrides ov{1, lbc, rbc, n_src, src_idx, src_val, 0};
    if (nx <= 512) return launch_reg<8, 2, 1, 32, true>(fl, fr, fg, p, p_next, M, nx, nullptr, ov, stream);
    if (nx <= 1024) return launch_reg<8, 4, 1, 32, true>(fl, fr, fg, p, p_next, M, nx, nullptr, ov, stream);
    if (nx <= 2048) return launch_reg<8, 8, 1, 32, true>(fl, fr, fg, p, p_next, M, nx, nullptr, ov, stream);
    return launch_reg<8, 16, 1, 32, true>(fl, fr, fg, p, p_next, M, nx, nullptr, ov, stream);
}

int solve_reg_dispatch(const double *dl, const double *d, const double *du, const double *b, double *x, int64_t M, int64_t nx,
                       int32_t *singular, cudaStream_t stream) {
    StepOverrides ov{};
    return reg_dispatch(dl, d, du, b, x, M, nx, singular, ov, stream);
}

int step_reg_dispatch(const double *dl, const double *d, const double *du, const double *p, double *p_next, int64_t M, int64_t nx,
                      int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val, int64_t src_val_stride,
                      int32_t *singular, cudaStream_t stream) {
    StepOverrides ov{1, lbc, rbc, n_src, src_idx, src_val, src_val_stride};
    return reg_dispatch(dl, d, du, p, p_next, M, nx, singular, ov, stream);
}

}  // namespace fbr
