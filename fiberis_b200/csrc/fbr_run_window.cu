// K3W  time-tiled integration of ONE LONG mesh through overlapping on-chip windows.
//
// The two sweeps of a step are first-order recurrences whose multipliers (l, g of fbr_factorize_f64)
// are below one in magnitude for a diffusion matrix, so the influence of a value decays
// geometrically along the mesh: after H cells the product of multipliers is below 2^-80 of
// anything the state holds, i.e. ~8 orders of magnitude below one fp64 rounding.  H (the "decay
// horizon", measured from the actual factors by fbr_decay_horizon_f64, with a rigorous bound for
// rows whose multiplier exceeds one) is typically tens of cells.
//
// That makes time blocking possible for meshes that do not fit on chip: a CTA loads a WINDOW of
// 32*W*8 consecutive cells (pressure + the three factors -> registers), advances it s time steps
// with the same register-resident sweeps as K3 (fbr_run_onchip.cu), and writes back only the cells
// it OWNS: the window minus s*Hf cells on the left and s*Hb cells on the right, whose values are
// contaminated by the missing neighbours (one horizon per step).  Windows overlap by exactly those
// halos, no CTA ever talks to another one, and a launch reads p^n from one buffer and writes
// p^{n+s} to another.  HBM traffic per cell-update drops from 40 B (one streaming solve per step)
// to (32 B * window/owned + 8 B) / s.
//
// Nothing is approximated beyond the 2^-80 truncation, and the path is only taken when the measured
// horizon is short enough (fbr_run_window_plan); otherwise the exact long-mesh kernels (K3L / K3S) run.
#include "fbr_common.cuh"

namespace fbr {

constexpr unsigned kFullW = 0xffffffffu;
constexpr int kWinCPT = 8;

struct WinArgs {
    const double *fl, *fr, *fg;  // (nx) factors
    const double *p_in;          // (nx) state at the start of the block
    double *p_out;               // (nx) state at the end of the block (owned cells of every window)
    int64_t nx;
    int64_t own, lead;  // owned cells per window; window k starts at cell k * own - lead
    int n_steps;        // steps in this block
    int nbf, nbb;       // number of preceding / following warps that can still reach a warp's carry
    int lbc, rbc;
    const int64_t *src_idx;
    int n_src;
    const double *src_table;  // (n_steps, n_src) rows of this block
};

__device__ const double kZeroWin = 0.0;

__device__ __forceinline__ void mov_if_eq_w(double &dst, double v, int j, int k) {
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, %3; @p mov.f64 %0, %1; }" : "+d"(dst) : "d"(v), "r"(j), "r"(k));
}

// 8 consecutive cells starting at `cell0` of a per-cell array; cells outside [0, nx) read `pad`
__device__ __forceinline__ void load8(const double *__restrict__ base, int64_t cell0, int64_t nx, double pad,
                                      double (&v)[kWinCPT]) {
    const double *p = base + cell0;
    if (cell0 >= 0 && cell0 + kWinCPT <= nx && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int j = 0; j < kWinCPT / 2; ++j) {
            const double2 t = __ldg(q + j);
            v[2 * j] = t.x;
            v[2 * j + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kWinCPT; ++j) {
            const int64_t i = cell0 + j;
            v[j] = (i >= 0 && i < nx) ? __ldg(base + i) : pad;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(32 * W, W >= 16 ? 1 : 2) run_window_kernel(const WinArgs a) {
    constexpr int CPT = kWinCPT;
    __shared__ __align__(16) double s_K[2][W][W];
    __shared__ __align__(16) double s_E[2][W];
    __shared__ double s_T[2][W];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nx = a.nx;
    const int64_t own0 = (int64_t)blockIdx.x * a.own;
    const int64_t own1 = own0 + a.own < nx ? own0 + a.own : nx;
    const int64_t cell0 = own0 - a.lead + (int64_t)tid * CPT;

    // ---------------- window setup ----------------
    double l[CPT], r[CPT], g[CPT], x[CPT];
    load8(a.fl, cell0, nx, 0.0, l);
    load8(a.fr, cell0, nx, 1.0, r);
    load8(a.fg, cell0, nx, 0.0, g);
    load8(a.p_in, cell0, nx, 0.0, x);

    // right-hand-side overrides owned by this thread (same scheme as fbr_run_onchip.cu), by GLOBAL cell
    int ov_j = -1, ov_stride = 0;
    bool ov_slow = false;
    const double *ovp = &kZeroWin;
    if (cell0 < nx && cell0 + CPT > 0) {
        constexpr unsigned kNoSlot = 0xFF, kZeroSlot = 0xFE;
        uint64_t ov = ~0ull;
        if (cell0 <= 0 && a.lbc != FBR_BC_NONE) ov = (ov & ~(0xFFull << (8 * (-cell0)))) | ((uint64_t)kZeroSlot << (8 * (-cell0)));
        const int64_t jl = nx - 1 - cell0;
        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) ov = (ov & ~(0xFFull << (8 * jl))) | ((uint64_t)kZeroSlot << (8 * jl));
        bool wide = false;
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {
            const int64_t s = a.src_idx[k];
            const int64_t jj = s - cell0;
            if (jj >= 0 && jj < CPT && s >= 0 && s < nx) {
                if (k >= 0xFE) wide = true;
                ov = (ov & ~(0xFFull << (8 * jj))) | ((uint64_t)(k & 0xFF) << (8 * jj));
            }
        }
        int n_ov = 0;
#pragma unroll 1
        for (int j = 0; j < CPT; ++j) {
            const unsigned so = (unsigned)(ov >> (8 * j)) & 0xFF;
            if (so != kNoSlot) {
                ++n_ov; ov_j = j;
                if (so != kZeroSlot) { ovp = a.src_table + so; ov_stride = a.n_src; }
                else { ovp = &kZeroWin; ov_stride = 0; }
            }
        }
        if (n_ov > 1 || wide) { ov_slow = true; ov_j = -1; ovp = &kZeroWin; ov_stride = 0; }
    }

    // scan multipliers (time-invariant within the block)
    double PfL[5], PbL[5], PfEx, PbEx;
    {
        double pf = 1.0, pb = 1.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) { pf *= -l[j]; pb *= -g[j]; }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const bool fa = lane >= (1 << s), ba = lane + (1 << s) < 32;
            PfL[s] = fa ? pf : 0.0;
            PbL[s] = ba ? pb : 0.0;
            const double upf = __shfl_up_sync(kFullW, pf, 1 << s);
            const double dnb = __shfl_down_sync(kFullW, pb, 1 << s);
            if (fa) pf *= upf;
            if (ba) pb *= dnb;
        }
        PfEx = __shfl_up_sync(kFullW, pf, 1);
        PbEx = __shfl_down_sync(kFullW, pb, 1);
        if (lane == 0) PfEx = 1.0;
        if (lane == 31) PbEx = 1.0;
        if (lane == 31) s_T[0][warp] = pf;
        if (lane == 0) s_T[1][warp] = pb;
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
    const int vf0 = warp - a.nbf > 0 ? warp - a.nbf : 0;          // forward carry: warps vf0 .. warp-1
    const int vb1 = warp + a.nbb < W - 1 ? warp + a.nbb : W - 1;  // backward carry: warps warp+1 .. vb1
    const bool near_f = a.nbf == 1, near_b = a.nbb == 1;
    double ov_val = a.n_steps > 0 ? __ldg(ovp) : 0.0;

    // ---------------- time loop ----------------
#pragma unroll 1
    for (int left = a.n_steps; left > 0; --left) {
        if (left == 1) ov_stride = 0;
        ovp += ov_stride;
        const double ov_next = __ldg(ovp);
        if (ov_j >= 0) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) mov_if_eq_w(x[k], ov_val, ov_j, k);
        }
        if (ov_slow) {
            const int64_t n = a.n_steps - left;
            if (cell0 <= 0 && cell0 + CPT > 0 && a.lbc != FBR_BC_NONE) {
#pragma unroll
                for (int k = 0; k < CPT; ++k) mov_if_eq_w(x[k], 0.0, (int)(-cell0), k);
            }
            const int64_t jl = nx - 1 - cell0;
            if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) {
#pragma unroll
                for (int k = 0; k < CPT; ++k) mov_if_eq_w(x[k], 0.0, (int)jl, k);
            }
#pragma unroll 1
            for (int q = 0; q < a.n_src; ++q) {
                const int64_t s = a.src_idx[q];
                const int64_t jj = s - cell0;
                if (jj >= 0 && jj < CPT && s >= 0 && s < nx) {
                    const double v = __ldg(a.src_table + n * a.n_src + q);
#pragma unroll
                    for (int k = 0; k < CPT; ++k) mov_if_eq_w(x[k], v, (int)jj, k);
                }
            }
        }
        // ---- forward sweep ----
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) y = fma(-l[j], y, x[j]);
#pragma unroll
        for (int s = 0; s < 5; ++s) y = fma(PfL[s], __shfl_up_sync(kFullW, y, 1 << s), y);
        double carry = __shfl_up_sync(kFullW, y, 1);
        if (lane == 0) carry = 0.0;
        if (lane == 31) s_E[0][warp] = y;
        __syncthreads();
        double c = 0.0;
        if (near_f) {
            if (warp > 0) c = s_E[0][warp - 1];
        } else {
            for (int v = vf0; v < warp; ++v) c = fma(s_K[0][warp][v], s_E[0][v], c);
        }
        y = fma(PfEx, c, carry);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            y = fma(-l[j], y, x[j]);
            x[j] = r[j] * y;
        }
        // ---- backward sweep ----
        double z = 0.0;
#pragma unroll
        for (int j = CPT - 1; j >= 0; --j) z = fma(-g[j], z, x[j]);
#pragma unroll
        for (int s = 0; s < 5; ++s) z = fma(PbL[s], __shfl_down_sync(kFullW, z, 1 << s), z);
        carry = __shfl_down_sync(kFullW, z, 1);
        if (lane == 31) carry = 0.0;
        if (lane == 0) s_E[1][warp] = z;
        __syncthreads();
        c = 0.0;
        if (near_b) {
            if (warp < W - 1) c = s_E[1][warp + 1];
        } else {
            for (int v = vb1; v > warp; --v) c = fma(s_K[1][warp][v], s_E[1][v], c);
        }
        z = fma(PbEx, c, carry);
#pragma unroll
        for (int j = CPT - 1; j >= 0; --j) {
            z = fma(-g[j], z, x[j]);
            x[j] = z;
        }
        ov_val = ov_next;
    }

    // ---------------- write back the owned cells ----------------
    double *dst = a.p_out + cell0;
    if (cell0 >= own0 && cell0 + CPT <= own1 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        double2 *q = reinterpret_cast<double2 *>(dst);
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j) q[j] = make_double2(x[2 * j], x[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int64_t i = cell0 + j;
            if (i >= own0 && i < own1) dst[j] = x[j];
        }
    }
}

// ---- decay horizon of the factors ----------------------------------------------------------------
// A product of consecutive multipliers normally only shrinks, but single multipliers can exceed one
// (the row after a source or Neumann row, cells where the mesh spacing jumps), so "the product has
// dropped below eps" is only final once it is below eps / G, G = the largest factor by which ANY
// run of consecutive multipliers can grow.  probe_kernel bounds G: every cell walks kProbe cells,
// recording the largest running product (local growth) and the product after kProbe cells (q).
// If every q < 1, any longer run is a kProbe-run times a shorter-growth run, so G = max local growth;
// otherwise the horizon is reported as "longer than cap" and the exact kernels are used.
// acc[0..3] = max local growth fwd, max q fwd, max local growth bwd, max q bwd (as ordered bit patterns).
constexpr int kProbe = 64;

__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {  // v >= 0: bit patterns are ordered
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void probe_kernel(const double *__restrict__ fl, const double *__restrict__ fg, int64_t nx, double *acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double gf = 1.0, qf = 0.0, gb = 1.0, qb = 0.0;
    if (i < nx) {
        double p = 1.0;
        int64_t k = i;
        for (int n = 0; n < kProbe && k >= 0; ++n, --k) {
            p *= fabs(__ldg(fl + k));
            if (!(p <= gf)) gf = p;  // NaN lands here and poisons the maximum -> unsupported
        }
        qf = k >= 0 ? p : 0.0;  // a run that leaves the mesh ends there
        p = 1.0; k = i;
        for (int n = 0; n < kProbe && k < nx; ++n, ++k) {
            p *= fabs(__ldg(fg + k));
            if (!(p <= gb)) gb = p;
        }
        qb = k < nx ? p : 0.0;
    }
    if (!(gf < 1e300)) gf = 1e300;
    if (!(gb < 1e300)) gb = 1e300;
    if (!(qf < 1e300)) qf = 1e300;
    if (!(qb < 1e300)) qb = 1e300;
    for (int s = 16; s > 0; s >>= 1) {
        gf = fmax(gf, __shfl_xor_sync(kFullW, gf, s));
        qf = fmax(qf, __shfl_xor_sync(kFullW, qf, s));
        gb = fmax(gb, __shfl_xor_sync(kFullW, gb, s));
        qb = fmax(qb, __shfl_xor_sync(kFullW, qb, s));
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_max_pos(acc, gf); atomic_max_pos(acc + 1, qf);
        atomic_max_pos(acc + 2, gb); atomic_max_pos(acc + 3, qb);
    }
}

// horizon[0] = 1 + the largest n for which some product |l_i l_{i-1} .. l_{i-n+1}| is still >= eps
// (so every run of >= horizon[0] multipliers is below eps); horizon[1] likewise for g going right;
// cap + 1 when that cannot be established within `cap` cells.
__global__ void horizon_kernel(const double *__restrict__ fl, const double *__restrict__ fg, int64_t nx, double log2_eps,
                               int cap, const double *__restrict__ acc, unsigned long long *horizon) {
    const double eps = exp2(log2_eps);
    const bool ok_f = acc[1] < 1.0, ok_b = acc[3] < 1.0;
    const double lo_f = eps / acc[0], lo_b = eps / acc[2];  // below this a product can never come back above eps
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int nf = 0, nb = 0;
    if (i < nx) {
        double p = 1.0;
        int64_t k = i;
        int n = 0, last = 0;
        bool sure = false;
        while (n < cap && k >= 0) {
            p *= fabs(__ldg(fl + k));
            ++n; --k;
            if (p >= eps) last = n;
            if (p < lo_f || p == 0.0) { sure = true; break; }
        }
        nf = (ok_f && (sure || k < 0)) ? last + 1 : cap + 1;
        p = 1.0; k = i; n = 0; last = 0; sure = false;
        while (n < cap && k < nx) {
            p *= fabs(__ldg(fg + k));
            ++n; ++k;
            if (p >= eps) last = n;
            if (p < lo_b || p == 0.0) { sure = true; break; }
        }
        nb = (ok_b && (sure || k >= nx)) ? last + 1 : cap + 1;
    }
    for (int s = 16; s > 0; s >>= 1) {
        nf = max(nf, __shfl_xor_sync(kFullW, nf, s));
        nb = max(nb, __shfl_xor_sync(kFullW, nb, s));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(horizon, (unsigned long long)nf);
        atomicMax(horizon + 1, (unsigned long long)nb);
    }
}

struct WinPlan {
    int W = 0;          // warps per window CTA
    int64_t s = 0;      // time steps per block
    int64_t own = 0, lead = 0, n_win = 0;
    double us_per_step = 0.0;  // model estimate
};

static inline int64_t round_up8(int64_t v) { return (v + 7) / 8 * 8; }

// Pick (W, s): a block of s steps costs  waves * (t_fixed + s * t_step) + t_launch  with
// waves = windows / resident CTAs; minimise the cost per step.  max_s caps the block length
// (snapshot interval).  Constants are measured orders of magnitude on a B200, not tuned per run.
static bool plan_window(int64_t nx, int64_t Hf, int64_t Hb, int64_t max_s, int sm_count, WinPlan *best) {
    const int64_t hf = round_up8(Hf > 0 ? Hf : 1), hb = round_up8(Hb > 0 ? Hb : 1);
    int force_w = 0;
    int64_t force_s = 0;
    if (const char *e = getenv("FBR_WIN_W")) force_w = atoi(e);
    if (const char *e = getenv("FBR_WIN_S")) force_s = atoll(e);
    if (force_s > max_s) force_s = max_s;
    bool found = false;
    for (int W : {8, 16}) {
        if (force_w && W != force_w) continue;
        const int64_t wc = 32 * (int64_t)W * kWinCPT;
        const int resident = (W >= 16 ? 1 : 2) * sm_count;
        const double t_step = W >= 16 ? 1.0 : 0.55, t_fixed = W >= 16 ? 4.0 : 3.0, t_launch = 2.5;
        for (int64_t s = 1; s <= max_s; ++s) {
            if (force_s && s != force_s) continue;
            const int64_t own = wc - s * (hf + hb);
            if (own < wc / 8) break;
            const int64_t n_win = (nx + own - 1) / own;
            const double waves = (double)((n_win + resident - 1) / resident);
            // a partially filled last wave still costs a full one; long runs of waves pipeline well
            const double cost = (waves * (t_fixed + s * t_step) + t_launch) / s;
            if (!found || cost < best->us_per_step) {
                found = true;
                best->W = W; best->s = s; best->own = own; best->lead = s * hf; best->n_win = n_win;
                best->us_per_step = cost;
            }
        }
    }
    return found;
}

}  // namespace fbr

using namespace fbr;

extern "C" int fbr_decay_horizon_f64(const double *fl, const double *fg, int64_t nx, double log2_eps, int cap,
                                     int64_t *horizon, double *work, void *stream) {
    FBR_REQUIRE(fl && fg && horizon && work && nx >= 1 && cap >= 1 && log2_eps < 0.0, "bad decay-horizon arguments");
    cudaStream_t st = as_stream(stream);
    FBR_CUDA(cudaMemsetAsync(work, 0, 4 * sizeof(double), st));
    FBR_CUDA(cudaMemsetAsync(horizon, 0, 2 * sizeof(int64_t), st));
    const int64_t nb = (nx + 255) / 256;
    probe_kernel<<<(unsigned)nb, 256, 0, st>>>(fl, fg, nx, work);
    if (int rc = check_launch("fbr_decay_horizon (probe)")) return rc;
    horizon_kernel<<<(unsigned)nb, 256, 0, st>>>(fl, fg, nx, log2_eps, cap, work,
                                                  reinterpret_cast<unsigned long long *>(horizon));
    return check_launch("fbr_decay_horizon");
}

extern "C" int fbr_run_window_plan(int64_t nx, int64_t horizon_fwd, int64_t horizon_bwd, int64_t snapshot_every,
                                   int64_t *plan4) {
    FBR_REQUIRE(nx >= 1 && horizon_fwd >= 0 && horizon_bwd >= 0 && plan4, "bad window-plan arguments");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    WinPlan pl;
    const int64_t max_s = snapshot_every >= 1 ? (snapshot_every < 4096 ? snapshot_every : 4096) : 4096;
    if (!plan_window(nx, horizon_fwd, horizon_bwd, max_s, f.sm_count, &pl))
        return fail(FBR_ERR_UNSUPPORTED, "decay horizon (%lld, %lld) is too long for the windowed path",
                    (long long)horizon_fwd, (long long)horizon_bwd);
    plan4[0] = pl.W; plan4[1] = pl.s; plan4[2] = pl.own; plan4[3] = pl.n_win;
    return FBR_OK;
}

extern "C" int64_t fbr_run_window_work_bytes(int64_t nx) { return 2 * nx * (int64_t)sizeof(double); }

extern "C" int fbr_run_window_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                  int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                  const double *src_table, int64_t snapshot_every, int64_t snap_row_stride, double *snap,
                                  double *p_final, int64_t horizon_fwd, int64_t horizon_bwd, double *work, void *stream) {
    FBR_REQUIRE(nx >= 1 && n_steps >= 0, "bad nx / n_steps");
    FBR_REQUIRE(fl && fr && fg && p0 && work, "NULL factor / initial-state / work pointer");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_table)), "sources given without src_idx / src_table");
    FBR_REQUIRE(!snap || (snapshot_every >= 1 && snap_row_stride >= nx), "bad snapshot layout");
    FBR_REQUIRE(horizon_fwd >= 0 && horizon_bwd >= 0, "bad decay horizon");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    cudaStream_t st = as_stream(stream);
    WinPlan pl;
    const int64_t max_s = snap ? (snapshot_every < 4096 ? snapshot_every : 4096) : 4096;
    if (!plan_window(nx, horizon_fwd, horizon_bwd, max_s, f.sm_count, &pl))
        return fail(FBR_ERR_UNSUPPORTED, "decay horizon (%lld, %lld) is too long for the windowed path",
                    (long long)horizon_fwd, (long long)horizon_bwd);
    // equalise the blocks inside one snapshot interval
    int64_t s_blk = pl.s;
    if (snap) {
        const int64_t nblk = (snapshot_every + s_blk - 1) / s_blk;
        s_blk = (snapshot_every + nblk - 1) / nblk;
    }
    const int64_t warp_span = 32 * kWinCPT;
    WinArgs a;
    a.fl = fl; a.fr = fr; a.fg = fg; a.nx = nx; a.own = pl.own; a.lead = pl.lead; a.lbc = lbc; a.rbc = rbc;
    a.src_idx = src_idx; a.n_src = n_src;
    a.nbf = (int)((horizon_fwd + warp_span - 1) / warp_span); if (a.nbf < 1) a.nbf = 1;
    a.nbb = (int)((horizon_bwd + warp_span - 1) / warp_span); if (a.nbb < 1) a.nbb = 1;
    double *buf[2] = {work, work + nx};
    const double *cur = p0;
    int64_t n = 0, snap_row = 0;
    int which = 0;
    while (n < n_steps) {
        int64_t len = s_blk;
        if (snap) {
            const int64_t to_snap = snapshot_every - n % snapshot_every;
            if (len > to_snap) len = to_snap;
        }
        if (len > n_steps - n) len = n_steps - n;
        const bool at_snap = snap && (n + len) % snapshot_every == 0;
        const bool at_end = n + len == n_steps;
        double *out;
        if (at_snap) out = snap + snap_row * snap_row_stride;
        else if (at_end && p_final) out = p_final;
        else { out = buf[which]; which ^= 1; }
        a.p_in = cur; a.p_out = out; a.n_steps = (int)len;
        a.src_table = src_table ? src_table + n * n_src : nullptr;
        if (pl.W >= 16) run_window_kernel<16><<<(unsigned)pl.n_win, 512, 0, st>>>(a);
        else run_window_kernel<8><<<(unsigned)pl.n_win, 256, 0, st>>>(a);
        if (int rc = check_launch("fbr_run_window_f64")) return rc;
        if (at_snap) ++snap_row;
        if (at_end && at_snap && p_final)
            FBR_CUDA(cudaMemcpyAsync(p_final, out, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
        cur = out;
        n += len;
    }
    if (n_steps == 0 && p_final) FBR_CUDA(cudaMemcpyAsync(p_final, p0, nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return FBR_OK;
}
