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
#include "fbr_sweep.cuh"

namespace fbr {

constexpr unsigned kFullW = 0xffffffffu;
constexpr int kWinCPT = 8;
constexpr int kWinList = 64;  // sources per window kept in shared memory (more: global scan)

// T = double (8 cells per thread) or float (the fp32 mode: 16 cells per thread, the same 64 bytes — a window holds
// twice the cells, the halos stay the same number of cells, and the sweeps run on the fp32 pipe)
#ifdef FBR_K3W_PROF  // debug build: clock64 stamps of a window's phases for a few CTAs (scratch/k3w_prof.py reads them)
__device__ long long g_k3w_prof[8 * 8 * 8];
#define K3W_STAMP(n) do { if (threadIdx.x == 0 && blockIdx.x % 17 == 0 && blockIdx.x / 17 < 8 && wi < 8) g_k3w_prof[((blockIdx.x / 17) * 8 + wi) * 8 + (n)] = clock64(); } while (0)
#else
#define K3W_STAMP(n) do { } while (0)
#endif

template <typename T>
struct WinArgsT {
    const T *fl, *fr, *fg;  // (nx) factors
    const T *p_in;          // (nx) state at the start of the block
    T *p_out;               // (nx) state at the end of the block (owned cells of every window)
    int64_t nx;
    int64_t own, lead;  // owned cells per window; window k starts at cell k * own - lead
    int n_steps;        // steps in this block
    int nbf, nbb;       // number of preceding / following warps that can still reach a warp's carry
    int64_t n_win;      // windows in the mesh
    int aligned;        // all four per-cell arrays are 16-byte aligned (asynchronous staging allowed)
    int lbc, rbc;
    const int64_t *src_idx;
    int n_src;
    const double *src_table;  // (n_steps, n_src) rows of this block
};
using WinArgs = WinArgsT<double>;

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

__device__ __forceinline__ void mov_if_eq_w(float &dst, float v, int j, int k) {
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, %3; @p mov.f32 %0, %1; }" : "+f"(dst) : "f"(v), "r"(j), "r"(k));
}
// fp32 mode: 16 consecutive cells
__device__ __forceinline__ void load8(const float *__restrict__ base, int64_t cell0, int64_t nx, float pad, float (&v)[16]) {
    const float *p = base + cell0;
    if (cell0 >= 0 && cell0 + 16 <= nx && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = __ldg(q + j);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t i = cell0 + j;
            v[j] = (i >= 0 && i < nx) ? __ldg(base + i) : pad;
        }
    }
}
__device__ __forceinline__ float lds_any(unsigned addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_any(unsigned addr, double) { return lds_f64(addr); }

// Staging of a window's rows in shared memory: BULK copies (cp.async.bulk, the 1-D TMA path) issued by four threads, one per
// array, completing on an mbarrier — round 1 had every thread issue sixteen 16-byte cp.async into XOR-swizzled rows, which
// cost 1 300 cycles of issue per window (scratch/k3w_prof.py).  The rows land LINEAR: thread t's 64 bytes of an array sit
// at byte 64 t, so the same 16-byte piece of every thread would hit 2 of the 8 bank groups.  Thread t therefore starts with
// piece (t / 2) % 4 — a quarter warp then covers all 8 groups — and rotates the four pieces back with two select stages.
template <typename R, int CPT>
__device__ __forceinline__ void unstage_row(const R *row, int tid, R (&v)[CPT]) {
    const int rot = (tid >> 1) & 3;
    const char *base = reinterpret_cast<const char *>(row) + tid * 64;
    int4 ld[4], mid[4], out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ld[j] = *reinterpret_cast<const int4 *>(base + 16 * ((j + rot) & 3));  // piece (j + rot) % 4
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int4 a = ld[(p + 3) & 3], b = ld[p];
        mid[p] = (rot & 1) ? a : b;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int4 a = mid[(p + 2) & 3], b = mid[p];
        out[p] = (rot & 2) ? a : b;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if constexpr (sizeof(R) == 8) {
            v[2 * p] = __hiloint2double(out[p].y, out[p].x);
            v[2 * p + 1] = __hiloint2double(out[p].w, out[p].z);
        } else {
            v[4 * p] = __int_as_float(out[p].x); v[4 * p + 1] = __int_as_float(out[p].y);
            v[4 * p + 2] = __int_as_float(out[p].z); v[4 * p + 3] = __int_as_float(out[p].w);
        }
    }
}

// the reverse: thread t's 64 bytes into byte 64 t of a linear row, conflict-free (piece (j + rot) % 4 goes out j-th)
template <typename R, int CPT>
__device__ __forceinline__ void stage_row_out(R *row, int tid, const R (&v)[CPT]) {
    const int rot = (tid >> 1) & 3;
    char *base = reinterpret_cast<char *>(row) + tid * 64;
    int4 in[4], mid[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if constexpr (sizeof(R) == 8)
            in[p] = make_int4(__double2loint(v[2 * p]), __double2hiint(v[2 * p]), __double2loint(v[2 * p + 1]), __double2hiint(v[2 * p + 1]));
        else
            in[p] = make_int4(__float_as_int(v[4 * p]), __float_as_int(v[4 * p + 1]), __float_as_int(v[4 * p + 2]), __float_as_int(v[4 * p + 3]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int4 a = in[(j + 1) & 3], b = in[j];
        mid[j] = (rot & 1) ? a : b;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // out-th piece = in[(j + rot) % 4]
        const int4 a = mid[(j + 2) & 3], b = mid[j];
        *reinterpret_cast<int4 *>(base + 16 * ((j + rot) & 3)) = (rot & 2) ? a : b;
    }
}

// Persistent CTAs: CTA b integrates windows b, b + gridDim.x, ...; while it advances one window
// through the time block, the next window's pressure and factors stream into shared memory
// (bulk copies), so HBM latency is hidden behind the sweeps.
// NS = scan stages (5: exact; 3 / 4 when the horizon is <= 64 / 128 cells); NEAR: the horizon fits
// one warp (256 cells), so a warp's carry comes from its immediate neighbour only.
template <typename R, int W, int NS, bool NEAR>
__global__ void __launch_bounds__(32 * W, W >= 16 ? 1 : 2) run_window_kernel(const WinArgsT<R> a) {
    constexpr bool F64 = sizeof(R) == 8;
    constexpr int CPT = 64 / (int)sizeof(R);
    constexpr int T = 32 * W;
    constexpr int WC = T * CPT;  // cells per window
    extern __shared__ __align__(128) double s_stage_raw[];  // [4][WC] of R: l, r, g, p of the NEXT window; [WC]: the state going out
    R *const s_stage = reinterpret_cast<R *>(s_stage_raw);
    R *const s_out = s_stage + 4 * WC;
    __shared__ __align__(16) R s_K[2][W][W];
    __shared__ __align__(16) R s_Ef[W + 1];  // s_Ef[w + 1] = zero-carry end value of warp w; s_Ef[0] = 0
    __shared__ __align__(16) R s_Eb[W + 1];  // s_Eb[w] = zero-carry first value of warp w; s_Eb[W] = 0
    __shared__ R s_dummy;                    // where lanes that have nothing to publish store
    __shared__ double s_T[2][W];
    __shared__ int s_list_n;
    __shared__ int s_list_k[kWinList];    // sources that fall into this window: slot ...
    __shared__ int s_list_off[kWinList];  // ... and offset from the window start
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nx = a.nx;
    // a window can be staged asynchronously when it lies inside the mesh and everything is 16-byte aligned
    auto stageable = [&](int64_t k) {
        const int64_t w0 = k * a.own - a.lead;
        return a.aligned && w0 >= 0 && w0 + WC <= nx;
    };
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar = smem_addr(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    constexpr unsigned kRowBytes = (unsigned)(WC * sizeof(R));
    // rows of window `kw` -> shared memory (factors, and the state when with_p): lane 0 of warps 0 .. 3 issues one copy each
    // (callers: after a barrier behind the last read of the buffer)
    auto stage_window = [&](int64_t kw, bool with_p) {
        const int64_t w0 = kw * a.own - a.lead;
        if (lane != 0 || warp >= (with_p ? 4 : 3)) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before async-proxy writes
        if (warp == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((with_p ? 4u : 3u) * kRowBytes) : "memory");
        const R *src = warp == 0 ? a.fl : warp == 1 ? a.fr : warp == 2 ? a.fg : a.p_in;
        bulk_g2s(smem_addr(s_stage) + warp * kRowBytes, src + w0, kRowBytes, bar);
    };
    unsigned n_staged = 0;  // staged windows consumed so far: the mbarrier's phase
    int64_t k = blockIdx.x;
    bool staged = k < a.n_win && stageable(k);
    bool first = true;
    if (staged) stage_window(k, false);  // time-invariant inputs of my first window: may run ahead of the previous time block

    // my source row (window-independent: read once, off every window's path — it was an exposed L2 round trip per window)
    const int64_t src_mine = tid < a.n_src ? __ldg(a.src_idx + tid) : -1;

    [[maybe_unused]] int wi = -1;
    for (; k < a.n_win; k += gridDim.x) {
    ++wi;
    K3W_STAMP(0);
    const int64_t own0 = k * a.own;
    const int64_t own1 = own0 + a.own < nx ? own0 + a.own : nx;
    const int64_t win0 = own0 - a.lead;
    const int64_t cell0 = win0 + (int64_t)tid * CPT;

    // ---------------- window setup ----------------
    // Everything up to the grid-dependency wait reads only time-invariant inputs, so under
    // programmatic dependent launch the first window's setup overlaps the previous time block.
    if (tid == 0) s_list_n = 0;
    R l[CPT], r[CPT], g[CPT], x[CPT];
    if (staged) {
        mbar_wait(bar, n_staged & 1u);
        ++n_staged;
        unstage_row<R, CPT>(s_stage, tid, l);
        unstage_row<R, CPT>(s_stage + WC, tid, r);
        unstage_row<R, CPT>(s_stage + 2 * WC, tid, g);
        if (!first) unstage_row<R, CPT>(s_stage + 3 * WC, tid, x);
    } else {
        load8(a.fl, cell0, nx, (R)0, l);
        load8(a.fr, cell0, nx, (R)1, r);
        load8(a.fg, cell0, nx, (R)0, g);
    }
    __syncthreads();
    K3W_STAMP(1);
    for (int q = tid; q < a.n_src; q += T) {  // one parallel pass over the source rows
        const int64_t s = q == tid ? src_mine : a.src_idx[q];
        if (s >= 0 && s < nx && s >= win0 && s < win0 + WC) {
            const int pos = atomicAdd(&s_list_n, 1);
            if (pos < kWinList) { s_list_k[pos] = q; s_list_off[pos] = (int)(s - win0); }
        }
    }
    __syncthreads();
    const int list_n = s_list_n;
    // most windows of a long mesh hold neither a source nor a boundary row: nothing to work out for them
    const bool plain_window = list_n == 0 && !(win0 <= 0 && a.lbc != FBR_BC_NONE) && !(win0 + WC >= nx && a.rbc != FBR_BC_NONE);

    // right-hand-side overrides owned by this thread (same scheme as fbr_run_onchip.cu), by GLOBAL cell
    int ov_j = -1, ov_stride = 0;
    bool ov_slow = false;
    const double *ovp = &kZeroWin;
    if (!plain_window && cell0 < nx && cell0 + CPT > 0) {
        constexpr unsigned kNoSlot = 0xFF, kZeroSlot = 0xFE;
        uint64_t ov[CPT / 8];  // one byte per cell: the slot that overrides it
#pragma unroll
        for (int q = 0; q < CPT / 8; ++q) ov[q] = ~0ull;
        auto slot_of = [&](int j) {
            uint64_t w = ov[0];
#pragma unroll
            for (int q = 1; q < CPT / 8; ++q)
                if (q == (j >> 3)) w = ov[q];
            return (unsigned)(w >> (8 * (j & 7))) & 0xFF;
        };
        auto put = [&](int j, unsigned v) {
#pragma unroll
            for (int q = 0; q < CPT / 8; ++q)
                if (q == (j >> 3)) ov[q] = (ov[q] & ~(0xFFull << (8 * (j & 7)))) | ((uint64_t)v << (8 * (j & 7)));
        };
        if (cell0 <= 0 && a.lbc != FBR_BC_NONE) put((int)(-cell0), kZeroSlot);
        const int64_t jl = nx - 1 - cell0;
        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) put((int)jl, kZeroSlot);
        bool wide = list_n > kWinList;
        // later duplicates win (matbuilder.py:156-161): the list is unordered, so keep the largest slot
        auto claim = [&](int q, int jj) {
            if (q >= 0xFE) wide = true;
            const unsigned cur = slot_of(jj);
            if (cur >= 0xFE || (unsigned)q > cur) put(jj, (unsigned)(q & 0xFF));
        };
        if (list_n <= kWinList) {
            for (int q = 0; q < list_n; ++q) {
                const int jj = s_list_off[q] - tid * CPT;
                if (jj >= 0 && jj < CPT) claim(s_list_k[q], jj);
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < a.n_src; ++q) {
                const int64_t s = a.src_idx[q];
                const int64_t jj = s - cell0;
                if (jj >= 0 && jj < CPT && s >= 0 && s < nx) claim(q, (int)jj);
            }
        }
        int n_ov = 0;
#pragma unroll 1
        for (int j = 0; j < CPT; ++j) {
            const unsigned so = slot_of(j);
            if (so != kNoSlot) {
                ++n_ov; ov_j = j;
                if (so != kZeroSlot) { ovp = a.src_table + so; ov_stride = a.n_src; }
                else { ovp = &kZeroWin; ov_stride = 0; }
            }
        }
        if (n_ov > 1 || wide) { ov_slow = true; ov_j = -1; ovp = &kZeroWin; ov_stride = 0; }
    }

    K3W_STAMP(2);
    // scan multipliers (time-invariant within the block)
    // (fp32 mode: products of up to 512 multipliers — formed in fp64 from the multipliers the sweeps use, rounded once)
    HalfProducts hf = {1.0, 1.0}, hb = {1.0, 1.0};
    double chunk_f, chunk_b;
    if constexpr (F64) {
        hf = half_products_fwd(l); hb = half_products_bwd(g);
        chunk_f = hf.a * hf.b; chunk_b = hb.a * hb.b;
    } else {
        chunk_f = 1.0; chunk_b = 1.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) { chunk_f *= (double)l[j]; chunk_b *= (double)g[j]; }
    }
    const ScanMultipliers sm = scan_multipliers(chunk_f, chunk_b, lane);
    R Pf[NS], Pb[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { Pf[s] = (R)sm.f[s]; Pb[s] = (R)sm.b[s]; }
    const R ex_f = (R)sm.ex_f, ex_b = (R)sm.ex_b;
    if (!NEAR) {
        if (lane == 31) s_T[0][warp] = sm.tot_f;
        if (lane == 0) s_T[1][warp] = sm.tot_b;
        __syncthreads();
        if (tid < W) {
            double kk = 1.0;
            for (int v = W - 1; v >= 0; --v) {
                if (v >= tid) s_K[0][tid][v] = (R)0;
                else { s_K[0][tid][v] = (R)kk; kk *= s_T[0][v]; }
            }
            kk = 1.0;
            for (int v = 0; v < W; ++v) {
                if (v <= tid) s_K[1][tid][v] = (R)0;
                else { s_K[1][tid][v] = (R)kk; kk *= s_T[1][v]; }
            }
        }
    }
    if (tid == 0) { s_Ef[0] = (R)0; s_Eb[W] = (R)0; }
    const bool any_slow = __syncthreads_or(ov_slow);
    K3W_STAMP(3);
    if (first) {  // ---- the previous time block's output is needed from here on ----
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    if (!staged || first) load8(a.p_in, cell0, nx, (R)0, x);
    first = false;
    K3W_STAMP(4);
    __syncthreads();  // everybody has emptied the staging buffer: stream the next window in
    {
        const int64_t kn = k + gridDim.x;
        staged = kn < a.n_win && stageable(kn);
        if (staged) stage_window(kn, true);
    }
    K3W_STAMP(5);
    const int vf0 = warp - a.nbf > 0 ? warp - a.nbf : 0;          // forward carry: warps vf0 .. warp-1
    const int vb1 = warp + a.nbb < W - 1 ? warp + a.nbb : W - 1;  // backward carry: warps warp+1 .. vb1
    const unsigned st_f = smem_addr(lane == 31 ? &s_Ef[warp + 1] : &s_dummy), ld_f = smem_addr(&s_Ef[warp]);
    const unsigned st_b = smem_addr(lane == 0 ? &s_Eb[warp] : &s_dummy), ld_b = smem_addr(&s_Eb[warp + 1]);
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;
    double ov_val = a.n_steps > 0 ? __ldg(ovp) : 0.0;

    // ---------------- time loop ----------------
#pragma unroll 1
    for (int left = a.n_steps; left > 0; --left) {
        if (ov_j >= 0) {  // only the threads that own an overridden cell come here
            if constexpr (F64) {
                switch (ov_j) {
                    case 0: x[0] = ov_val; break;
                    case 1: x[1] = ov_val; break;
                    case 2: x[2] = ov_val; break;
                    case 3: x[3] = ov_val; break;
                    case 4: x[4] = ov_val; break;
                    case 5: x[5] = ov_val; break;
                    case 6: x[6] = ov_val; break;
                    default: x[7] = ov_val; break;
                }
            } else {
#pragma unroll
                for (int q = 0; q < CPT; ++q) mov_if_eq_w(x[q], (R)ov_val, ov_j, q);
            }
            if (left == 1) ov_stride = 0;
            ovp += ov_stride;
            ov_val = __ldg(ovp);  // value for the next step
        }
        if (any_slow) {
            if (ov_slow) {
                const int64_t n = a.n_steps - left;
                if (cell0 <= 0 && cell0 + CPT > 0 && a.lbc != FBR_BC_NONE) {
#pragma unroll
                    for (int q = 0; q < CPT; ++q) mov_if_eq_w(x[q], (R)0, (int)(-cell0), q);
                }
                const int64_t jl = nx - 1 - cell0;
                if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) {
#pragma unroll
                    for (int q = 0; q < CPT; ++q) mov_if_eq_w(x[q], (R)0, (int)jl, q);
                }
#pragma unroll 1
                for (int q = 0; q < a.n_src; ++q) {
                    const int64_t s = a.src_idx[q];
                    const int64_t jj = s - cell0;
                    if (jj >= 0 && jj < CPT && s >= 0 && s < nx) {
                        const R v = (R)__ldg(a.src_table + n * a.n_src + q);
#pragma unroll
                        for (int u = 0; u < CPT; ++u) mov_if_eq_w(x[u], v, (int)jj, u);
                    }
                }
            }
        }
        // ---- forward sweep: y_i = b_i - l_i y_{i-1}, then t_i = r_i y_i ----
        {
            [[maybe_unused]] double ya;
            R y;
            if constexpr (F64) y = chunk_end_fwd(x, l, hf.b, ya);
            else y = chunk_end_fwd_plain<R, CPT>(x, l);
#pragma unroll
            for (int s = 0; s < NS; ++s) y = fma_any<R>(Pf[s], __shfl_up_sync(kAllLanes, y, 1 << s), y);
            const R carry = and_mask(__shfl_up_sync(kAllLanes, y, 1), keep_f);
            sts_any(st_f, y);
            __syncthreads();
            R c;
            if (NEAR) {
                c = lds_any(ld_f, R());
            } else {
                c = (R)0;
                for (int v = vf0; v < warp; ++v) c = fma_any<R>(s_K[0][warp][v], s_Ef[v + 1], c);
            }
            if constexpr (F64) chunk_apply_fwd(x, l, r, hf.a, ya, fma(ex_f, c, carry));
            else chunk_apply_fwd_plain<R, CPT>(x, l, r, fma_any<R>(ex_f, c, carry));
        }
        // ---- backward sweep: x_i = t_i - g_i x_{i+1} ----
        {
            [[maybe_unused]] double za;
            R z;
            if constexpr (F64) z = chunk_end_bwd(x, g, hb.b, za);
            else z = chunk_end_bwd_plain<R, CPT>(x, g);
#pragma unroll
            for (int s = 0; s < NS; ++s) z = fma_any<R>(Pb[s], __shfl_down_sync(kAllLanes, z, 1 << s), z);
            const R carry = and_mask(__shfl_down_sync(kAllLanes, z, 1), keep_b);
            sts_any(st_b, z);
            __syncthreads();
            R c;
            if (NEAR) {
                c = lds_any(ld_b, R());
            } else {
                c = (R)0;
                for (int v = vb1; v > warp; --v) c = fma_any<R>(s_K[1][warp][v], s_Eb[v], c);
            }
            if constexpr (F64) chunk_apply_bwd(x, g, hb.a, za, fma(ex_b, c, carry));
            else chunk_apply_bwd_plain<R, CPT>(x, g, fma_any<R>(ex_b, c, carry));
        }
    }

    K3W_STAMP(6);
    // ---------------- write back the owned cells ----------------
    // Whole windows inside the mesh leave through shared memory and ONE bulk store (cp.async.bulk shared -> global) of the
    // owned range: 512 threads issuing four 16-byte stores each held the window for 1 400 cycles (scratch/k3w_prof.py);
    // the bulk store drains while the next window is being set up.
    R *dst = a.p_out + cell0;
    const bool bulk_out = a.aligned && win0 >= 0 && win0 + WC <= nx && own1 - own0 == a.own &&
                          (reinterpret_cast<uintptr_t>(a.p_out + own0) & 15) == 0 && (a.own * sizeof(R)) % 16 == 0;
    if (bulk_out) {  // (CTA-uniform)
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous window's store has read s_out
        __syncthreads();
        stage_row_out<R, CPT>(s_out, tid, x);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my generic-proxy writes -> visible to the async proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.p_out + own0),
                         "r"(smem_addr(s_out + (own0 - win0))), "r"((unsigned)(a.own * sizeof(R))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    } else if (cell0 >= own0 && cell0 + CPT <= own1 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        if constexpr (F64) {
            double2 *q = reinterpret_cast<double2 *>(dst);
#pragma unroll
            for (int j = 0; j < CPT / 2; ++j) q[j] = make_double2(x[2 * j], x[2 * j + 1]);
        } else {
            float4 *q = reinterpret_cast<float4 *>(dst);
#pragma unroll
            for (int j = 0; j < CPT / 4; ++j) q[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int64_t i = cell0 + j;
            if (i >= own0 && i < own1) dst[j] = x[j];
        }
    }
    K3W_STAMP(7);
    }  // windows of this CTA
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // my bulk stores are complete before the CTA retires
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

// v >= 0: bit patterns are ordered.  The maxima settle after the first few warps; everybody else only LOOKS (one L2
// read) instead of queueing an atomic on the same four addresses (312 500 warps at 1e7 cells: 1.3 ms -> 0.1 ms).
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (bits > __ldcg(reinterpret_cast<const unsigned long long *>(addr))) atomicMax(reinterpret_cast<unsigned long long *>(addr), bits);
}
__device__ __forceinline__ void atomic_max_seen(unsigned long long *addr, unsigned long long v) {
    if (v > __ldcg(addr)) atomicMax(addr, v);
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
        atomic_max_seen(horizon, (unsigned long long)nf);
        atomic_max_seen(horizon + 1, (unsigned long long)nb);
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
static bool plan_window(int64_t nx, int64_t Hf, int64_t Hb, int64_t max_s, int sm_count, WinPlan *best, int cpt = kWinCPT) {
    const int64_t hf = round_up8(Hf > 0 ? Hf : 1), hb = round_up8(Hb > 0 ? Hb : 1);
    int force_w = 0;
    int64_t force_s = 0;
    if (const char *e = getenv("FBR_WIN_W")) force_w = atoi(e);
    if (const char *e = getenv("FBR_WIN_S")) force_s = atoll(e);
    if (force_s > max_s) force_s = max_s;
    bool found = false;
    for (int W : {8, 16}) {
        if (force_w && W != force_w) continue;
        const int64_t wc = 32 * (int64_t)W * cpt;
        const int resident = (W >= 16 ? 1 : 2) * sm_count;
        // us per window: fixed (load / setup / store) + per time step; measured on B200 (C2 / C5 sweeps,
        // profiles/r01_k3w_sweep.txt).  Dependent launches hide the launch gap.
        // (round 2, after the bulk-copy staging, fitted to the (W, s) sweeps of C2 and C5 — profiles/r02_k3w_sweep_v2.txt: a
        // window costs 2.9 / 1.5 us + 0.57 / 0.52 us per step (16 / 8 warps); two 8-warp CTAs sharing an SM step in 0.64 us
        // each, so C2 prefers 143 windows on 148 SMs to 241 on 296 slots)
        const double t_fixed = W >= 16 ? 2.9 : 1.5, t_launch = 0.3;
        for (int64_t s = 1; s <= max_s; ++s) {
            if (force_s && s != force_s) continue;
            const int64_t own = wc - s * (hf + hb);
            if (own < wc / 8) break;
            const int64_t n_win = (nx + own - 1) / own;
            const double t_step = W >= 16 ? 0.57 : n_win <= sm_count ? 0.52 : 0.64;
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

extern "C" int fbr_run_window_plan_f32(int64_t nx, int64_t horizon_fwd, int64_t horizon_bwd, int64_t snapshot_every,
                                       int64_t *plan4) {
    FBR_REQUIRE(nx >= 1 && horizon_fwd >= 0 && horizon_bwd >= 0 && plan4, "bad window-plan arguments");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    WinPlan pl;
    const int64_t max_s = snapshot_every >= 1 ? (snapshot_every < 4096 ? snapshot_every : 4096) : 4096;
    if (!plan_window(nx, horizon_fwd, horizon_bwd, max_s, f.sm_count, &pl, 16))
        return fail(FBR_ERR_UNSUPPORTED, "decay horizon (%lld, %lld) is too long for the windowed path",
                    (long long)horizon_fwd, (long long)horizon_bwd);
    plan4[0] = pl.W; plan4[1] = pl.s; plan4[2] = pl.own; plan4[3] = pl.n_win;
    return FBR_OK;
}

extern "C" int64_t fbr_run_window_work_bytes(int64_t nx) { return 2 * nx * (int64_t)sizeof(double); }
// fp32 mode: the three factors and the initial state narrowed once + two state buffers, each padded to 16 bytes
static inline int64_t pitch_f32(int64_t nx) { return (nx + 3) / 4 * 4; }
extern "C" int64_t fbr_run_window_f32_work_bytes(int64_t nx) { return 6 * pitch_f32(nx) * (int64_t)sizeof(float); }

namespace fbr {
__global__ void narrow4_kernel(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                               const double *__restrict__ d, float *__restrict__ out, int64_t nx, int64_t pitch) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (int64_t)gridDim.x * blockDim.x) {
        out[i] = (float)__ldg(a + i);
        out[pitch + i] = (float)__ldg(b + i);
        out[2 * pitch + i] = (float)__ldg(c + i);
        out[3 * pitch + i] = (float)__ldg(d + i);
    }
}
}  // namespace fbr

// fl, fr, fg, p0, the two state buffers: arrays of R (fp32 mode: narrowed by the caller below)
template <typename R>
static int run_window_impl(const R *fl, const R *fr, const R *fg, const R *p0, int64_t nx,
                                  int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                  const double *src_table, int64_t snapshot_every, int64_t snap_row_stride, R *snap,
                                  R *p_final, int64_t horizon_fwd, int64_t horizon_bwd, R *buf0, R *buf1, void *stream) {
    constexpr int kCPT = 64 / (int)sizeof(R);
    FBR_REQUIRE(nx >= 1 && n_steps >= 0, "bad nx / n_steps");
    FBR_REQUIRE(fl && fr && fg && p0 && buf0 && buf1, "NULL factor / initial-state / work pointer");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_table)), "sources given without src_idx / src_table");
    FBR_REQUIRE(snapshot_every >= 0 && (!snap || (snapshot_every >= 1 && snap_row_stride >= nx)), "bad snapshot layout");
    FBR_REQUIRE(horizon_fwd >= 0 && horizon_bwd >= 0, "bad decay horizon");
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    cudaStream_t st = as_stream(stream);
    WinPlan pl;
    // snapshot_every >= 1 caps the block length even without a snapshot buffer (a caller that streams
    // rows out itself, one call per interval, then gets the same window geometry — the same bits — as
    // the single-call path); 0 = no cap
    const int64_t max_s = snapshot_every >= 1 ? (snapshot_every < 4096 ? snapshot_every : 4096) : 4096;
    if (!plan_window(nx, horizon_fwd, horizon_bwd, max_s, f.sm_count, &pl, kCPT))
        return fail(FBR_ERR_UNSUPPORTED, "decay horizon (%lld, %lld) is too long for the windowed path",
                    (long long)horizon_fwd, (long long)horizon_bwd);
    // equalise the blocks inside one snapshot interval
    int64_t s_blk = pl.s;
    if (snapshot_every >= 1) {
        const int64_t nblk = (snapshot_every + s_blk - 1) / s_blk;
        s_blk = (snapshot_every + nblk - 1) / nblk;
    }
    const int64_t warp_span = 32 * kCPT;
    WinArgsT<R> a;
    a.fl = fl; a.fr = fr; a.fg = fg; a.nx = nx; a.own = pl.own; a.lead = pl.lead; a.lbc = lbc; a.rbc = rbc;
    a.src_idx = src_idx; a.n_src = n_src;
    a.nbf = (int)((horizon_fwd + warp_span - 1) / warp_span); if (a.nbf < 1) a.nbf = 1;
    a.nbb = (int)((horizon_bwd + warp_span - 1) / warp_span); if (a.nbb < 1) a.nbb = 1;
    const bool no_pdl = getenv("FBR_WIN_NO_PDL") != nullptr;  // tuning aid
    a.n_win = pl.n_win;
    const size_t smem = 5 * (size_t)(32 * pl.W) * 64;  // staging: l, r, g, p of the next window + the state going out, 64 bytes per thread each
    // kernel variant: how many scan stages the horizon needs, and whether carries stop at the neighbouring warp
    const int64_t hmax = horizon_fwd > horizon_bwd ? horizon_fwd : horizon_bwd;
    const bool near = a.nbf == 1 && a.nbb == 1 && !getenv("FBR_WIN_GENERAL");
    // (2^ns lanes of kCPT cells each must span the horizon: 64 / 128 cells in fp64, 128 / 256 in the fp32 mode)
    const int ns = !near ? 5 : hmax <= 8 * kCPT ? 3 : hmax <= 16 * kCPT ? 4 : 5;
    void (*kern)(const WinArgsT<R>) = nullptr;
    if (pl.W >= 16) kern = !near ? run_window_kernel<R, 16, 5, false> : ns == 3 ? run_window_kernel<R, 16, 3, true>
                           : ns == 4 ? run_window_kernel<R, 16, 4, true> : run_window_kernel<R, 16, 5, true>;
    else kern = !near ? run_window_kernel<R, 8, 5, false> : ns == 3 ? run_window_kernel<R, 8, 3, true>
                : ns == 4 ? run_window_kernel<R, 8, 4, true> : run_window_kernel<R, 8, 5, true>;
    FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t resident = (int64_t)(pl.W >= 16 ? 1 : 2) * f.sm_count;  // persistent CTAs, one wave
    R *buf[2] = {buf0, buf1};
    const R *cur = p0;
    int64_t n = 0, snap_row = 0;
    int which = 0;
    while (n < n_steps) {
        int64_t len = s_blk;
        if (snapshot_every >= 1) {
            const int64_t to_snap = snapshot_every - n % snapshot_every;
            if (len > to_snap) len = to_snap;
        }
        if (len > n_steps - n) len = n_steps - n;
        const bool at_snap = snap && (n + len) % snapshot_every == 0;
        const bool at_end = n + len == n_steps;
        R *out;
        if (at_snap) out = snap + snap_row * snap_row_stride;
        else if (at_end && p_final) out = p_final;
        else { out = buf[which]; which ^= 1; }
        a.p_in = cur; a.p_out = out; a.n_steps = (int)len;
        a.src_table = src_table ? src_table + n * n_src : nullptr;
        // every block but the first is a programmatic dependent launch: its factor loads and setup
        // overlap the previous block's tail (the first one must see the caller's earlier work whole)
        auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        // asynchronous staging pays when a CTA walks several windows; a lone window loads directly
        a.aligned = pl.n_win > resident && al16(fl) && al16(fr) && al16(fg) && al16(cur);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(pl.n_win < resident ? pl.n_win : resident));
        cfg.blockDim = dim3(pl.W >= 16 ? 512 : 256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (n > 0 && !no_pdl) ? 1 : 0;
        FBR_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
        if (int rc = check_launch("fbr_run_window_f64")) return rc;
        if (at_snap) ++snap_row;
        if (at_end && at_snap && p_final)
            FBR_CUDA(cudaMemcpyAsync(p_final, out, nx * sizeof(R), cudaMemcpyDeviceToDevice, st));
        cur = out;
        n += len;
    }
    if (n_steps == 0 && p_final) FBR_CUDA(cudaMemcpyAsync(p_final, p0, nx * sizeof(R), cudaMemcpyDeviceToDevice, st));
    return FBR_OK;
}

extern "C" int fbr_run_window_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                  int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                  const double *src_table, int64_t snapshot_every, int64_t snap_row_stride, double *snap,
                                  double *p_final, int64_t horizon_fwd, int64_t horizon_bwd, double *work, void *stream) {
    FBR_REQUIRE(nx >= 1 && work, "bad nx / NULL work pointer");
    return run_window_impl<double>(fl, fr, fg, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every,
                                   snap_row_stride, snap, p_final, horizon_fwd, horizon_bwd, work, work + nx, stream);
}

extern "C" int fbr_run_window_f32(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                  int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                  const double *src_table, int64_t snapshot_every, int64_t snap_row_stride, float *snap,
                                  float *p_final, int64_t horizon_fwd, int64_t horizon_bwd, void *work, void *stream) {
    FBR_REQUIRE(nx >= 1 && fl && fr && fg && p0 && work, "bad nx / NULL factor / initial-state / work pointer");
    FBR_REQUIRE((reinterpret_cast<uintptr_t>(work) & 15) == 0, "work must be 16-byte aligned");
    const int64_t pitch = pitch_f32(nx);
    float *w = static_cast<float *>(work);
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    const int64_t want = (nx + 255) / 256, cap = (int64_t)f.sm_count * 16;
    narrow4_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(fl, fr, fg, p0, w, nx, pitch);
    if (int rc = check_launch("fbr_run_window_f32 (narrow)")) return rc;
    return run_window_impl<float>(w, w + pitch, w + 2 * pitch, w + 3 * pitch, nx, n_steps, lbc, rbc, src_idx, n_src, src_table,
                                  snapshot_every, snap_row_stride, snap, p_final, horizon_fwd, horizon_bwd, w + 4 * pitch,
                                  w + 5 * pitch, stream);
}

#ifdef FBR_K3W_PROF
extern "C" int fbr_debug_window_prof(long long *host) {
    return (int)cudaMemcpyFromSymbol(host, fbr::g_k3w_prof, sizeof(fbr::g_k3w_prof));
}
#endif
