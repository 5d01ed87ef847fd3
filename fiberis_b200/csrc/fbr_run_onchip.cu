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
// Three kernels share this scheme: run_onchipx_kernel (64 bytes of state per thread, per-step scalars staged through
// shared memory; with ZONAL the matrices are assembled and factorised in the per-member prologue), run_group_kernel
// (short meshes: groups of 2..16 lanes per member, several members per warp) and the generic run_onchip_kernel
// (1..8 cells per thread; single short models and the fp32 mode below 512 cells), whose per-step scalars are
// prefetched one step ahead straight from global memory (L2) by the thread that owns the cell.
#include "fbr_common.cuh"
#include "fbr_sweep.cuh"
#include "fbr_factor_reg.cuh"

#include <atomic>
#include <cmath>
#include <cstring>
#include <type_traits>

namespace fbr {

constexpr int kMaxWarps = 16;  // per team
constexpr int kMaxSrc = 253;  // a source column is one byte of a descriptor (0xFE / 0xFF are reserved)
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
    int64_t snap_row_stride, snap_model_stride;  // in elements
    void *snap;     // element (row, slot, i) at row * snap_row_stride + slot * snap_model_stride + i
    void *p_final;  // (M, nx) of T
    const int64_t *obs_idx;
    int n_obs;
    const double *obs;    // (n_steps + 1, n_obs)
    const double *obs_w;  // (n_obs) or NULL
    double *misfit;       // (M)
    int blk_steps;        // fp64 x8 kernel: time steps per staged block of per-step scalars
    int prefetch_rows;    // fp64 x8 kernel, short runs: the NEXT member's four rows stream into shared memory (bulk copies)
    // on-chip assembly mode (fbr_run_zonal_f64): no factor arrays, the members' matrices are assembled (K1 arithmetic)
    // and factorised in registers from the shared mesh and n_zones diffusivities per member
    const double *z_mesh;        // (nx)
    const double *z_value;       // (M, z_n)
    const int32_t *z_of_cell;    // (nx)
    int z_n;
    double z_dt;
    const double *z_sigma;       // (nx) PML damping or NULL
    int32_t *z_singular;         // (M) preset to INT32_MAX by the caller; atomicMin of 1 + zero-pivot row; may be NULL
    // fbr_run_opts
    int64_t snap_cell_stride;    // element stride between the cells of a recorded row (1: contiguous)
    double scan_eps;             // skip scan stages / carry chains whose multiplier products are all below this; < 0: never
    int reserve_ctas;
    int64_t n_samples;           // 0: observations on the step times
    const int32_t *sample_row;
    const double *sample_w;
    const int32_t *row_first_sample;
    int base_sample, first_sample, last_sample;
    int32_t *scan_mode;          // (M) or NULL: how each member was scanned
    unsigned long long *ticket;  // zeroed word of this launch, or NULL: members beyond the first deal are handed out on demand
    int snap_wide;               // fp32 mode, a handful of recorded members: `snap` holds doubles, rows are widened at the store
    const int32_t *order;        // (M) or NULL: with a ticket, the k-th member handed out is order[k] (fbr_run_opts::member_order)
};

// Rows of A(dt) for CPT consecutive cells of member m, exactly K1's arithmetic (fbr_assemble.cu, matbuilder.py:21-81).
template <int CPT>
__device__ __forceinline__ void assemble_zonal_chunk(const RunArgs &a, int64_t m, int cell0, unsigned src_mask, double (&dl)[CPT],
                                                     double (&d)[CPT], double (&du)[CPT]) {
    const int nx = (int)a.nx;
    const double *zv = a.z_value + m * a.z_n;
    double xs[CPT + 2], Ds[CPT + 2];
#pragma unroll
    for (int j = 0; j < CPT + 2; ++j) {
        int i = cell0 - 1 + j;
        i = i < 0 ? 0 : (i > nx - 1 ? nx - 1 : i);
        xs[j] = __ldg(a.z_mesh + i);
        Ds[j] = __ldg(zv + (a.z_of_cell ? __ldg(a.z_of_cell + i) : i));  // zone value, or the member's own (nx) row
    }
    double he[CPT + 1];  // harmonic means between neighbours: he[j] couples cells cell0 + j - 1 and cell0 + j
#pragma unroll
    for (int j = 0; j < CPT + 1; ++j)
        he[j] = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, Ds[j]), Ds[j + 1]), __dadd_rn(Ds[j], Ds[j + 1]));
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int i = cell0 + j;
        double aa = 0.0, bb = 0.0, cc = 0.0;
        if (i > 0 && i < nx - 1) {
            const double dxm = __dsub_rn(xs[j + 1], xs[j]), dxp = __dsub_rn(xs[j + 2], xs[j + 1]);
            const double sum = __dadd_rn(dxm, dxp);
            const double al = __ddiv_rn(__dmul_rn(he[j], a.z_dt), __dmul_rn(__dmul_rn(dxm, sum), 0.5));
            const double ar = __ddiv_rn(__dmul_rn(he[j + 1], a.z_dt), __dmul_rn(__dmul_rn(dxp, sum), 0.5));
            aa = -al;
            bb = __dadd_rn(__dadd_rn(1.0, al), ar);
            cc = -ar;
        }
        if (i == 0) {
            if (a.lbc == FBR_BC_DIRICHLET) bb = 1.0;
            else if (a.lbc == FBR_BC_NEUMANN) { bb = -1.0; cc = 1.0; }
        }
        if (i == nx - 1) {
            if (a.rbc == FBR_BC_DIRICHLET) bb = 1.0;
            else if (a.rbc == FBR_BC_NEUMANN) { bb = -1.0; aa = 1.0; }
        }
        if ((src_mask >> j) & 1) { aa = 0.0; bb = 1.0; cc = 0.0; }
        if (a.z_sigma && i < nx) bb = __dadd_rn(bb, __dmul_rn(a.z_dt, __ldg(a.z_sigma + i)));
        if (i == 0) aa = 0.0;
        if (i >= nx - 1) cc = 0.0;
        if (i >= nx) { aa = 0.0; bb = 1.0; }  // identity rows behind the mesh
        dl[j] = aa; d[j] = bb; du[j] = cc;
    }
}

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// x[j] = v for a run-time j, in place, without dynamic register indexing: one compare and one
// predicated move per owned cell (inline PTX so that ptxas keeps x[] where it is).
__device__ __forceinline__ void mov_if_eq(double &dst, double v, int j, int k) {
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, %3; @p mov.f64 %0, %1; }" : "+d"(dst) : "d"(v), "r"(j), "r"(k));
}
__device__ __forceinline__ void mov_if_eq(float &dst, float v, int j, int k) {
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, %3; @p mov.f32 %0, %1; }" : "+f"(dst) : "f"(v), "r"(j), "r"(k));
}
template <typename T, int CPT>
__device__ __forceinline__ void set_at(T (&x)[CPT], int j, T v) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) mov_if_eq(x[k], v, j, k);
}
// v = x[j] for a run-time j: binary select tree on the bits of j
template <typename T, int CPT>
__device__ __forceinline__ T get_at(const T (&x)[CPT], int j) {
    if (CPT == 1) return x[0];
    const bool b0 = j & 1, b1 = j & 2, b2 = j & 4;
    T a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = (2 * k + 1 < CPT) ? (b0 ? x[(2 * k + 1) % CPT] : x[(2 * k) % CPT]) : x[(2 * k) % CPT];
    if (CPT == 2) return a[0];
    const T c0 = b1 ? a[1] : a[0], c1 = b1 ? a[3] : a[2];
    if (CPT == 4) return c0;
    return b2 ? c1 : c0;
}
// if (cond) c = fma(t, c, e): predicated FMA, no select
__device__ __forceinline__ void fma_if(double &c, double t, double e, bool cond) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p fma.rn.f64 %0, %1, %0, %2; }" : "+d"(c) : "d"(t), "d"(e), "r"((int)cond));
}
__device__ __forceinline__ void fma_if(float &c, float t, float e, bool cond) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p fma.rn.f32 %0, %1, %0, %2; }" : "+f"(c) : "f"(t), "f"(e), "r"((int)cond));
}

__device__ const double kZero = 0.0;

constexpr unsigned kSlotNone = 0xFF, kSlotZero = 0xFE;

// ticket words of the on-demand member schedule: one per launch in flight, zeroed on the launch's stream
constexpr unsigned kTicketSlots = 1024;
__device__ unsigned long long g_k3_ticket[kTicketSlots];

// Order in which a CTA walks the models.  A recorded model (snapshots every few steps) takes about
// twice as long as the others, so a recorded model picked up in the last round would be the tail of
// the whole launch (C3: +3.5 ms of 27).  With A recorded models (short lists only) and G CTAs:
//   round 0 : the A CTAs at the END of the grid integrate one recorded model each, the other G - A
//             CTAs take the first G - A unrecorded models;
//   later   : unrecorded models are dealt G at a time, the A "recorded" CTAs LAST in every deal, so
//             they are the ones that go without when the final deal is short.
// preset: snapshot slot of the model (-1 none), or -2 = look it up (long lists, no reordering).
struct ModelCursor {
    int64_t n_rec, j_rec, k, p, n_plain;
    bool round0;
};
__device__ __forceinline__ bool recorded_before(const RunArgs &a, int64_t q) {  // duplicate of an earlier entry?
    for (int64_t e = 0; e < q; ++e)
        if (a.snap_models[e] == a.snap_models[q]) return true;
    return false;
}
__device__ __forceinline__ ModelCursor model_cursor(const RunArgs &a) {
    ModelCursor c;
    const int64_t G = gridDim.x, cp = G - 1 - blockIdx.x;
    c.n_rec = (a.snap && a.snap_models && a.n_snap_models <= 16 && a.n_snap_models < G) ? a.n_snap_models : 0;
    if (c.n_rec == 0) {
        c.j_rec = 0; c.k = blockIdx.x; c.p = 0; c.round0 = true; c.n_plain = a.M;
        return c;
    }
    int64_t distinct = 0;
    for (int64_t q = 0; q < c.n_rec; ++q) distinct += !recorded_before(a, q);
    c.n_plain = a.M - distinct;
    c.j_rec = cp < c.n_rec ? cp : c.n_rec;
    c.round0 = cp >= c.n_rec;
    c.p = c.round0 ? cp - c.n_rec : G - c.n_rec + cp;
    c.k = c.round0 ? c.p : (G - c.n_rec) + c.p;
    return c;
}
__device__ __forceinline__ bool next_model(const RunArgs &a, ModelCursor &c, int64_t &m, long long &preset) {
    const int64_t G = gridDim.x;
    if (c.n_rec == 0) {
        if (a.ticket && !c.round0) {
            // Members cost differently (4 / 5 scan stages, the warp table: 940 / 1025 / 1150 cycles per step on C3), so a
            // fixed deal leaves the CTAs that drew the expensive ones as the tail of the launch: after the first deal a
            // CTA takes the next member when it is free.  Which CTA integrates a member does not change a bit of it.
            __shared__ unsigned long long s_next;
            __syncthreads();
            if (threadIdx.x == 0) s_next = (unsigned long long)G + atomicAdd(a.ticket, 1ULL);
            __syncthreads();
            c.k = (int64_t)s_next;
        }
        if (c.k >= a.M) return false;
        m = (a.ticket && a.order) ? (int64_t)__ldg(a.order + c.k) : c.k;
        c.k += G;
        c.round0 = false;
        preset = !a.snap ? -1 : !a.snap_models ? (long long)m : -2;
        return true;
    }
    if (c.j_rec < c.n_rec) {
        m = a.snap_models[c.j_rec];
        preset = c.j_rec;
        c.j_rec = c.n_rec;
        return true;
    }
    if (c.k >= c.n_plain) return false;
    // the k-th model that is not recorded
    m = c.k;
    for (;;) {
        int64_t below = 0;
        for (int64_t q = 0; q < c.n_rec; ++q) below += (a.snap_models[q] <= m) && !recorded_before(a, q);
        if (c.k + below == m) break;
        m = c.k + below;
    }
    c.k = c.round0 ? (G - c.n_rec) + c.p : c.k + G;
    c.round0 = false;
    preset = -1;
    return true;
}

// WIDE (fp32 mode, a few short models): the state is kept and written in float, but factors and arithmetic stay in
// fp64 — ONE model has no throughput reason to round its factors, and on a uniform mesh every cell's l, r, g round to
// fp32 the same way: the bias of the conserved mean is then systematic (C1, 1000 steps: 1.04e-5 with float factors
// whatever the chunking; the fp32-rounded STATE alone stays near 1e-6).
template <typename T, int CPT, int W, bool ZONAL, bool WIDE>
__global__ void __launch_bounds__(32 * W, (16 / W) > 0 ? (16 / W) : 1) run_onchip_kernel(const RunArgs a) {
    static_assert(CPT <= 8, "slot bytes are packed into one 64-bit word");
    using A = typename std::conditional<WIDE, double, T>::type;  // arithmetic / factor type
    // cross-warp carry: C_w = sum_v K[dir][w][v] * E[dir][v], K = products of the warp-total
    // multipliers between warp v and warp w (time-invariant, zero where warp v does not feed warp w),
    // E = this step's zero-carry warp-end values
    __shared__ __align__(16) A s_K[2][W][W];
    __shared__ __align__(16) A s_E[2][W];
    __shared__ double s_T[2][W];
    __shared__ double s_red[W];
    __shared__ long long s_snap_slot;
    __shared__ Mat2 s_zw[ZONAL ? W : 1];
    __shared__ double s_zc[ZONAL ? W : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nx = a.nx;
    const int64_t cell0 = (int64_t)tid * CPT;
    unsigned src_rows = 0;  // which of my cells are source rows of the matrix (on-chip assembly mode)
    if (ZONAL) {
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {
            const int64_t jj = a.src_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) src_rows |= 1u << jj;
        }
    }

    ModelCursor cursor = model_cursor(a);
    int64_t m;
    long long preset;
    while (next_model(a, cursor, m, preset)) {
        // ---------------- per-model setup ----------------
        A l[CPT], r[CPT], g[CPT], x[CPT];
        double ld[CPT], gd[CPT];  // the multipliers in fp64: the scan multipliers are products of these, rounded once
        if constexpr (ZONAL) {
            // on-chip assembly (K1's arithmetic, bit-identical diagonals) + factorisation in registers: a single short
            // model needs ONE launch instead of three (C1: the one-thread factorisation alone took 44 us)
            double zr[CPT];
            assemble_zonal_chunk<CPT>(a, m, (int)cell0, src_rows, ld, zr, gd);
            const int64_t left = nx - cell0;
            const int bad = factor_rows_in_registers<CPT, W>(ld, zr, gd, (int)cell0, (int)(left < 0 ? 0 : left > CPT ? CPT : left),
                                                             lane, warp, s_zw, s_zc);
            if (bad && a.z_singular) atomicMin(a.z_singular + m, bad);
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int64_t i = cell0 + j;
                l[j] = (A)ld[j]; r[j] = (A)zr[j]; g[j] = (A)gd[j];
                x[j] = i < nx ? (A)__ldg(a.p0 + m * a.p0_stride + i) : (A)0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int64_t i = cell0 + j;
                if (i < nx) {
                    ld[j] = __ldg(a.fl + m * nx + i);
                    gd[j] = __ldg(a.fg + m * nx + i);
                    l[j] = (A)ld[j];
                    r[j] = (A)__ldg(a.fr + m * nx + i);
                    g[j] = (A)gd[j];
                    x[j] = (A)__ldg(a.p0 + m * a.p0_stride + i);
                } else {  // padding cell: identity row, decoupled
                    ld[j] = 0.0; gd[j] = 0.0;
                    l[j] = (A)0; r[j] = (A)1; g[j] = (A)0; x[j] = (A)0;
                }
            }
        }
        // Which of my cells are overridden in the right-hand side / observed?  The common case — at
        // most one of each per thread — is a register index (ov_j / ob_j, -1 = none) plus a pointer
        // that walks down a column of src_table / obs; a boundary row reads a constant zero with
        // stride 0.  Threads owning several such cells take a slow generic path.
        int ov_j = -1, ob_j = -1, ov_stride = 0, ob_stride = 0;
        bool ov_slow = false, ob_slow = false;
        const double *ovp = &kZero, *obp = &kZero;
        if (cell0 < nx) {
            uint64_t ov = ~0ull;  // one slot byte per owned cell
            if (cell0 == 0 && a.lbc != FBR_BC_NONE) ov = (ov & ~0xFFull) | kSlotZero;
            const int64_t jl = nx - 1 - cell0;
            if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE)
                ov = (ov & ~(0xFFull << (8 * jl))) | ((uint64_t)kSlotZero << (8 * jl));
#pragma unroll 1
            for (int k = 0; k < a.n_src; ++k) {  // ascending k: later duplicates win (matbuilder.py:156-161)
                const int64_t jj = a.src_idx[k] - cell0;
                if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) ov = (ov & ~(0xFFull << (8 * jj))) | ((uint64_t)k << (8 * jj));
            }
            int n_ov = 0, n_ob = 0;
#pragma unroll 1
            for (int j = 0; j < CPT; ++j) {
                const unsigned so = (unsigned)(ov >> (8 * j)) & 0xFF;
                if (so != kSlotNone) {
                    ++n_ov; ov_j = j;
                    if (so != kSlotZero) { ovp = a.src_table + so; ov_stride = a.n_src; }
                    else { ovp = &kZero; ov_stride = 0; }
                }
            }
#pragma unroll 1
            for (int k = 0; k < a.n_obs; ++k) {
                const int64_t jj = a.obs_idx[k] - cell0;
                if (jj >= 0 && jj < CPT && a.obs_idx[k] < nx) {
                    ++n_ob; ob_j = (int)jj; ob_stride = a.n_obs;
                    obp = a.obs + a.n_obs + k;  // row 1 = state after the first step
                }
            }
            if (n_ov > 1) { ov_slow = true; ov_j = -1; ovp = &kZero; ov_stride = 0; }
            if (n_ob > 1) { ob_slow = true; ob_j = -1; obp = &kZero; ob_stride = 0; }
        }
        const double ob_w = (ob_j >= 0 && a.obs_w) ? a.obs_w[(obp - a.obs) % a.n_obs] : 1.0;

        // scan multipliers (time-invariant): stage products of the chunk multipliers; lanes that a
        // stage does not touch get a zero multiplier so the time loop needs no predicates.  The products are
        // formed in fp64 from the fp64 factors and rounded ONCE (fp32 mode: a product of up to 256 rounded
        // multipliers would otherwise carry their accumulated rounding into every carry of every step).
        A PfL[5], PbL[5], PfEx, PbEx;
        {
            double pf = 1.0, pb = 1.0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) { pf *= -ld[j]; pb *= -gd[j]; }
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const bool fa = lane >= (1 << s), ba = lane + (1 << s) < 32;
                PfL[s] = fa ? (A)pf : (A)0;
                PbL[s] = ba ? (A)pb : (A)0;
                const double upf = __shfl_up_sync(kFull, pf, 1 << s);
                const double dnb = __shfl_down_sync(kFull, pb, 1 << s);
                if (fa) pf *= upf;
                if (ba) pb *= dnb;
            }
            // pf = product over lanes 0..lane, pb = product over lanes lane..31
            const double ex_f = __shfl_up_sync(kFull, pf, 1), ex_b = __shfl_down_sync(kFull, pb, 1);
            PfEx = lane == 0 ? (A)1 : (A)ex_f;
            PbEx = lane == 31 ? (A)1 : (A)ex_b;
            if (W > 1) {
                __syncthreads();  // previous model's readers are done with the shared scratch
                if (lane == 31) s_T[0][warp] = pf;
                if (lane == 0) s_T[1][warp] = pb;
                __syncthreads();
                if (tid < W) {  // row tid of both tables
                    double k = 1.0;
                    for (int v = W - 1; v >= 0; --v) {  // forward: warps v < tid feed warp tid
                        if (v >= tid) s_K[0][tid][v] = (A)0;
                        else { s_K[0][tid][v] = (A)k; k *= s_T[0][v]; }
                    }
                    k = 1.0;
                    for (int v = 0; v < W; ++v) {  // backward: warps v > tid feed warp tid
                        if (v <= tid) s_K[1][tid][v] = (A)0;
                        else { s_K[1][tid][v] = (A)k; k *= s_T[1][v]; }
                    }
                }
            }
        }
        // snapshot slot of this model (-1: not recorded)
        if (tid == 0) {
            long long slot = preset;
            if (preset == -2) {
                slot = -1;
                for (int64_t q = 0; q < a.n_snap_models; ++q)
                    if (a.snap_models[q] == m) slot = q;
            }
            s_snap_slot = slot;
        }
        __syncthreads();
        const long long snap_slot = s_snap_slot;
        double acc = 0.0;
        const int snap_every = a.snapshot_every > 0x7fffffff ? 0x7fffffff : (int)a.snapshot_every;
        int snap_countdown = snap_slot >= 0 ? snap_every : 0x7fffffff;  // never reaches 0 when not recorded
        T *snap_dst = reinterpret_cast<T *>(a.snap) + (snap_slot >= 0 ? snap_slot : 0) * a.snap_model_stride + cell0 * a.snap_cell_stride;
        const int64_t snap_stride = a.snap_row_stride;
        // per-step scalars are prefetched one step ahead straight from global memory (L2) by the
        // threads that own an overridden / observed cell; everybody else re-reads a constant
        double ov_val = a.n_steps > 0 ? __ldg(ovp) : 0.0;
        double ob_val = a.n_steps > 0 ? __ldg(obp) : 0.0;

        // ---------------- time loop ----------------
#pragma unroll 1
        for (int left = (int)a.n_steps; left > 0; --left) {
            if (left == 1) { ov_stride = 0; ob_stride = 0; }  // no row after the last one
            ovp += ov_stride;
            obp += ob_stride;
            const double ov_next = __ldg(ovp), ob_next = __ldg(obp);

            // right-hand side: b = p^n with boundary / source overrides (matbuilder.py:45-76)
            if (ov_j >= 0) set_at<A, CPT>(x, ov_j, (A)ov_val);
            if (ov_slow) {  // several overridden cells in this thread: re-derive them (rare, slow)
                const int64_t n = a.n_steps - left;
                if (cell0 == 0 && a.lbc != FBR_BC_NONE) x[0] = (A)0;
                const int64_t jl = nx - 1 - cell0;
                if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) set_at<A, CPT>(x, (int)jl, (A)0);
#pragma unroll 1
                for (int k = 0; k < a.n_src; ++k) {
                    const int64_t jj = a.src_idx[k] - cell0;
                    if (jj >= 0 && jj < CPT) set_at<A, CPT>(x, (int)jj, (A)__ldg(a.src_table + n * a.n_src + k));
                }
            }
            // ---- forward sweep: y_i = b_i - l_i y_{i-1} ----
            A y = (A)0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) y = fma_t<A>(-l[j], y, x[j]);
#pragma unroll
            for (int s = 0; s < 5; ++s) y = fma_t<A>(PfL[s], __shfl_up_sync(kFull, y, 1 << s), y);
            A carry = __shfl_up_sync(kFull, y, 1);
            if (lane == 0) carry = (A)0;
            if (W > 1) {
                if (lane == 31) s_E[0][warp] = y;
                __syncthreads();
                A c = (A)0;
#pragma unroll
                for (int v = 0; v < W - 1; ++v) c = fma_t<A>(s_K[0][warp][v], s_E[0][v], c);
                carry = fma_t<A>(PfEx, c, carry);
            }
            y = carry;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                y = fma_t<A>(-l[j], y, x[j]);
                x[j] = r[j] * y;  // t_j = r_j y_j
            }
            // ---- backward sweep: x_i = t_i - g_i x_{i+1} ----
            A z = (A)0;
#pragma unroll
            for (int j = CPT - 1; j >= 0; --j) z = fma_t<A>(-g[j], z, x[j]);
#pragma unroll
            for (int s = 0; s < 5; ++s) z = fma_t<A>(PbL[s], __shfl_down_sync(kFull, z, 1 << s), z);
            carry = __shfl_down_sync(kFull, z, 1);
            if (lane == 31) carry = (A)0;
            if (W > 1) {
                if (lane == 0) s_E[1][warp] = z;
                __syncthreads();
                A c = (A)0;
#pragma unroll
                for (int v = 1; v < W; ++v) c = fma_t<A>(s_K[1][warp][v], s_E[1][v], c);
                carry = fma_t<A>(PbEx, c, carry);
            }
            z = carry;
#pragma unroll
            for (int j = CPT - 1; j >= 0; --j) {
                z = fma_t<A>(-g[j], z, x[j]);
                x[j] = WIDE ? (A)(T)z : z;  // (the recurrence itself continues unrounded: z is not the stored state)
            }
            // ---- fused objective ----
            if (ob_j >= 0) {  // one observed cell: its weight is applied once, after the loop
                const double res = (double)get_at<A, CPT>(x, ob_j) - ob_val;
                acc = fma(res, res, acc);
            }
            if (ob_slow) {
                const int64_t n = a.n_steps - left;
#pragma unroll 1
                for (int k = 0; k < a.n_obs; ++k) {
                    const int64_t jj = a.obs_idx[k] - cell0;
                    if (jj >= 0 && jj < CPT) {
                        const double res = (double)get_at<A, CPT>(x, (int)jj) - __ldg(a.obs + (n + 1) * a.n_obs + k);
                        acc = fma((a.obs_w ? a.obs_w[k] : 1.0) * res, res, acc);
                    }
                }
            }
            // ---- sampled snapshot ----
            if (--snap_countdown == 0) {
                snap_countdown = snap_every;
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    if (cell0 + j < nx) snap_dst[j * a.snap_cell_stride] = (T)x[j];
                snap_dst += snap_stride;
            }
            ov_val = ov_next;
            ob_val = ob_next;
        }

        // ---------------- per-model epilogue ----------------
        if (a.p_final) {
            T *dst = reinterpret_cast<T *>(a.p_final) + m * nx + cell0;
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (cell0 + j < nx) dst[j] = x[j];
        }
        if (a.misfit) {
            acc *= ob_w;  // 1 unless this thread took the single-observation fast path with weights
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
            if (lane == 0) s_red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < W; ++w) tot += s_red[w];
                a.misfit[m] = tot;
            }
        }
        __syncthreads();  // shared scratch is reused by the next model
    }
}

#ifdef FBR_K3_PROF  // debug build: where a member's time goes in run_onchipx_kernel (scratch/k3_prof.py reads it)
__device__ long long g_k3_prof[8 * 8 * 4];  // [CTA slot][member of that CTA][prologue, block boundaries, step runs, epilogue]
#define K3_CLK() clock64()
#else
#define K3_CLK() 0ll
#endif

// ---- fp64, 8 cells per thread: the time loop built from the half-chain sweep blocks ---------------
// Same algorithm and results as run_onchip_kernel<double, 8, W> (the cross-warp carry stays EXACT:
// ensemble members are not required to have a short decay horizon), on a register diet so that the
// loop holds almost nothing but the sweeps:
//   * which of a thread's cells is overridden / observed, and by which column of src_table / obs, is
//     one packed 32-bit descriptor; addresses are rebuilt from the step number inside the (divergent)
//     branch only the owning threads enter;
//   * snapshots are taken between runs of an inner loop that knows nothing about them;
//   * threads that own several overridden or observed cells (adjacent sources ...) are served by a
//     generic path guarded by one CTA-uniform flag.
// descriptor: [4:0] overridden cell (31 none)  [12:5] src column (0xFE: constant zero)
//             [17:13] observed cell (31 none)  [25:18] obs column
//             [26] the thread's FIRST cell is the left boundary row (right-hand side 0, folded into the factors)
constexpr unsigned kDescNone = 0x1F;  // "no cell" in a descriptor's cell field
constexpr unsigned kDescObShift = 13, kDescSrcColShift = 5, kDescObColShift = 18;
constexpr unsigned kDescZeroFirst = 1u << 26;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_k3() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_k3() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// v = x[j] for a run-time j (any CPT): a chain of selects, no dynamic register indexing
template <typename T, int CPT>
__device__ __forceinline__ T pick_at(const T (&x)[CPT], int j) {
    if constexpr (CPT <= 8) return get_at<T, CPT>(x, j);
    else {  // 16 cells (fp32 mode): a select tree on the four bits of j — 15 selects on 4 predicates instead of a chain of
            // 15 compare-and-select pairs
        static_assert(CPT == 16, "select tree for 16 cells");
        const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8;
        T a[8], c[4];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = b0 ? x[2 * k + 1] : x[2 * k];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = b1 ? a[2 * k + 1] : a[2 * k];
        const T d0 = b2 ? c[1] : c[0], d1 = b2 ? c[3] : c[2];
        return b3 ? d1 : d0;
    }
}

// A thread's 64 bytes straight to global memory (16-byte stores when aligned).
template <typename T, int CPT>
__device__ __forceinline__ void store_row_direct(T *__restrict__ dst, int64_t cell0, int64_t nx, const T (&x)[CPT], int64_t cs = 1) {
    if (cs != 1) {  // cell-major result layout (Data2D's (cell, time)): one element per cell, cs elements apart
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (cell0 + j < nx) dst[j * cs] = x[j];
        return;
    }
    if (cell0 + CPT <= nx && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        constexpr int PER = 16 / sizeof(T);
#pragma unroll
        for (int q = 0; q < CPT / PER; ++q) {
            if constexpr (sizeof(T) == 8) reinterpret_cast<double2 *>(dst)[q] = make_double2((double)x[2 * q], (double)x[2 * q + 1]);
            else reinterpret_cast<float4 *>(dst)[q] = make_float4((float)x[(4 * q) % CPT], (float)x[(4 * q + 1) % CPT],
                                                                   (float)x[(4 * q + 2) % CPT], (float)x[(4 * q + 3) % CPT]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (cell0 + j < nx) dst[j] = x[j];
    }
}

// fp32 mode, fbr_run_opts::snap_f64: a thread's 16 floats go out as doubles (the audited members of a sweep: their launch
// is a handful of CTAs, and a widening pass AFTER it would have to squeeze past the sweep's persistent grid — measured:
// 4.6 ms for 82 MB on the four free CTA slots, which made the end-to-end figure of the fp32 mode wait for the audit chain).
template <int CPT>
__device__ __forceinline__ void store_row_wide(double *__restrict__ dst, int64_t cell0, int64_t nx, const float (&x)[CPT], int64_t cs) {
    if (cs == 1 && cell0 + CPT <= nx && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < CPT / 2; ++q) reinterpret_cast<double2 *>(dst)[q] = make_double2((double)x[2 * q], (double)x[2 * q + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (cell0 + j < nx) dst[j * cs] = (double)x[j];
    }
}

// One row of a warp's 32 * CPT cells to global memory with coalesced stores.  A thread holds CPT
// consecutive cells (64 bytes); written directly, every store instruction would touch 32 half-filled
// sectors 64 bytes apart (measured: 2.3 us per recorded row).  Instead the warp transposes through a
// private shared-memory patch (rows padded by one element: conflict-free) and each instruction
// writes 32 consecutive elements.
template <typename T, int CPT>
__device__ __forceinline__ void store_row_coalesced(T *__restrict__ row, int64_t warp_cell0, int64_t nx, const T (&x)[CPT],
                                                    T *patch, int lane, int64_t cs = 1) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPT; ++j) patch[lane * (CPT + 1) + j] = x[j];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = 32 * i + lane;  // cell within the warp's 32 * CPT
        const T v = patch[(c / CPT) * (CPT + 1) + (c % CPT)];
        if (warp_cell0 + c < nx) row[(warp_cell0 + c) * cs] = v;
    }
}

// How a member's sweeps are scanned (decided per member from its own factors, CTA-uniform):
//   kScanExact : 5 Kogge-Stone stages per warp and the W x W table of warp-total products across warps — exact;
//   kScanNear  : 5 stages, carry from the neighbouring warp only — every interior warp's total product (32 * CPT
//                consecutive multipliers) is below the threshold, so longer carry chains vanish;
//   kScanNear4 : as kScanNear with 4 stages — every product of 16 consecutive chunks is below the threshold too.
// The threshold (RunArgs::scan_eps, default 2^-80) sits 8 orders of magnitude below fp64 rounding; the skipped terms
// are products of that size with partial sums of the sweep.  B200: 18.3 / 16.4 / 15.4 ms per C3 sweep for the bare
// loops (scratch/ubench_k3v.cu) — the table costs 6 shared loads and 3 dependent FMAs behind every barrier.
constexpr int kScanExact = 0, kScanNear = 1, kScanNear4 = 2;

// LEAN: the instantiation for launches that record nothing (no snapshots) — the step-at-a-time use, where the per-member
// prologue is all that counts: without the copies of the step loop that carry snapshot code the register allocator leaves
// far fewer spills on that path (W = 2: 192 -> see the ptxas log).
template <typename T, int CPT, int W, int MINB, bool ZONAL, bool LEAN = false>
__global__ void __launch_bounds__(32 * W, MINB) run_onchipx_kernel(const RunArgs a) {
    __shared__ Mat2 s_zw[ZONAL ? W : 1];
    __shared__ double s_zc[ZONAL ? W : 1];
    static_assert(CPT * sizeof(T) == 64, "a thread owns 64 bytes of state: 8 doubles or 16 floats");
    __shared__ __align__(16) T s_K[2][W][W];
    // zero-carry warp totals of the current sweep: forward total of warp v at [0][1 + v] ([0][0] stays zero),
    // backward total of warp v at [1][1 + v] ([1][W + 1] stays zero) — a warp's neighbour is one fixed slot away
    __shared__ __align__(16) T s_E[2][W + 2];
    __shared__ T s_dummy;
    __shared__ double s_zero, s_trash;  // constant zero source value; where non-observers "record"
    __shared__ double s_T[2][W];
    __shared__ double s_red[W];
    __shared__ long long s_snap_slot;
    __shared__ double s_prev[kMaxObs], s_base[kMaxObs];  // observed cells: last row of the previous block; baseline sample
    extern __shared__ __align__(16) double s_dyn[];  // staged per-step scalars, see the time loop
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nx = a.nx;
    const int64_t cell0 = (int64_t)tid * CPT;
    if (tid == 0) { s_zero = 0.0; s_E[0][0] = (T)0; s_E[1][W + 1] = (T)0; }
    const unsigned st_f = smem_addr(lane == 31 ? &s_E[0][1 + warp] : &s_dummy);
    const unsigned st_b = smem_addr(lane == 0 ? &s_E[1][1 + warp] : &s_dummy);
    const int keep_f = lane == 0 ? 0 : -1, keep_b = lane == 31 ? 0 : -1;
    const int n_steps = (int)a.n_steps;

    // Which of this thread's cells are overridden / observed depends only on the mesh length, the
    // source / observation indices and the boundary codes — not on the member: work it out once.
    unsigned desc = kDescNone | (kDescNone << kDescObShift);
    bool ov_slow = false, ob_slow = false;
    if (cell0 < nx) {
        // slot of every owned cell (none / constant zero / src column), a bit mask of the cells that have one
        unsigned have = 0, slot_of_last = kSlotNone;
        int last = -1, n_ov = 0, n_ob = 0;
        auto claim = [&](int j, unsigned so) {  // later claims win (matbuilder.py:156-161: sources over boundary rows)
            if (!((have >> j) & 1)) { have |= 1u << j; ++n_ov; }
            if (n_ov == 1 || j == last) { last = j; slot_of_last = so; }
            else last = -2;  // several cells: the generic path re-derives them every step
        };
        auto is_source = [&](int64_t cell) {
            for (int k = 0; k < a.n_src; ++k)
                if (a.src_idx[k] == cell) return true;
            return false;
        };
        // The left boundary row (cell 0, right-hand side 0) is folded into the factors (see the per-model setup) unless a
        // source sits on it; every other overridden cell takes its value from shared memory.
        const int64_t jl = nx - 1 - cell0;
        if (cell0 == 0 && a.lbc != FBR_BC_NONE) {
            if (!is_source(0)) desc |= kDescZeroFirst;
            else claim(0, kSlotZero);
        }
        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) claim((int)jl, kSlotZero);
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {  // ascending k: later duplicates win
            const int64_t jj = a.src_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) claim((int)jj, (unsigned)k);
        }
        if (n_ov == 1) desc = (desc & ~0x1FFFu) | (unsigned)last | (slot_of_last << kDescSrcColShift);
#pragma unroll 1
        for (int k = 0; k < a.n_obs; ++k) {
            const int64_t jj = a.obs_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.obs_idx[k] < nx) {
                ++n_ob;
                desc = (desc & ~(0x1FFFu << kDescObShift)) | ((unsigned)jj << kDescObShift) | ((unsigned)k << kDescObColShift);
            }
        }
        if (n_ov > 1) { ov_slow = true; desc |= kDescNone; }
        if (n_ob > 1) { ob_slow = true; desc |= kDescNone << kDescObShift; }
    }
    // The common copies of the step loop serve overridden cells in a thread's FIRST or LAST register only (two selects, no
    // branch: boundary rows, sources on chunk boundaries); any other layout takes the generic copies.
    const bool mid_ov = (desc & 0x1F) != kDescNone && (desc & 0x1F) != 0 && (desc & 0x1F) != CPT - 1;
    const bool any_slow = __syncthreads_or(ov_slow || ob_slow || mid_ov);
    // 16-byte loads of a thread's 8 cells when the rows allow it (nx even keeps every member's row aligned)
    const bool vec_ok = (nx % 2 == 0) && cell0 + CPT <= nx && ((reinterpret_cast<uintptr_t>(a.fl) | reinterpret_cast<uintptr_t>(a.fr) |
                         reinterpret_cast<uintptr_t>(a.fg) | reinterpret_cast<uintptr_t>(a.p0)) & 15) == 0 && a.p0_stride % 2 == 0;
    auto load_row = [&](const double *row, double pad, T (&v)[CPT]) -> double {  // returns the fp64 product of the row's chunk
        double prod = 1.0;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(row + cell0) & 15);
        if (vec_ok || (cell0 + CPT <= nx && mis == 0)) {
            const double2 *q = reinterpret_cast<const double2 *>(row + cell0);
#pragma unroll
            for (int j = 0; j < CPT / 2; ++j) {
                const double2 t = __ldg(q + j);
                v[2 * j] = (T)t.x;
                v[2 * j + 1] = (T)t.y;
                prod *= t.x * t.y;
            }
        } else if (cell0 + CPT <= nx && mis == 8) {
            // a full chunk starting on an odd element (odd nx: every other member's row): 8-byte load, 16-byte loads, 8-byte load
            const double first = __ldg(row + cell0), last = __ldg(row + cell0 + CPT - 1);
            v[0] = (T)first;
            prod = first * last;
            const double2 *q = reinterpret_cast<const double2 *>(row + cell0 + 1);
#pragma unroll
            for (int j = 0; j < CPT / 2 - 1; ++j) {
                const double2 t = __ldg(q + j);
                v[2 * j + 1] = (T)t.x;
                v[2 * j + 2] = (T)t.y;
                prod *= t.x * t.y;
            }
            v[CPT - 1] = (T)last;
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const double t = cell0 + j < nx ? __ldg(row + cell0 + j) : pad;  // padding cell: identity row
                v[j] = (T)t;
                prod *= t;
            }
        }
        return prod;
    };

    unsigned src_rows = 0;  // which of my cells are source rows of the matrix (on-chip assembly mode)
    if (ZONAL) {
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {
            const int64_t jj = a.src_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) src_rows |= 1u << jj;
        }
    }
    const bool has_ov = (desc & 0x1F) != kDescNone, has_ob = ((desc >> kDescObShift) & 0x1F) != kDescNone;
    const bool ov_zero = ((desc >> kDescSrcColShift) & 0xFF) == kSlotZero;
    const bool ov_first = (desc & 0x1F) == 0, ov_last = (desc & 0x1F) == CPT - 1;
    // Per-step scalars never travel through a global load inside the sweeps: a load in flight across a sweep
    // shares a scoreboard with its shuffles / shared-memory reads and stalls them (+25 % per step on B200,
    // scratch/ubench_k3.cu), and a load consumed on the spot exposes L2 latency on an in-order warp.  Instead the
    // time axis is cut into blocks of `blk` steps:
    //   * the block's source values stream into shared memory (cp.async) while the previous block is integrated;
    //   * observed pressures are only WRITTEN during the block (a store has no result latency) into a trace, and
    //     the residuals of the observation samples a block completes are summed cooperatively at the next block
    //     boundary (samples on the step times, or interpolated onto the observations' own time axis).
    const int blk = a.blk_steps;
    double *s_src = s_dyn;                                  // [2][blk][n_src]
    double *s_obs = s_src + 2 * blk * a.n_src;              // [2][blk][n_obs] observation samples the block completes
    double *s_tr = s_obs + 2 * blk * a.n_obs;               // [2][blk][n_obs] trace of the observed cells
    T *s_patch = reinterpret_cast<T *>(s_tr + 2 * blk * a.n_obs + warp * (32 * 9));  // [W] row-transposition patches (2304 B each)
    // Short runs (the step-at-a-time use: n_steps = 1 .. 16) live on the latency of loading a member's four rows: while a
    // member is integrated the NEXT one's rows arrive in shared memory by four bulk copies (cp.async.bulk + mbarrier) and are
    // read with the rotated start piece that keeps linear 64-byte rows conflict-free (csrc/fbr_run_pass.cu).  Round 2's first
    // two attempts at this (16-byte cp.async; bulk copies read with 4-way conflicts) were slower than no prefetch at all.
    [[maybe_unused]] double *s_rows = s_tr + 2 * blk * a.n_obs + W * (32 * 9);  // (behind the patches; fp64 instantiation only)
    [[maybe_unused]] const int64_t row_pitch = (int64_t)W * 32 * CPT;  // doubles per staged row (the team's width)
    __shared__ __align__(8) unsigned long long s_rbar;
    [[maybe_unused]] const unsigned rbar = smem_addr(&s_rbar);
    [[maybe_unused]] unsigned n_rows_taken = 0;
    constexpr bool kCanPrefetch = LEAN && !ZONAL && sizeof(T) == 8;
    const bool prefetch = kCanPrefetch && a.prefetch_rows;
    if (kCanPrefetch && prefetch) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    [[maybe_unused]] auto stage_member = [&](int64_t mm) {  // lane 0 of warps 0 .. min(W, 4) - 1 (W = 1: lane 0 issues all four)
        if (lane != 0) return;
        const unsigned bytes = (unsigned)(nx * sizeof(double));
        const double *src[4] = {a.fl + mm * nx, a.fr + mm * nx, a.fg + mm * nx, a.p0 + mm * a.p0_stride};
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (warp == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rbar), "r"(4u * bytes) : "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (warp == (W >= 4 ? q : W >= 2 ? (q & 1) : 0)) bulk_g2s(smem_addr(s_rows + q * row_pitch), src[q], bytes, rbar);
    };
    [[maybe_unused]] auto take_row = [&](int q, double pad, T (&v)[CPT]) {  // my 64 bytes of staged row q
        if constexpr (sizeof(T) == 8) {
            const double *row = s_rows + q * row_pitch;
            if (cell0 + CPT <= nx) {
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
                    v[2 * pp] = (rot & 2) ? mid[(pp + 2) & 3].x : mid[pp].x;
                    v[2 * pp + 1] = (rot & 2) ? mid[(pp + 2) & 3].y : mid[pp].y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j) v[j] = cell0 + j < nx ? row[cell0 + j] : pad;
            }
        }
    };
    const int64_t cs = a.snap_cell_stride;

    ModelCursor cursor = model_cursor(a);
    int64_t m;
    long long preset;
    if (kCanPrefetch && prefetch && cursor.k < a.M) stage_member(cursor.k);  // (LEAN: nothing is recorded, members are k, k + G, ..)
    while (next_model(a, cursor, m, preset)) {
        // observation samples completed by rows (lo, hi] (for lo == 0 also those that only need the initial state)
        [[maybe_unused]] const long long k3t0 = K3_CLK();
        [[maybe_unused]] long long k3_bnd = 0, k3_run = 0;
        auto sample_range = [&](int lo, int hi, int &ja, int &jb) {
            if (a.sample_row) { ja = __ldg(a.row_first_sample + (lo == 0 ? 0 : lo + 1)); jb = __ldg(a.row_first_sample + hi + 1); }
            else { ja = lo == 0 ? 0 : lo + 1; jb = hi + 1; }
        };
        auto stage_block = [&](int b) {  // source values of steps [b blk, ..) and the observations those steps complete
            const int n0 = b * blk;
            const int rows = n_steps - n0 < blk ? n_steps - n0 : blk;
            if (rows <= 0) return;
            const int buf = b & 1;
            for (int e = tid; e < rows * a.n_src; e += 32 * W)
                cp_async8(s_src + buf * blk * a.n_src + e, a.src_table + (int64_t)n0 * a.n_src + e);
            if (a.n_obs > 0) {  // up to blk samples of the window (denser observations: the rest is read from global memory)
                int ja, jb;
                sample_range(n0, n0 + rows, ja, jb);
                ja = ja < a.first_sample ? a.first_sample : ja;
                jb = jb > a.last_sample ? a.last_sample : jb;
                jb = jb > ja + blk ? ja + blk : jb;
                for (int e = tid; e < (jb - ja) * a.n_obs; e += 32 * W)
                    cp_async8(s_obs + buf * blk * a.n_obs + e, a.obs + (int64_t)ja * a.n_obs + e);
            }
        };
        // the first block's scalars start their way into shared memory before anything else of this member is touched:
        // their latency hides behind the loads of the member's rows (it sat in front of the time loop before: one exposed
        // global -> shared round trip per member, which the step-at-a-time use pays for every single step)
        stage_block(0);
        cp_async_commit_k3();
        // ---------------- per-model setup ----------------
        T l[CPT], r[CPT], g[CPT], x[CPT];
        double cpf_d = 1.0, cpb_d = 1.0, tot_f_d = 0.0, tot_b_d = 0.0;  // fp64 chunk / warp products (used by the fp32 mode)
        if constexpr (ZONAL) {
            double zl[CPT], zr[CPT], zg[CPT];
            assemble_zonal_chunk<CPT>(a, m, (int)cell0, src_rows, zl, zr, zg);
            const int64_t left = nx - cell0;
            const int bad = factor_rows_in_registers<CPT, W>(zl, zr, zg, (int)cell0, (int)(left < 0 ? 0 : left > CPT ? CPT : left),
                                                             lane, warp, s_zw, s_zc);
            if (bad && a.z_singular) atomicMin(a.z_singular + m, bad);
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                l[j] = (T)zl[j]; r[j] = (T)zr[j]; g[j] = (T)zg[j];
                if (sizeof(T) == 4) { cpf_d *= zl[j]; cpb_d *= zg[j]; }
            }
        } else if (kCanPrefetch && prefetch) {
            mbar_wait(rbar, n_rows_taken & 1u);
            ++n_rows_taken;
            take_row(0, 0.0, l); take_row(1, 1.0, r); take_row(2, 0.0, g); take_row(3, 0.0, x);
            __syncthreads();  // everybody has its rows: the buffer is free for the next member's
            if (cursor.k < a.M) stage_member(cursor.k);
        } else {
            cpf_d = load_row(a.fl + m * nx, 0.0, l);
            load_row(a.fr + m * nx, 1.0, r);
            cpb_d = load_row(a.fg + m * nx, 0.0, g);
        }
        if (!(kCanPrefetch && prefetch)) load_row(a.p0 + m * a.p0_stride, 0.0, x);
        // Left boundary row (right-hand side always 0, matbuilder.py:49-55) in a thread's first register: with r_0 := 0 and
        // l_1 := 0 the stale x_0 never reaches anything (t_0 = r_0 y_0 = 0, y_1 = b_1 - l_1 * 0 = b_1 — the same bits) and
        // x_0 = -g_0 x_1 comes out of the backward sweep as before: the override costs nothing per step.
        if (desc & kDescZeroFirst) { r[0] = (T)0; l[1] = (T)0; }

        // scan multipliers: products of the chunk multipliers over runs of lanes / warps.  In the fp32 mode they are formed
        // in fp64 from the fp64 factors and rounded ONCE — a product of up to 512 rounded multipliers would otherwise carry
        // their accumulated rounding into every carry of every step.
        ScanMultipliersT<T> sm;
        if constexpr (sizeof(T) == 8) sm = scan_multipliers<T>(chunk_product<T, CPT>(l), chunk_product<T, CPT>(g), lane);
        else {
            const ScanMultipliersT<double> smd = scan_multipliers<double>(cpf_d, cpb_d, lane);
#pragma unroll
            for (int q = 0; q < 5; ++q) { sm.f[q] = (T)smd.f[q]; sm.b[q] = (T)smd.b[q]; }
            sm.ex_f = (T)smd.ex_f; sm.ex_b = (T)smd.ex_b; sm.tot_f = (T)smd.tot_f; sm.tot_b = (T)smd.tot_b;
            tot_f_d = smd.tot_f; tot_b_d = smd.tot_b;
        }
        if (W > 1) {  // (the barrier that ends every member protects the shared scratch of the previous one)
            if (lane == 31) s_T[0][warp] = sizeof(T) == 8 ? (double)sm.tot_f : tot_f_d;
            if (lane == 0) s_T[1][warp] = sizeof(T) == 8 ? (double)sm.tot_b : tot_b_d;
            __syncthreads();
            if (tid < W) {  // row tid of both tables
                double k = 1.0;
                for (int v = W - 1; v >= 0; --v) {  // forward: warps v < tid feed warp tid
                    if (v >= tid) s_K[0][tid][v] = (T)0;
                    else { s_K[0][tid][v] = (T)k; k *= s_T[0][v]; }
                }
                k = 1.0;
                for (int v = 0; v < W; ++v) {  // backward: warps v > tid feed warp tid
                    if (v <= tid) s_K[1][tid][v] = (T)0;
                    else { s_K[1][tid][v] = (T)k; k *= s_T[1][v]; }
                }
            }
        }
        if (tid == 0) {
            long long slot = preset;
            if (preset == -2) {
                slot = -1;
                for (int64_t q = 0; q < a.n_snap_models; ++q)
                    if (a.snap_models[q] == m) slot = q;
            }
            s_snap_slot = slot;
        }
        // observed cells: the initial state is "row 0" of the trace; no baseline yet
        if (has_ob) s_prev[(desc >> kDescObColShift) & 0xFF] = (double)pick_at<T, CPT>(x, (int)((desc >> kDescObShift) & (CPT - 1)));
        if (ob_slow) {
#pragma unroll 1
            for (int k = 0; k < a.n_obs; ++k) {
                const int64_t jj = a.obs_idx[k] - cell0;
                if (jj >= 0 && jj < CPT) s_prev[k] = (double)pick_at<T, CPT>(x, (int)jj);
            }
        }
        for (int k = tid; k < a.n_obs; k += 32 * W) s_base[k] = 0.0;
        // scan mode of this member (CTA-uniform): which far terms vanish below the threshold?
        bool far_ok = false;
        if (a.scan_eps > 0.0) far_ok = fabs((double)sm.f[4]) <= a.scan_eps && fabs((double)sm.b[4]) <= a.scan_eps;
        far_ok = __syncthreads_and(far_ok);
        bool near_ok = a.scan_eps > 0.0 && !any_slow;
        for (int v = 1; v + 1 < W; ++v)
            near_ok = near_ok && fabs((double)s_T[0][v]) <= a.scan_eps && fabs((double)s_T[1][v]) <= a.scan_eps;
        const int scan_mode = near_ok ? (far_ok ? kScanNear4 : kScanNear) : kScanExact;
        if (a.scan_mode && tid == 0) a.scan_mode[m] = scan_mode;
        const long long snap_slot = s_snap_slot;
        const bool recorded = snap_slot >= 0;
        double acc = 0.0;

        // Residuals of the observation samples whose last needed state lies in rows (lo, hi] — the block just integrated
        // (trace buffer `buf`; row lo itself is s_prev) — plus, for lo == 0, those that only need the initial state.
        auto flush_rows = [&](int lo, int hi, int buf) {
            if (!a.sample_row) {
                // observations on the step times: sample j is row j, trace and staged observations share one layout
                const double *tr = s_tr + buf * blk * a.n_obs, *so = s_obs + buf * blk * a.n_obs;
                const int n = (hi - lo) * a.n_obs;
                for (int e = tid; e < n; e += 32 * W) {
                    const double res = tr[e] - so[e];
                    const double wgt = a.obs_w ? __ldg(a.obs_w + e % a.n_obs) : 1.0;
                    acc = fma(wgt * res, res, acc);
                }
                return;
            }
            int ja, jb;
            sample_range(lo, hi, ja, jb);
            const int n_obs = a.n_obs;
            const double *tr0 = s_tr + (buf * blk - lo - 1) * n_obs;
            auto sample = [&](int j, int k) -> double {  // simulated channel k at observation sample j
                const int row = __ldg(a.sample_row + j);
                const double w = __ldg(a.sample_w + j);
                if (row < 0) return __longlong_as_double(0x7ff8000000000000ll);  // outside the simulated time range
                const double p_lo = row == lo ? s_prev[k] : tr0[row * n_obs + k];
                if (w == 0.0) return p_lo;
                return fma(w, tr0[(row + 1) * n_obs + k] - p_lo, p_lo);
            };
            if (a.base_sample >= ja && a.base_sample < jb) {  // "initial time calibration" (baseline_model_generator.py:398)
                for (int k = tid; k < n_obs; k += 32 * W) s_base[k] = sample(a.base_sample, k);
                __syncthreads();
            }
            ja = ja < a.first_sample ? a.first_sample : ja;
            jb = jb > a.last_sample ? a.last_sample : jb;
            const int j_staged = ja + blk;  // samples [ja, j_staged) were staged into shared memory with the block
            const double *so = s_obs + (buf * blk - ja) * n_obs;
            // (sample, channel) pairs dealt round-robin; j and k advance without a division per pair
            const int dj = (32 * W) / n_obs, dk = (32 * W) % n_obs;
            int j = ja + tid / n_obs, k = tid % n_obs;
            for (; j < jb; j += dj) {
                const double o = j < j_staged ? so[j * n_obs + k] : __ldg(a.obs + (int64_t)j * n_obs + k);
                const double res = (sample(j, k) - s_base[k]) - o;
                const double wgt = a.obs_w ? __ldg(a.obs_w + k) : 1.0;
                acc = fma(wgt * res, res, acc);
                k += dk;
                if (k >= n_obs) { k -= n_obs; ++j; }
            }
        };
        unsigned src_addr = smem_addr(&s_zero), tr_addr = smem_addr(&s_trash);
        const unsigned src_inc = (has_ov && !ov_zero) ? 8u * a.n_src : 0u, tr_inc = has_ob ? 8u * a.n_obs : 0u;

        // ---------------- time loop: runs of steps between block boundaries / snapshots ----------------
        const int snap_every = a.snapshot_every > 0x7fffffff ? 0x7fffffff : (int)a.snapshot_every;
        int to_snap = snap_slot >= 0 ? snap_every : 0x7fffffff;  // never reaches 0 when not recorded
        T *snap_row = reinterpret_cast<T *>(a.snap) + (snap_slot >= 0 ? snap_slot : 0) * a.snap_model_stride;
        // a handful of CTAs (single models): the row goes out with plain 16-byte stores straight from the
        // registers — no latency on the step; a full grid of recording members is bandwidth-bound and
        // takes the transposing, fully coalesced path
        const bool direct_rows = gridDim.x <= 64;
        int done = 0;
        [[maybe_unused]] const long long k3t1 = K3_CLK();
        while (done < n_steps) {
            const int b = done / blk;
            [[maybe_unused]] const long long k3ta = K3_CLK();
            if (done % blk == 0) {  // entering block b (CTA-uniform)
                cp_async_wait_all_k3();
                __syncthreads();  // block b's inputs have landed; everybody is done with block b - 1
                if (b > 0 && a.n_obs > 0) {
                    flush_rows((b - 1) * blk, b * blk, (b - 1) & 1);
                    __syncthreads();  // s_prev is replaced; block b + 1 is staged
                    for (int k = tid; k < a.n_obs; k += 32 * W) s_prev[k] = s_tr[(((b - 1) & 1) * blk + blk - 1) * a.n_obs + k];
                }
                stage_block(b + 1);
                cp_async_commit_k3();
                const int buf = b & 1;
                if (has_ov && !ov_zero) src_addr = smem_addr(s_src + buf * blk * a.n_src + ((desc >> kDescSrcColShift) & 0xFF));
                if (has_ob) tr_addr = smem_addr(s_tr + buf * blk * a.n_obs + ((desc >> kDescObColShift) & 0xFF));
            }
            int run = blk - done % blk;
            if (run > n_steps - done) run = n_steps - done;
            [[maybe_unused]] const long long k3tb = K3_CLK();
            // several copies of the step loop: the common ones carry no code for threads that own several
            // overridden / observed cells, nor for snapshots, and only the scan stages their mode needs
            auto steps = [&](auto mode_tag, auto slow_tag, auto rec_tag, int n_begin, int n_end) {
                constexpr int MODE = decltype(mode_tag)::value;
                constexpr bool SLOW = decltype(slow_tag)::value, REC = decltype(rec_tag)::value;
                constexpr int NS = MODE == kScanNear4 ? 4 : 5;
                // the value overriding this thread's cell (a source sample, or the constant zero) is read from shared
                // memory one step AHEAD, next to the backward sweep's barrier, whose wait hides the read's latency
                T v_ov = (T)lds_f64(src_addr);
#pragma unroll 1
                for (int n = n_begin; n < n_end; ++n) {
                    // right-hand side: b = p^n with boundary / source overrides (matbuilder.py:45-76)
                    if (!SLOW) {  // every overridden cell of the team sits in a first / last register
                        x[0] = ov_first ? v_ov : x[0];
                        x[CPT - 1] = ov_last ? v_ov : x[CPT - 1];
                    } else if (has_ov) {  // only the threads that own an overridden cell
                        const int jo = (int)(desc & 0x1F);
                        if (jo == 0) x[0] = v_ov;
                        else if (jo == CPT - 1) x[CPT - 1] = v_ov;
                        else set_at<T, CPT>(x, jo, v_ov);
                    }
                    if (SLOW && ov_slow) {  // several overridden cells in this thread: re-derive them (rare)
                        if (cell0 == 0 && a.lbc != FBR_BC_NONE) x[0] = (T)0;
                        const int64_t jl = nx - 1 - cell0;
                        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) set_at<T, CPT>(x, (int)jl, (T)0);
#pragma unroll 1
                        for (int k = 0; k < a.n_src; ++k) {
                            const int64_t jj = a.src_idx[k] - cell0;
                            if (jj >= 0 && jj < CPT) set_at<T, CPT>(x, (int)jj, (T)__ldg(a.src_table + (int64_t)n * a.n_src + k));
                        }
                    }
                    // ---- forward sweep: y_i = b_i - l_i y_{i-1}, t_i = r_i y_i ----
                    {
                        T y = chunk_end_fwd_plain<T, CPT>(x, l);
#pragma unroll
                        for (int s = 0; s < NS; ++s) y = fma_any<T>(sm.f[s], __shfl_up_sync(kAllLanes, y, 1 << s), y);
                        T c = and_mask(__shfl_up_sync(kAllLanes, y, 1), keep_f);
                        if (W > 1) {
                            sts_any(st_f, y);
                            __syncthreads();
                            if (MODE == kScanExact) {
                                T cw = (T)0;
#pragma unroll
                                for (int v = 0; v < W - 1; ++v) cw = fma_any<T>(s_K[0][warp][v], s_E[0][1 + v], cw);
                                c = fma_any<T>(sm.ex_f, cw, c);
                            } else c = fma_any<T>(sm.ex_f, s_E[0][warp], c);
                        }
                        chunk_apply_fwd_plain<T, CPT>(x, l, r, c);
                    }
                    // ---- backward sweep: x_i = t_i - g_i x_{i+1} ----
                    {
                        T z = chunk_end_bwd_plain<T, CPT>(x, g);
#pragma unroll
                        for (int s = 0; s < NS; ++s) z = fma_any<T>(sm.b[s], __shfl_down_sync(kAllLanes, z, 1 << s), z);
                        T c = and_mask(__shfl_down_sync(kAllLanes, z, 1), keep_b);
                        src_addr += src_inc;
                        v_ov = (T)lds_f64(src_addr);  // next step's override value (one row past the block's last is never used)
                        if (W > 1) {
                            sts_any(st_b, z);
                            __syncthreads();
                            if (MODE == kScanExact) {
                                T cw = (T)0;
#pragma unroll
                                for (int v = 1; v < W; ++v) cw = fma_any<T>(s_K[1][warp][v], s_E[1][1 + v], cw);
                                c = fma_any<T>(sm.ex_b, cw, c);
                            } else c = fma_any<T>(sm.ex_b, s_E[1][warp + 2], c);
                        }
                        chunk_apply_bwd_plain<T, CPT>(x, g, c);
                    }
                    // ---- fused objective: record the observed pressure (non-observers hit the dummy slot) ----
                    sts_f64(tr_addr, (double)pick_at<T, CPT>(x, (int)((desc >> kDescObShift) & (CPT - 1))));
                    tr_addr += tr_inc;
                    if (SLOW && ob_slow) {  // several observed cells in this thread: record each of them
                        double *row = s_tr + (((n / blk) & 1) * blk + n % blk) * a.n_obs;
#pragma unroll 1
                        for (int k = 0; k < a.n_obs; ++k) {
                            const int64_t jj = a.obs_idx[k] - cell0;
                            if (jj >= 0 && jj < CPT) row[k] = (double)pick_at<T, CPT>(x, (int)jj);
                        }
                    }
                    // ---- sampled snapshot (CTA-uniform countdown; never fires for unrecorded members) ----
                    if (REC && --to_snap == 0) {
                        to_snap = snap_every;
                        bool stored = false;
                        if constexpr (sizeof(T) == 4) {
                            if (a.snap_wide) {  // same element offsets, 8-byte elements
                                double *wide = reinterpret_cast<double *>(a.snap) + (snap_row - reinterpret_cast<T *>(a.snap));
                                store_row_wide<CPT>(wide + cell0 * cs, cell0, nx, x, cs);
                                stored = true;
                            }
                        }
                        if (!stored) {
                            if (direct_rows) store_row_direct<T, CPT>(snap_row + cell0 * cs, cell0, nx, x, cs);
                            else store_row_coalesced<T, CPT>(snap_row, (int64_t)warp * 32 * CPT, nx, x, s_patch, lane, cs);
                        }
                        snap_row += a.snap_row_stride;
                    }
                }
            };
            using I0 = std::integral_constant<int, kScanExact>;
            using I1 = std::integral_constant<int, kScanNear>;
            using I2 = std::integral_constant<int, kScanNear4>;
            if constexpr (LEAN) {  // nothing is recorded in this instantiation (the launcher checked)
                if (scan_mode == kScanExact) {
                    if (any_slow) steps(I0{}, std::true_type{}, std::false_type{}, done, done + run);
                    else steps(I0{}, std::false_type{}, std::false_type{}, done, done + run);
                } else if (scan_mode == kScanNear) steps(I1{}, std::false_type{}, std::false_type{}, done, done + run);
                else steps(I2{}, std::false_type{}, std::false_type{}, done, done + run);
            } else if (scan_mode == kScanExact) {
                if (recorded) {  // (the copies for unrecorded members carry no snapshot code)
                    if (any_slow) steps(I0{}, std::true_type{}, std::true_type{}, done, done + run);
                    else steps(I0{}, std::false_type{}, std::true_type{}, done, done + run);
                } else {
                    if (any_slow) steps(I0{}, std::true_type{}, std::false_type{}, done, done + run);
                    else steps(I0{}, std::false_type{}, std::false_type{}, done, done + run);
                }
            } else if (scan_mode == kScanNear) {
                if (recorded) steps(I1{}, std::false_type{}, std::true_type{}, done, done + run);
                else steps(I1{}, std::false_type{}, std::false_type{}, done, done + run);
            } else {
                if (recorded) steps(I2{}, std::false_type{}, std::true_type{}, done, done + run);
                else steps(I2{}, std::false_type{}, std::false_type{}, done, done + run);
            }
            done += run;
#ifdef FBR_K3_PROF
            { const long long k3tc = K3_CLK(); k3_bnd += k3tb - k3ta; k3_run += k3tc - k3tb; }
#endif
        }
        [[maybe_unused]] const long long k3t2 = K3_CLK();
        cp_async_wait_all_k3();
        __syncthreads();
        if (n_steps > 0 && a.n_obs > 0) {
            const int b_last = (n_steps - 1) / blk;
            flush_rows(b_last * blk, n_steps, b_last & 1);
        }

        // ---------------- per-model epilogue ----------------
        if (a.p_final)
            store_row_coalesced<T, CPT>(reinterpret_cast<T *>(a.p_final) + m * nx, (int64_t)warp * 32 * CPT, nx, x, s_patch, lane);
        if (a.misfit) {
            for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
            if (lane == 0) s_red[warp] = acc;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < W; ++w) tot += s_red[w];
                a.misfit[m] = tot;
            }
        }
        __syncthreads();  // shared scratch is reused by the next model
#ifdef FBR_K3_PROF
        if (tid == 0 && blockIdx.x % 71 == 0 && blockIdx.x / 71 < 8) {
            static_assert(true, "");
            long long *q = g_k3_prof + (blockIdx.x / 71) * 32;
            int slot = 0;
            while (slot < 8 && q[slot * 4 + 2] != 0) ++slot;
            if (slot < 8) { q[slot * 4] = k3t1 - k3t0; q[slot * 4 + 1] = k3_bnd; q[slot * 4 + 2] = k3_run; q[slot * 4 + 3] = K3_CLK() - k3t2; }
        }
#endif
    }
}

// ---- short meshes: several members per warp ---------------------------------------------------------------------
// Meshes of up to 64 / 128 cells (the reference's own models have 200) would leave most lanes of a warp idle with 8
// cells per thread, and most of a thread's registers idle with fewer.  Here a member is owned by a GROUP of L = 2 .. 16
// lanes (8 cells per thread), a warp integrates 32 / L members in lock-step (all members share n_steps), shuffles stay
// inside a group (log2 L scan stages), and nothing in the time loop touches shared memory or a barrier.  Per-step
// scalars (source values, observations) are prefetched one step ahead by the threads that own the cell.
template <typename T, int L>
__global__ void __launch_bounds__(128, 4) run_group_kernel(const RunArgs a) {
    constexpr int CPT = 8, GPW = 32 / L, NS = L == 16 ? 4 : L == 8 ? 3 : L == 4 ? 2 : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane & (L - 1), grp = lane / L;
    const int nx = (int)a.nx, cell0 = sub * CPT;
    const int64_t n_groups = (int64_t)gridDim.x * 4 * GPW;
    // which of my cells are overridden / observed (member-independent)
    int ov_j = -1, ob_j = -1, ov_col = -1, ob_col = -1;
    bool ov_slow = false, ob_slow = false;
    if (cell0 < nx) {
        uint64_t ov = ~0ull;
        if (cell0 == 0 && a.lbc != FBR_BC_NONE) ov = (ov & ~0xFFull) | kSlotZero;
        const int jl = nx - 1 - cell0;
        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) ov = (ov & ~(0xFFull << (8 * jl))) | ((uint64_t)kSlotZero << (8 * jl));
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {  // ascending k: later duplicates win (matbuilder.py:156-161)
            const int64_t jj = a.src_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) ov = (ov & ~(0xFFull << (8 * jj))) | ((uint64_t)k << (8 * jj));
        }
        int n_ov = 0, n_ob = 0;
#pragma unroll 1
        for (int j = 0; j < CPT; ++j) {
            const unsigned so = (unsigned)(ov >> (8 * j)) & 0xFF;
            if (so != kSlotNone) { ++n_ov; ov_j = j; ov_col = so == kSlotZero ? -1 : (int)so; }
        }
#pragma unroll 1
        for (int k = 0; k < a.n_obs; ++k) {
            const int64_t jj = a.obs_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.obs_idx[k] < nx) { ++n_ob; ob_j = (int)jj; ob_col = k; }
        }
        if (n_ov > 1) { ov_slow = true; ov_j = -1; ov_col = -1; }
        if (n_ob > 1) { ob_slow = true; ob_j = -1; ob_col = -1; }
    }
    const double ob_w = (ob_j >= 0 && a.obs_w) ? a.obs_w[ob_col] : 1.0;
    const int keep_f = sub == 0 ? 0 : -1, keep_b = sub == L - 1 ? 0 : -1;
    const int snap_every = a.snapshot_every > 0x7fffffff ? 0x7fffffff : (int)a.snapshot_every;

    // the loop bound is uniform over the warp; groups past the last member redo member M - 1 and write nothing
    for (int64_t m_raw = ((int64_t)blockIdx.x * 4 + warp) * GPW + grp; m_raw - grp < a.M; m_raw += n_groups) {
        const bool live = m_raw < a.M;
        const int64_t m = live ? m_raw : a.M - 1;
        T l[CPT], r[CPT], g[CPT], x[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int i = cell0 + j;
            if (i < nx) {
                l[j] = (T)__ldg(a.fl + m * nx + i);
                r[j] = (T)__ldg(a.fr + m * nx + i);
                g[j] = (T)__ldg(a.fg + m * nx + i);
                x[j] = (T)__ldg(a.p0 + m * a.p0_stride + i);
            } else {  // padding cell: identity row, decoupled
                l[j] = (T)0; r[j] = (T)1; g[j] = (T)0; x[j] = (T)0;
            }
        }
        // stage multipliers of the group scans (zero where a stage does not reach: no predicates in the loop)
        T Pf[NS], Pb[NS], PfEx, PbEx;
        {
            T pf = chunk_product<T, CPT>(l), pb = chunk_product<T, CPT>(g);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const bool fa = sub >= (1 << s), ba = sub + (1 << s) < L;
                Pf[s] = fa ? pf : (T)0;
                Pb[s] = ba ? pb : (T)0;
                const T upf = __shfl_up_sync(kFull, pf, 1 << s, L), dnb = __shfl_down_sync(kFull, pb, 1 << s, L);
                if (fa) pf *= upf;
                if (ba) pb *= dnb;
            }
        }
        // snapshot slot of this member (-1: not recorded)
        long long slot = -1;
        if (live && a.snap) {
            if (!a.snap_models) slot = m;
            else
                for (int64_t q = 0; q < a.n_snap_models; ++q)
                    if (a.snap_models[q] == m) slot = q;
        }
        int snap_countdown = slot >= 0 ? snap_every : 0x7fffffff;
        T *snap_dst = reinterpret_cast<T *>(a.snap) + (slot >= 0 ? slot : 0) * a.snap_model_stride + cell0 * a.snap_cell_stride;
        const double *ovp = ov_col >= 0 ? a.src_table + ov_col : &kZero, *obp = ob_col >= 0 ? a.obs + a.n_obs + ob_col : &kZero;
        int ov_stride = ov_col >= 0 ? a.n_src : 0, ob_stride = ob_col >= 0 ? a.n_obs : 0;
        double ov_val = a.n_steps > 0 ? __ldg(ovp) : 0.0, ob_val = a.n_steps > 0 ? __ldg(obp) : 0.0, acc = 0.0;

#pragma unroll 1
        for (int left = (int)a.n_steps; left > 0; --left) {
            if (left == 1) { ov_stride = 0; ob_stride = 0; }  // no row after the last one
            ovp += ov_stride;
            obp += ob_stride;
            const double ov_next = __ldg(ovp), ob_next = __ldg(obp);
            if (ov_j >= 0) set_at<T, CPT>(x, ov_j, (T)ov_val);
            if (ov_slow) {  // several overridden cells in this thread: re-derive them (rare)
                const int64_t n = a.n_steps - left;
                if (cell0 == 0 && a.lbc != FBR_BC_NONE) x[0] = (T)0;
                const int jl = nx - 1 - cell0;
                if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) set_at<T, CPT>(x, jl, (T)0);
#pragma unroll 1
                for (int k = 0; k < a.n_src; ++k) {
                    const int64_t jj = a.src_idx[k] - cell0;
                    if (jj >= 0 && jj < CPT) set_at<T, CPT>(x, (int)jj, (T)__ldg(a.src_table + n * a.n_src + k));
                }
            }
            {   // forward sweep
                T y = chunk_end_fwd_plain<T, CPT>(x, l);
#pragma unroll
                for (int s = 0; s < NS; ++s) y = fma_any<T>(Pf[s], __shfl_up_sync(kFull, y, 1 << s, L), y);
                const T c = and_mask(__shfl_up_sync(kFull, y, 1, L), keep_f);
                chunk_apply_fwd_plain<T, CPT>(x, l, r, c);
            }
            {   // backward sweep
                T z = chunk_end_bwd_plain<T, CPT>(x, g);
#pragma unroll
                for (int s = 0; s < NS; ++s) z = fma_any<T>(Pb[s], __shfl_down_sync(kFull, z, 1 << s, L), z);
                const T c = and_mask(__shfl_down_sync(kFull, z, 1, L), keep_b);
                chunk_apply_bwd_plain<T, CPT>(x, g, c);
            }
            if (ob_j >= 0) {  // one observed cell: its weight is applied once, after the loop
                const double res = (double)get_at<T, CPT>(x, ob_j) - ob_val;
                acc = fma(res, res, acc);
            }
            if (ob_slow) {
                const int64_t n = a.n_steps - left;
#pragma unroll 1
                for (int k = 0; k < a.n_obs; ++k) {
                    const int64_t jj = a.obs_idx[k] - cell0;
                    if (jj >= 0 && jj < CPT) {
                        const double res = (double)get_at<T, CPT>(x, (int)jj) - __ldg(a.obs + (n + 1) * a.n_obs + k);
                        acc = fma((a.obs_w ? a.obs_w[k] : 1.0) * res, res, acc);
                    }
                }
            }
            if (--snap_countdown == 0) {
                snap_countdown = snap_every;
                store_row_direct<T, CPT>(snap_dst, cell0, nx, x, a.snap_cell_stride);
                snap_dst += a.snap_row_stride;
            }
            ov_val = ov_next;
            ob_val = ob_next;
        }
        if (a.p_final && live) store_row_direct<T, CPT>(reinterpret_cast<T *>(a.p_final) + m * nx + cell0, cell0, nx, x);
        if (a.misfit) {
            acc *= ob_w;
#pragma unroll
            for (int s = 1; s < L; s <<= 1) acc += __shfl_xor_sync(kFull, acc, s, L);
            if (sub == 0 && live) a.misfit[m] = acc;
        }
    }
}

template <typename T, int L>
static int launch_group(const RunArgs &a, cudaStream_t stream) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    auto kern = run_group_kernel<T, L>;
    int occ = 0;
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "group run kernel does not fit on an SM");
    const int64_t per_cta = 4 * (32 / L);
    int64_t grid = (int64_t)occ * f.sm_count;
    if (grid > (a.M + per_cta - 1) / per_cta) grid = (a.M + per_cta - 1) / per_cta;
    kern<<<(unsigned)grid, 128, 0, stream>>>(a);
    return check_launch("fbr_run (member groups)");
}

template <typename T, int CPT, int W, bool ZONAL = false>
static int launch_onchipx(const RunArgs &a, cudaStream_t stream) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    // 16 warps per SM (128 registers per thread).  A fifth 4-warp CTA per SM at 96 registers was measured again in round 2
    // with the leaner loops: 11 spill loads per step, 25.7 against 21.3 ms per C3 sweep.
    constexpr int kFull16 = (16 / W) > 0 ? (16 / W) : 1;
    void (*kern)(const RunArgs) = run_onchipx_kernel<T, CPT, W, kFull16, ZONAL>;
    if constexpr (!ZONAL && sizeof(T) == 8) {
        if (!a.snap && !getenv("FBR_K3_NO_LEAN")) kern = run_onchipx_kernel<T, CPT, W, kFull16, ZONAL, true>;
    }
    RunArgs b = a;
    const int cols = a.n_src + 2 * a.n_obs;
    // short runs without records (the step-at-a-time use): prefetch the next member's rows through shared memory
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool prefetch = false;
    if constexpr (!ZONAL && sizeof(T) == 8)
        prefetch = !a.snap && !getenv("FBR_K3_NO_LEAN") && !getenv("FBR_K3_NO_PREFETCH") && a.n_steps <= 16 && a.nx % 2 == 0 &&
                   a.p0_stride % 2 == 0 && al16(a.fl) && al16(a.fr) && al16(a.fg) && al16(a.p0) && a.M > 1;
    b.prefetch_rows = prefetch;
    const size_t rows_bytes = prefetch ? 4 * (size_t)W * 32 * CPT * sizeof(double) : 0;
    // Time steps per staged block of per-step scalars: the longest block that costs no resident CTA (block boundaries
    // carry three barriers and the residual sums; C3: 64 steps 25.2 ms, 32 steps 25.5 ms, 128 steps lose a CTA: 42.7 ms).
    // The audited side launch of an ensemble must end up with the SAME footprint as the sweep (see ensemble.py): the
    // choice depends only on the column count and the team width, not on M.
    auto smem_for = [&](int blk) { return (2 * (size_t)blk * (size_t)cols + (size_t)W * 32 * 9) * sizeof(double) + rows_bytes; };
    int occ = 0, occ_ref = 0;
    size_t smem = smem_for(8);
    if (smem > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ref, kern, 32 * W, smem));
    b.blk_steps = 8;
    occ = occ_ref;
    for (int blk : {64, 32, 16}) {
        const size_t need = smem_for(blk);
        if (need > (size_t)f.smem_optin) continue;
        if (need > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        int o = 0;
        FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 32 * W, need));
        if (o == occ_ref) { b.blk_steps = blk; smem = need; occ = o; break; }
    }
    if (const char *e = getenv("FBR_K3_BLK")) {  // tuning aid
        b.blk_steps = atoi(e) > 0 ? atoi(e) : 8;
        smem = smem_for(b.blk_steps);
        if (smem > 48 * 1024) FBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, smem));
    }
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "run kernel does not fit on an SM (threads=%d)", 32 * W);
    int64_t grid = (int64_t)occ * f.sm_count - a.reserve_ctas;  // fbr_run_opts::reserve_ctas
    if (grid < 1) grid = 1;
    if (grid > a.M) grid = a.M;
    // long sweeps that record nothing: members after the first deal are handed out on demand (next_model)
    b.ticket = nullptr;
    if (!a.snap && !prefetch && a.M > grid && a.n_steps >= 64 && !getenv("FBR_K3_STATIC_DEAL")) {
        static std::atomic<unsigned> next_slot{0};
        unsigned long long *base = nullptr;
        FBR_CUDA(cudaGetSymbolAddress((void **)&base, g_k3_ticket));
        b.ticket = base + next_slot.fetch_add(1) % kTicketSlots;  // concurrent launches (streams, threads) never share a word
        FBR_CUDA(cudaMemsetAsync(b.ticket, 0, sizeof(unsigned long long), stream));
    }
    kern<<<(unsigned)grid, 32 * W, smem, stream>>>(b);
    return check_launch("fbr_run (on-chip, 64 B per thread)");
}

// 64 bytes of state per thread: 8 cells in fp64, 16 cells in the fp32 mode
template <typename T>
static int launch_x(const RunArgs &a, cudaStream_t stream) {
    constexpr int CPT = 64 / sizeof(T);
    const int64_t threads = (a.nx + CPT - 1) / CPT;
    if (a.z_mesh) {  // on-chip assembly + factorisation (fbr_run_zonal_f64 / _f32: assembled and factorised in fp64)
        if (threads <= 32) return launch_onchipx<T, CPT, 1, true>(a, stream);
        if (threads <= 64) return launch_onchipx<T, CPT, 2, true>(a, stream);
        if (threads <= 128) return launch_onchipx<T, CPT, 4, true>(a, stream);
        if (threads <= 256) return launch_onchipx<T, CPT, 8, true>(a, stream);
        return launch_onchipx<T, CPT, 16, true>(a, stream);
    }
    if (threads <= 32) return launch_onchipx<T, CPT, 1>(a, stream);
    if (threads <= 64) return launch_onchipx<T, CPT, 2>(a, stream);
    if (threads <= 128) return launch_onchipx<T, CPT, 4>(a, stream);
    if (threads <= 256) return launch_onchipx<T, CPT, 8>(a, stream);
    return launch_onchipx<T, CPT, 16>(a, stream);
}

template <typename T, int CPT, int W, bool ZONAL, bool WIDE>
static int launch_onchip(const RunArgs &a, cudaStream_t stream) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    auto kern = run_onchip_kernel<T, CPT, W, ZONAL, WIDE>;
    int occ = 0;
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, 0));
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "run kernel does not fit on an SM (threads=%d)", 32 * W);
    int64_t grid = (int64_t)occ * f.sm_count;
    if (grid > a.M) grid = a.M;
    kern<<<(unsigned)grid, 32 * W, 0, stream>>>(a);
    return check_launch("fbr_run (on-chip)");
}

template <typename T, int CPT, bool ZONAL = false, bool WIDE = false>
static int launch_cpt(const RunArgs &a, cudaStream_t stream) {
    const int64_t threads = (a.nx + CPT - 1) / CPT;
    if (threads <= 32) return launch_onchip<T, CPT, 1, ZONAL, WIDE>(a, stream);
    if (threads <= 64) return launch_onchip<T, CPT, 2, ZONAL, WIDE>(a, stream);
    if (threads <= 128) return launch_onchip<T, CPT, 4, ZONAL, WIDE>(a, stream);
    if (ZONAL) return fail(FBR_ERR_UNSUPPORTED, "on-chip assembly of short models covers up to 128 threads");
    if (threads <= 256) return launch_onchip<T, CPT, 8, false, WIDE>(a, stream);
    return launch_onchip<T, CPT, 16, false, WIDE>(a, stream);
}

template <typename T>
static int run_dispatch(const RunArgs &a, cudaStream_t stream) {
    FBR_REQUIRE(a.M >= 1 && a.nx >= 1 && a.n_steps >= 0, "M, nx must be >= 1 and n_steps >= 0");
    FBR_REQUIRE(a.n_steps <= 0x7fffffff, "n_steps per launch is limited to 2^31 - 1");
    FBR_REQUIRE(((a.fl && a.fr && a.fg) || a.z_mesh) && a.p0, "NULL factor / initial-state pointer");
    FBR_REQUIRE(a.p0_stride == 0 || a.p0_stride >= a.nx, "bad p0_stride");
    FBR_REQUIRE(a.lbc >= 0 && a.lbc <= 2 && a.rbc >= 0 && a.rbc <= 2, "bad boundary code");
    FBR_REQUIRE(a.n_src >= 0 && a.n_src <= kMaxSrc, "n_src=%d outside [0, %d]", a.n_src, kMaxSrc);
    FBR_REQUIRE(a.n_src == 0 || (a.src_idx && a.src_table), "sources given without src_idx / src_table");
    FBR_REQUIRE(a.n_obs >= 0 && a.n_obs <= kMaxObs, "n_obs=%d outside [0, %d]", a.n_obs, kMaxObs);
    FBR_REQUIRE(a.n_obs == 0 || (a.obs_idx && a.obs && a.misfit), "observations given without obs / misfit");
    FBR_REQUIRE(!a.snap || (a.snapshot_every >= 1 && a.n_snap_models >= 1), "snapshot_every / n_snap_models must be >= 1");
    FBR_REQUIRE(!a.snap || a.snap_models || a.n_snap_models == a.M, "snap_models == NULL needs n_snap_models == M");
    FBR_REQUIRE(!a.snap || a.snap_cell_stride != 1 || (a.snap_row_stride >= a.nx && a.snap_model_stride >= a.nx),
                "snapshot strides must be >= nx");
    FBR_REQUIRE(!a.snap || a.snap_cell_stride >= 1, "snap_cell_stride must be >= 1");
    FBR_REQUIRE(a.n_samples == 0 || (a.n_obs > 0 && a.sample_row && a.sample_w && a.row_first_sample),
                "sampled observations need obs_idx, sample_row, sample_w and row_first_sample");
    FBR_REQUIRE(a.n_samples == 0 || (a.first_sample >= 0 && a.last_sample <= a.n_samples && a.base_sample < a.n_samples),
                "sample window / baseline index outside [0, n_samples)");
    FBR_REQUIRE(a.n_samples == 0 || a.base_sample < 0 || a.base_sample <= a.first_sample,
                "base_sample must not lie behind first_sample");
    if (a.nx > fbr_run_max_nx())
        return fail(FBR_ERR_UNSUPPORTED, "nx=%lld exceeds the on-chip run kernel limit %lld", (long long)a.nx,
                    (long long)fbr_run_max_nx());
    // ONE step with reused factors, nothing recorded but the new state (the step-at-a-time use): the persistent kernel's
    // per-member set-up (scan-mode detection, staging of per-step scalars, 12 loop copies) buys nothing here — K2's
    // register-resident kernel in its FACTORS mode does the two sweeps straight from the rows it prefetches
    // (65536 x 512: 0.25 -> 0.22 ms).
    if constexpr (sizeof(T) == 8) {
        if (a.n_steps == 1 && !a.snap && a.n_obs == 0 && a.p_final && !a.z_mesh && a.p0_stride == a.nx && a.nx > 256 &&
            a.nx <= kSolveRegMaxNx && a.n_samples == 0 && !a.scan_mode && !getenv("FBR_K3_NO_STEP_KERNEL"))
            return step_factors_dispatch(a.fl, a.fr, a.fg, a.p0, reinterpret_cast<double *>(a.p_final), a.M, a.nx, a.lbc, a.rbc,
                                         a.src_idx, a.n_src, a.src_table, stream);
    }
    if (a.snap_wide) {  // fbr_run_opts::snap_f64
        FBR_REQUIRE(sizeof(T) == 4 && a.snap && a.M <= 64 && a.nx >= 512,
                    "snap_f64 is an option of the fp32 variants for at most 64 recorded members of >= 512 cells");
        return launch_x<T>(a, stream);
    }
    int cpt = a.nx >= 256 ? 8 : a.nx >= 128 ? 4 : a.nx >= 64 ? 2 : 1;
    // a few members of 128..255 cells (C1: one model of 200 cells) are a latency problem: shorter chains over more warps
    // (2 cells per thread, 4 warps: 0.48 us per step against 0.52 with 4 cells per thread and 0.58 with 8)
    if (a.M < 32 && a.nx >= 128 && a.nx < 256) cpt = 2;
    if (const char *e = getenv("FBR_CPT")) {  // tuning aid
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4 || v == 8) && (a.nx + v - 1) / v <= 32 * kMaxWarps) cpt = v;
    }
    // fp64 above 128 cells / fp32 mode from 512 cells: the 64-byte-per-thread kernel (a partly filled warp of 8-cell
    // threads still beats narrower chunks: 32768 members x 200 cells 17.1 -> 10.3 ms)
    const char *min_env = getenv("FBR_K3_X_MIN_NX");  // tuning aid
    const int64_t x_min = min_env ? atoll(min_env) : (sizeof(T) == 8 ? 129 : 512);
    // (a handful of members below 256 cells is a latency problem, not a throughput one: the generic kernel's shorter
    // chains over more threads win there — C1, one 200-cell model: 0.50 against 0.75 us per step)
    const bool few_short = sizeof(T) == 8 && a.nx < 256 && a.M < 32 && !min_env;
    if constexpr (sizeof(T) == 8) {
        if (few_short && a.z_mesh && a.n_samples == 0) {  // one launch for a short single model (on-chip assembly, fp64)
            switch (cpt) {
                case 1: return launch_cpt<T, 1, true>(a, stream);
                case 2: return launch_cpt<T, 2, true>(a, stream);
                case 4: return launch_cpt<T, 4, true>(a, stream);
                default: return launch_cpt<T, 8, true>(a, stream);
            }
        }
    }
    // observations on their own time axis are evaluated by the 64-byte-per-thread kernel only
    const bool needs_x = a.n_samples > 0;
    if (needs_x || (a.nx >= x_min && !few_short && !getenv("FBR_K3_GENERIC"))) return launch_x<T>(a, stream);
    // short meshes, more than a handful of members: groups of 8 / 16 lanes per member, several members per warp
    if (a.nx > 8 && a.nx <= 128 && a.M >= 8 && !a.z_mesh && !getenv("FBR_K3_GENERIC") && !getenv("FBR_K3_NO_GROUPS")) {
        if (a.nx <= 16) return launch_group<T, 2>(a, stream);
        if (a.nx <= 32) return launch_group<T, 4>(a, stream);
        return a.nx <= 64 ? launch_group<T, 8>(a, stream) : launch_group<T, 16>(a, stream);
    }
    if constexpr (sizeof(T) == 4) {
        if (a.M < 32) {  // fp32 mode, a few short models: float state, fp64 factors and arithmetic (see run_onchip_kernel)
            switch (cpt) {
                case 1: return launch_cpt<T, 1, false, true>(a, stream);
                case 2: return launch_cpt<T, 2, false, true>(a, stream);
                case 4: return launch_cpt<T, 4, false, true>(a, stream);
                default: return launch_cpt<T, 8, false, true>(a, stream);
            }
        }
    }
    switch (cpt) {
        case 1: return launch_cpt<T, 1>(a, stream);
        case 2: return launch_cpt<T, 2>(a, stream);
        case 4: return launch_cpt<T, 4>(a, stream);
        default: return launch_cpt<T, 8>(a, stream);
    }
}

// fbr_run_opts -> RunArgs (NULL: all defaults).  Callers compiled against an older, shorter struct are accepted.
static int apply_opts(RunArgs &a, const fbr_run_opts *opts) {
    fbr_run_opts o;
    memset(&o, 0, sizeof(o));
    if (opts) {
        FBR_REQUIRE(opts->struct_bytes >= 8 && opts->struct_bytes <= (int)sizeof(o), "fbr_run_opts.struct_bytes=%d is not a size this library knows",
                    opts->struct_bytes);
        memcpy(&o, opts, (size_t)opts->struct_bytes);
    }
    FBR_REQUIRE(o.reserve_ctas >= 0, "cannot reserve a negative number of CTA slots");
    a.reserve_ctas = o.reserve_ctas;
    a.snap_cell_stride = o.snap_cell_stride == 0 ? 1 : o.snap_cell_stride;
    const double log2_eps = o.scan_log2_eps == 0.0 ? -80.0 : o.scan_log2_eps;
    a.scan_eps = log2_eps > 0.0 ? -1.0 : exp2(log2_eps);
    if (getenv("FBR_K3_EXACT")) a.scan_eps = -1.0;  // tuning aid
    a.scan_mode = o.scan_mode;
    a.order = o.member_order;
    a.snap_wide = o.snap_f64 != 0;
    a.n_samples = o.n_samples;
    if (o.n_samples > 0) {
        a.sample_row = o.sample_row; a.sample_w = o.sample_w; a.row_first_sample = o.row_first_sample;
        a.base_sample = o.base_sample; a.first_sample = o.first_sample; a.last_sample = o.last_sample;
    } else {  // observations on the step times: sample j is the state after step j, all of them count, nothing subtracted
        a.sample_row = nullptr; a.sample_w = nullptr; a.row_first_sample = nullptr;
        a.base_sample = -1; a.first_sample = 1; a.last_sample = (int)(a.n_steps + 1);
    }
    return FBR_OK;
}


// fbr_rank_members_f64: members ranked by decreasing cost figure, in half-octave classes (counting sort, one CTA): the order
// a sweep hands its members out in (fbr_run_opts::member_order).  Members of a class keep no particular order.
constexpr int kRankBins = 4096;  // sign 0: 11 exponent bits + the leading mantissa bit
__device__ __forceinline__ int rank_bin(double c) {
    const long long bits = __double_as_longlong(c);
    if (!(c > 0.0)) return kRankBins - 1;  // zero, negative, NaN: the last class
    return kRankBins - 1 - (int)(bits >> 51);  // larger cost -> earlier class
}
__global__ void __launch_bounds__(1024) rank_members_kernel(const double *cost, int64_t M, int32_t *order) {
    __shared__ int s_cnt[kRankBins];
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = tid; b < kRankBins; b += 1024) s_cnt[b] = 0;
    __syncthreads();
    for (int64_t m = tid; m < M; m += 1024) atomicAdd(&s_cnt[rank_bin(cost[m])], 1);
    __syncthreads();
    // exclusive prefix sum over the classes: 4 consecutive classes per thread
    int c[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[j] = s_cnt[4 * tid + j]; sum += c[j]; }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(kFull, inc, d); if (lane >= d) inc += up; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(kFull, wi, d); if (lane >= d) wi += up; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    int start = s_warp[warp] + inc - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_cnt[4 * tid + j] = start; start += c[j]; }
    __syncthreads();
    for (int64_t m = tid; m < M; m += 1024) order[atomicAdd(&s_cnt[rank_bin(cost[m])], 1)] = (int32_t)m;
}

}  // namespace fbr

using namespace fbr;

extern "C" int fbr_rank_members_f64(const double *cost, int64_t M, int32_t *order, void *stream) {
    FBR_REQUIRE(cost && order, "NULL argument");
    FBR_REQUIRE(M >= 1 && M <= 0x7fffffff, "M = %lld out of range", (long long)M);
    rank_members_kernel<<<1, 1024, 0, as_stream(stream)>>>(cost, M, order);
    return check_launch("fbr_rank_members_f64");
}

extern "C" int64_t fbr_run_max_nx(void) { return 8 * 32 * (int64_t)kMaxWarps; }

#define FBR_FILL_ARGS()                                                                                   \
    RunArgs a;                                                                                            \
    a.fl = fl; a.fr = fr; a.fg = fg; a.p0 = p0; a.p0_stride = p0_stride; a.M = M; a.nx = nx;              \
    a.n_steps = n_steps; a.lbc = lbc; a.rbc = rbc; a.src_idx = src_idx; a.n_src = n_src;                  \
    a.src_table = src_table; a.snapshot_every = snapshot_every; a.snap_models = snap_models;              \
    a.n_snap_models = n_snap_models; a.snap_row_stride = snap_row_stride;                                  \
    a.snap_model_stride = snap_model_stride; a.snap = snap; a.p_final = p_final; a.obs_idx = obs_idx;             \
    a.n_obs = obs_idx ? n_obs : 0; a.obs = obs; a.obs_w = obs_w; a.misfit = misfit; a.blk_steps = 0; a.prefetch_rows = 0;     \
    a.z_mesh = nullptr; a.z_value = nullptr; a.z_of_cell = nullptr; a.z_n = 0; a.z_dt = 0.0;              \
    a.z_sigma = nullptr; a.z_singular = nullptr; a.ticket = nullptr;                                     \
    if (int rc_opts = apply_opts(a, opts)) return rc_opts;

extern "C" int fbr_run_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                           int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                           const int64_t *src_idx, int n_src, const double *src_table, int64_t snapshot_every,
                           const int64_t *snap_models, int64_t n_snap_models, int64_t snap_row_stride,
                           int64_t snap_model_stride, double *snap, double *p_final,
                           const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                           double *misfit, const fbr_run_opts *opts, void *stream) {
    FBR_FILL_ARGS();
    return run_dispatch<double>(a, as_stream(stream));
}

extern "C" int fbr_run_f32(const double *fl, const double *fr, const double *fg, const double *p0,
                           int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                           const int64_t *src_idx, int n_src, const double *src_table, int64_t snapshot_every,
                           const int64_t *snap_models, int64_t n_snap_models, int64_t snap_row_stride,
                           int64_t snap_model_stride, float *snap, float *p_final,
                           const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                           double *misfit, const fbr_run_opts *opts, void *stream) {
    FBR_FILL_ARGS();
    return run_dispatch<float>(a, as_stream(stream));
}

extern "C" int fbr_run_zonal_f64(const double *mesh, const double *zone_value, int n_zones, const int32_t *zone_of_cell,
                                 double dt, const double *sigma, int32_t *singular, const double *p0, int64_t p0_stride,
                                 int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                 const double *src_table, int64_t snapshot_every, const int64_t *snap_models,
                                 int64_t n_snap_models, int64_t snap_row_stride, int64_t snap_model_stride, double *snap,
                                 double *p_final, const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                                 double *misfit, const fbr_run_opts *opts, void *stream) {
    const double *fl = nullptr, *fr = nullptr, *fg = nullptr;
    FBR_FILL_ARGS();
    FBR_REQUIRE(mesh && zone_value && n_zones >= 1, "NULL mesh / zone arguments");
    FBR_REQUIRE(zone_of_cell || n_zones == nx, "zone_of_cell == NULL means one diffusivity per cell: n_zones must equal nx");
    FBR_REQUIRE((nx > 128 || (M < 32 && nx >= 2)) && nx <= fbr_run_max_nx(),
                "on-chip assembly covers 128 < nx <= %lld (a few short models: 2 <= nx), got nx = %lld, M = %lld",
                (long long)fbr_run_max_nx(), (long long)nx, (long long)M);
    a.z_mesh = mesh; a.z_value = zone_value; a.z_of_cell = zone_of_cell; a.z_n = n_zones; a.z_dt = dt; a.z_sigma = sigma;
    a.z_singular = singular;
    return run_dispatch<double>(a, as_stream(stream));
}

extern "C" int fbr_run_zonal_f32(const double *mesh, const double *zone_value, int n_zones, const int32_t *zone_of_cell,
                                 double dt, const double *sigma, int32_t *singular, const double *p0, int64_t p0_stride,
                                 int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                 const double *src_table, int64_t snapshot_every, const int64_t *snap_models,
                                 int64_t n_snap_models, int64_t snap_row_stride, int64_t snap_model_stride, float *snap,
                                 float *p_final, const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                                 double *misfit, const fbr_run_opts *opts, void *stream) {
    const double *fl = nullptr, *fr = nullptr, *fg = nullptr;
    FBR_FILL_ARGS();
    FBR_REQUIRE(mesh && zone_value && n_zones >= 1, "NULL mesh / zone arguments");
    FBR_REQUIRE(zone_of_cell || n_zones == nx, "zone_of_cell == NULL means one diffusivity per cell: n_zones must equal nx");
    FBR_REQUIRE(nx >= 512 && nx <= fbr_run_max_nx(), "fp32 on-chip assembly covers 512 <= nx <= %lld (got %lld)",
                (long long)fbr_run_max_nx(), (long long)nx);
    a.z_mesh = mesh; a.z_value = zone_value; a.z_of_cell = zone_of_cell; a.z_n = n_zones; a.z_dt = dt; a.z_sigma = sigma;
    a.z_singular = singular;
    return run_dispatch<float>(a, as_stream(stream));
}

#ifdef FBR_K3_PROF
extern "C" int fbr_debug_k3_prof(long long *host, int reset) {
    if (reset) { long long z[8 * 8 * 4] = {0}; return (int)cudaMemcpyToSymbol(fbr::g_k3_prof, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(host, fbr::g_k3_prof, sizeof(fbr::g_k3_prof));
}
#endif
