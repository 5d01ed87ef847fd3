// K3L  persistent time integration of ONE LONG mesh spread over the whole GPU (cooperative launch).
//
// Same arithmetic as K3 (fbr_run_onchip.cu): pressure and the three factor arrays of every cell stay
// in registers for all time steps; a sweep is local pass -> warp scan -> cross-warp carry -> second
// local pass.  What changes is one more level of the carry hierarchy: a CTA owns one TILE of
// 32*W*CPT consecutive cells, all tiles are co-resident (cudaLaunchCooperativeKernel), and the carry
// between tiles goes through global memory (L2) with one grid-wide barrier per sweep:
//     tile k publishes its zero-carry end value E_k;  barrier;
//     C_k = sum_{v<k} Kg[k][v] * E_v   with Kg = products of whole-tile multipliers (time-invariant).
// Products of |multipliers| < 1 underflow to an exact 0 after a few tiles, so a row of Kg has a short
// non-zero "horizon" and the dot product only visits that (skipping exact zeros is exact).
// Covers meshes up to (resident CTAs) * 1024 cells (~6e5 on a B200): BASELINE.json configs[1] (1e5 cells).
#include <cooperative_groups.h>

#include "fbr_common.cuh"

namespace cg = cooperative_groups;

namespace fbr {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kCoopCPT = 8;
constexpr int kCoopW = 4;
constexpr int kCoopTile = 32 * kCoopW * kCoopCPT;  // cells per CTA
constexpr int kRing = 16;      // published tile-end values are kept for this many steps
constexpr int kMaxP2P = 12;    // widest carry horizon served by point-to-point flags (else grid barrier)

// Flagged 16-byte message (the "LL" idea of collective libraries): a double travels as two 8-byte
// words, each carrying half of the value and the 32-bit step number.  8-byte stores are atomic, so a
// reader that sees the expected step in BOTH words has the whole value — no fences on either side.
__device__ __forceinline__ void ll_publish(ulonglong2 *slot, double v, unsigned step) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    ulonglong2 m;
    m.x = (bits & 0xffffffffull) | ((unsigned long long)step << 32);
    m.y = (bits >> 32) | ((unsigned long long)step << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(m.x), "l"(m.y) : "memory");
}
__device__ __forceinline__ double ll_wait(const ulonglong2 *slot, unsigned step) {
    unsigned long long a, b;
    do {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(slot) : "memory");
    } while ((unsigned)(a >> 32) != step || (unsigned)(b >> 32) != step);
    return __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
}

struct CoopArgs {
    const double *fl, *fr, *fg;  // (nx) factors of this model
    const double *p0;            // (nx)
    int64_t nx, n_steps;
    int lbc, rbc;
    const int64_t *src_idx;
    int n_src;
    const double *src_table;  // (n_steps, n_src)
    int64_t snapshot_every;
    double *snap;  // (n_steps / snapshot_every, nx) rows of this model, or NULL
    int64_t snap_row_stride;
    double *p_final;  // (nx) or NULL
    double *work;     // Ttile[2][n_tiles], E[2][n_tiles], message ring[2][kRing][n_tiles] (16 B each), max horizon
    int n_tiles;
};

__device__ const double kZeroCoop = 0.0;

__device__ __forceinline__ void mov_if_eq_d(double &dst, double v, int j, int k) {
    asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, %3; @p mov.f64 %0, %1; }" : "+d"(dst) : "d"(v), "r"(j), "r"(k));
}

__global__ void __launch_bounds__(32 * kCoopW, 4) run_coop_kernel(const CoopArgs a) {
    constexpr int CPT = kCoopCPT, W = kCoopW;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double s_row[];  // [2][n_tiles] rows of the tile-level K tables
    __shared__ __align__(16) double s_K[2][W][W];
    __shared__ __align__(16) double s_E[2][W];
    __shared__ double s_T[2][W];
    __shared__ double s_pre[2][W];  // product of the warp totals before (fwd) / after (bwd) each warp
    __shared__ double s_carry[2];
    __shared__ int s_hor[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, n_tiles = a.n_tiles;
    const int64_t nx = a.nx;
    const int64_t cell0 = ((int64_t)tile * (32 * W) + tid) * CPT;
    double *Tg = a.work;                                    // [2][n_tiles]
    double *Eg = a.work + 2 * n_tiles;                      // [2][n_tiles]  (grid-barrier mode)
    ulonglong2 *ring = reinterpret_cast<ulonglong2 *>(Eg + 2 * n_tiles);  // [2][kRing][n_tiles] flagged messages
    long long *max_hor = reinterpret_cast<long long *>(ring + 2 * kRing * n_tiles);
    double *Kf = s_row, *Kb = s_row + n_tiles;

    // ---------------- setup ----------------
    double l[CPT], r[CPT], g[CPT], x[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int64_t i = cell0 + j;
        if (i < nx) {
            l[j] = __ldg(a.fl + i); r[j] = __ldg(a.fr + i); g[j] = __ldg(a.fg + i); x[j] = __ldg(a.p0 + i);
        } else {
            l[j] = 0.0; r[j] = 1.0; g[j] = 0.0; x[j] = 0.0;
        }
    }
    // right-hand-side overrides owned by this thread (same scheme as fbr_run_onchip.cu)
    int ov_j = -1, ov_stride = 0;
    bool ov_slow = false;
    const double *ovp = &kZeroCoop;
    if (cell0 < nx) {
        constexpr unsigned kNoSlot = 0xFF, kZeroSlot = 0xFE;
        uint64_t ov = ~0ull;  // one slot byte per owned cell
        if (cell0 == 0 && a.lbc != FBR_BC_NONE) ov = (ov & ~0xFFull) | kZeroSlot;
        const int64_t jl = nx - 1 - cell0;
        if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) ov = (ov & ~(0xFFull << (8 * jl))) | ((uint64_t)kZeroSlot << (8 * jl));
        bool wide = false;  // a source index that does not fit a slot byte
#pragma unroll 1
        for (int k = 0; k < a.n_src; ++k) {
            const int64_t jj = a.src_idx[k] - cell0;
            if (jj >= 0 && jj < CPT && a.src_idx[k] < nx) {
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
                else { ovp = &kZeroCoop; ov_stride = 0; }
            }
        }
        if (n_ov > 1 || wide) { ov_slow = true; ov_j = -1; ovp = &kZeroCoop; ov_stride = 0; }
    }
    // scan multipliers
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
            const double upf = __shfl_up_sync(kFullMask, pf, 1 << s);
            const double dnb = __shfl_down_sync(kFullMask, pb, 1 << s);
            if (fa) pf *= upf;
            if (ba) pb *= dnb;
        }
        PfEx = __shfl_up_sync(kFullMask, pf, 1);
        PbEx = __shfl_down_sync(kFullMask, pb, 1);
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
            s_pre[0][tid] = k;  // product of all warp totals before warp tid
            k = 1.0;
            for (int v = 0; v < W; ++v) {
                if (v <= tid) s_K[1][tid][v] = 0.0;
                else { s_K[1][tid][v] = k; k *= s_T[1][v]; }
            }
            s_pre[1][tid] = k;  // product of all warp totals after warp tid
        }
        if (tid == 0) {  // whole-tile multipliers
            double tf = 1.0, tb = 1.0;
            for (int v = 0; v < W; ++v) { tf *= s_T[0][v]; tb *= s_T[1][v]; }
            Tg[tile] = tf;
            Tg[n_tiles + tile] = tb;
            for (int q = 0; q < 2 * kRing; ++q) ring[q * n_tiles + tile] = make_ulonglong2(0ull, 0ull);
            if (tile == 0) *max_hor = 0;
        }
    }
    __threadfence();
    grid.sync();
    // rows of the tile-level K tables, and how far back / forward they are non-zero
    if (tid == 0) {
        double k = 1.0;
        int hor = 0;
        for (int v = tile - 1; v >= 0; --v) {
            Kf[v] = k;
            if (k != 0.0) hor = tile - v;
            k *= __ldcg(Tg + v);
        }
        s_hor[0] = hor;
    }
    if (tid == 32) {
        double k = 1.0;
        int hor = 0;
        for (int v = tile + 1; v < n_tiles; ++v) {
            Kb[v] = k;
            if (k != 0.0) hor = v - tile;
            k *= __ldcg(Tg + n_tiles + v);
        }
        s_hor[1] = hor;
    }
    __syncthreads();
    if (tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(max_hor), (unsigned long long)max(s_hor[0], s_hor[1]));
    __threadfence();
    grid.sync();
    // short horizons (the normal case): tiles synchronise point to point through per-tile flags;
    // otherwise every sweep ends in a grid-wide barrier
    const bool p2p = *reinterpret_cast<volatile long long *>(max_hor) <= kMaxP2P;
    const int hor_f = s_hor[0], hor_b = s_hor[1];
    unsigned step = 0;
    const double pre_f = s_pre[0][warp], pre_b = s_pre[1][warp];

    const int snap_every = a.snapshot_every > 0x7fffffff ? 0x7fffffff : (int)a.snapshot_every;
    int snap_countdown = a.snap ? snap_every : 0x7fffffff;
    double *snap_dst = a.snap + cell0;
    double ov_val = a.n_steps > 0 ? __ldg(ovp) : 0.0;

    // ---------------- time loop ----------------
#pragma unroll 1
    for (int left = (int)a.n_steps; left > 0; --left) {
        if (left == 1) ov_stride = 0;
        ovp += ov_stride;
        const double ov_next = __ldg(ovp);
        if (ov_j >= 0) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) mov_if_eq_d(x[k], ov_val, ov_j, k);
        }
        if (ov_slow) {
            const int64_t n = a.n_steps - left;
            if (cell0 == 0 && a.lbc != FBR_BC_NONE) x[0] = 0.0;
            const int64_t jl = nx - 1 - cell0;
            if (jl >= 0 && jl < CPT && a.rbc != FBR_BC_NONE) {
#pragma unroll
                for (int k = 0; k < CPT; ++k) mov_if_eq_d(x[k], 0.0, (int)jl, k);
            }
#pragma unroll 1
            for (int q = 0; q < a.n_src; ++q) {
                const int64_t jj = a.src_idx[q] - cell0;
                if (jj >= 0 && jj < CPT) {
                    const double v = __ldg(a.src_table + n * a.n_src + q);
#pragma unroll
                    for (int k = 0; k < CPT; ++k) mov_if_eq_d(x[k], v, (int)jj, k);
                }
            }
        }
        // ---- forward sweep ----
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < CPT; ++j) y = fma(-l[j], y, x[j]);
#pragma unroll
        for (int s = 0; s < 5; ++s) y = fma(PfL[s], __shfl_up_sync(kFullMask, y, 1 << s), y);
        double carry = __shfl_up_sync(kFullMask, y, 1);
        if (lane == 0) carry = 0.0;
        if (lane == 31) s_E[0][warp] = y;
        __syncthreads();
        double c = 0.0;
#pragma unroll
        for (int v = 0; v < W - 1; ++v) c = fma(s_K[0][warp][v], s_E[0][v], c);
        ++step;
        ulonglong2 *rf = ring + (step % kRing) * n_tiles, *rb = ring + (kRing + step % kRing) * n_tiles;
        if (tid == 32 * W - 1) {  // zero-carry end value of the whole tile
            const double e = fma(s_T[0][W - 1], c, y);
            if (p2p) ll_publish(rf + tile, e, step);
            else { Eg[tile] = e; __threadfence(); }
        }
        if (!p2p) grid.sync();
        if (warp == 0) {  // tile-level carry: dot product over the non-zero horizon of the K row
            double ck = 0.0;
            for (int d = 1 + lane; d <= hor_f; d += 32)
                ck = fma(Kf[tile - d], p2p ? ll_wait(rf + tile - d, step) : __ldcg(Eg + tile - d), ck);
            for (int s = 16; s > 0; s >>= 1) ck += __shfl_xor_sync(kFullMask, ck, s);
            if (lane == 0) s_carry[0] = ck;
        }
        __syncthreads();
        c = fma(pre_f, s_carry[0], c);
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
        for (int s = 0; s < 5; ++s) z = fma(PbL[s], __shfl_down_sync(kFullMask, z, 1 << s), z);
        carry = __shfl_down_sync(kFullMask, z, 1);
        if (lane == 31) carry = 0.0;
        if (lane == 0) s_E[1][warp] = z;
        __syncthreads();
        c = 0.0;
#pragma unroll
        for (int v = 1; v < W; ++v) c = fma(s_K[1][warp][v], s_E[1][v], c);
        if (tid == 0) {
            const double e = fma(s_T[1][0], c, z);
            if (p2p) ll_publish(rb + tile, e, step);
            else { Eg[n_tiles + tile] = e; __threadfence(); }
        }
        if (!p2p) grid.sync();
        if (warp == 0) {
            double ck = 0.0;
            for (int d = 1 + lane; d <= hor_b; d += 32)
                ck = fma(Kb[tile + d], p2p ? ll_wait(rb + tile + d, step) : __ldcg(Eg + n_tiles + tile + d), ck);
            for (int s = 16; s > 0; s >>= 1) ck += __shfl_xor_sync(kFullMask, ck, s);
            if (lane == 0) s_carry[1] = ck;
        }
        __syncthreads();
        c = fma(pre_b, s_carry[1], c);
        z = fma(PbEx, c, carry);
#pragma unroll
        for (int j = CPT - 1; j >= 0; --j) {
            z = fma(-g[j], z, x[j]);
            x[j] = z;
        }
        if (--snap_countdown == 0) {
            snap_countdown = snap_every;
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (cell0 + j < nx) snap_dst[j] = x[j];
            snap_dst += a.snap_row_stride;
        }
        ov_val = ov_next;
    }
    if (a.p_final) {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
            if (cell0 + j < nx) a.p_final[cell0 + j] = x[j];
    }
}

}  // namespace fbr

using namespace fbr;

static int coop_capacity(int *max_tiles) {
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    int occ = 0;
    // dynamic smem depends on the tile count; query with the largest row we could need
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, run_coop_kernel, 32 * kCoopW, 2 * 4 * 148 * sizeof(double)));
    *max_tiles = occ * f.sm_count;
    return FBR_OK;
}

extern "C" int64_t fbr_run_long_max_nx(void) {
    int tiles = 0;
    if (coop_capacity(&tiles) != FBR_OK) return 0;
    return (int64_t)tiles * kCoopTile;
}

extern "C" int64_t fbr_run_long_work_bytes(int64_t nx) {
    const int64_t tiles = (nx + kCoopTile - 1) / kCoopTile;
    return ((4 + 4 * kRing) * tiles + 2) * (int64_t)sizeof(double);
}

extern "C" int fbr_run_long_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t M,
                                int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                                int64_t snap_model_stride, double *snap, double *p_final, double *work, void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 1 && n_steps >= 0 && n_steps <= 0x7fffffff, "bad M / nx / n_steps");
    FBR_REQUIRE(fl && fr && fg && p0 && work, "NULL factor / initial-state / work pointer");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_table)), "sources given without src_idx / src_table");
    FBR_REQUIRE(!snap || (snapshot_every >= 1 && snap_row_stride >= nx && snap_model_stride >= nx), "bad snapshot layout");
    int max_tiles = 0;
    if (int rc = coop_capacity(&max_tiles)) return rc;
    const int64_t tiles = (nx + kCoopTile - 1) / kCoopTile;
    if (tiles > max_tiles)
        return fail(FBR_ERR_UNSUPPORTED, "nx=%lld needs %lld co-resident tiles, the device holds %d", (long long)nx,
                    (long long)tiles, max_tiles);
    const size_t smem = 2 * (size_t)tiles * sizeof(double);
    for (int64_t m = 0; m < M; ++m) {
        CoopArgs a;
        a.fl = fl + m * nx; a.fr = fr + m * nx; a.fg = fg + m * nx; a.p0 = p0 + m * nx;
        a.nx = nx; a.n_steps = n_steps; a.lbc = lbc; a.rbc = rbc; a.src_idx = src_idx; a.n_src = n_src;
        a.src_table = src_table; a.snapshot_every = snapshot_every;
        a.snap = snap ? snap + m * snap_model_stride : nullptr;
        a.snap_row_stride = snap_row_stride;
        a.p_final = p_final ? p_final + m * nx : nullptr;
        a.work = work; a.n_tiles = (int)tiles;
        void *params[] = {(void *)&a};
        FBR_CUDA(cudaLaunchCooperativeKernel((const void *)run_coop_kernel, dim3((unsigned)tiles), dim3(32 * kCoopW),
                                             params, smem, as_stream(stream)));
        if (int rc = check_launch("fbr_run_long_f64")) return rc;
    }
    return FBR_OK;
}
