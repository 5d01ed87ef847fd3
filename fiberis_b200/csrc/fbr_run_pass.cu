// K3P  exact time integration of meshes of ANY length in ONE pass over HBM per step: 40 B per cell-update
// (three factors + p^n in, p^{n+1} out).  Replaces the level hierarchy of K3S (88 B per cell-update) as the
// exact fallback behind the windowed kernel K3W — stiff long meshes, whose decay horizon is too long for windows.
//
// The mesh is cut into tiles of 256 W cells (W = 2, 4 or 8, see kPWMin); a CTA of W warps solves a tile in registers with
// the sweeps of K3 (8 cells per thread, warp scans, exact carry table across the warps).  What couples the tiles is one number per
// tile and sweep direction — the carry — and both carries are obtained without a second pass over the data:
//   * forward carry: y at the last cell of the tile on the left.  It depends only on p^n, which the previous launch
//     finished: every tile leaves, next to p^{n+1}, the zero-carry end value of ITS forward sweep for the NEXT step
//     (Ef) — the next launch folds the Ef of the tiles on its left with the time-invariant tile products Mf
//     (exact: the fold stops when the running product underflows to exactly 0, or at tile 0);
//   * backward carry: x at the first cell of the tile on the right.  It depends on this step's forward results, so it
//     is resolved inside the launch by a decoupled look-back: tiles are handed out right to left (atomic ticket), a
//     tile publishes its zero-carry aggregate as soon as its own forward sweep is done (its right-hand neighbours
//     hold earlier tickets: they run or are done) and folds the aggregates on its right exactly like the forward ones
//     — always the same association, so the result does not depend on the schedule (value and epoch-tagged status in
//     one 16-byte word, no reset between launches).
// One launch per time step; snapshot rows are written in place and read back as the next step's input.
// Latency is what this pipeline fights (a tile is ~2.5 us of dependent work, 16 warps per SM): the next tile's four rows
// stream into shared memory while the current one is solved — four 16 KB BULK copies (cp.async.bulk, the 1-D TMA path,
// completion on an mbarrier) issued by one thread: 16-byte cp.async by every thread (16 per thread and tile) topped out
// at 4.0 TB/s in isolation against 6.9 TB/s for the bulk copies, conflicting shared-memory reads of the linear rows
// included (scratch/ubench_bulk.cu, profiles/r02_ubench_bulk_copy.txt) —, its ticket is drawn a tile ahead, the
// forward fold is issued before the rows are touched, and a tile's aggregate travels with its status in ONE 16-byte
// word (one store to publish, one load to poll — no fence, no second round trip).
#include "fbr_common.cuh"
#include "fbr_sweep.cuh"

namespace fbr {

// Warps per tile W (a tile = 256 W cells, 16 / W CTAs per SM): what couples the tiles costs per TILE, what a tile costs
// grows with its width.  Ordinary meshes (the fold ends at the neighbouring tile): narrow tiles win — C5's mesh 89 / 92 /
// 101 us per step for W = 2 / 4 / 8; stiff meshes (K3P's own domain: the fold spans tens of tiles) want few, wide tiles —
// 4e6 cells at alpha ~ 3e4: 154 / 102 / 88 us.  The caller says which (fbr_run_pass_tiled_f64); default 8.
constexpr int kPWMin = 2;  // narrowest tile: sizes the workspace

#ifdef FBR_K3P_PROF  // debug build: clock64 stamps of a tile's phases for a few CTAs (scratch/k3p_prof.py reads them)
__device__ long long g_k3p_prof[8 * 24 * 12];
#define K3P_STAMP(k) do { if (tid == 0 && blockIdx.x % 37 == 0 && blockIdx.x / 37 < 8 && it < 24) g_k3p_prof[((blockIdx.x / 37) * 24 + it) * 12 + (k)] = clock64(); } while (0)
#else
#define K3P_STAMP(k) do { } while (0)
#endif

struct PassArgs {
    const double *fl, *fr, *fg;
    const double *p_in;
    double *p_out;
    int64_t nx, n_tiles;
    int lbc, rbc;
    const int64_t *src_idx;
    int n_src;
    const double *src_now, *src_next;  // (n_src) values overriding the right-hand side of this / the next step (next: NULL at the end)
    double *Mf, *Mb;                   // (n_tiles) products of the multipliers over each tile (written by the setup pass)
    const double *Ef_now;              // (n_tiles) zero-carry forward end value of every tile for THIS step
    double *Ef_next;                   // (n_tiles) the same for the next step
    double2 *Lw;                       // (n_tiles) zero-carry backward aggregates of this launch as (value, epoch * 4 + 1) words
    unsigned *ticket;                  // one counter, never reset: every launch adds n_tiles + gridDim.x to it (each CTA's last,
    unsigned ticket_base;              // failing draw included), so launch e starts at e * (n_tiles + grid) = ticket_base
    unsigned epoch;                    // launch number, from 1 (0 = the setup launch)
};

// (value, tag) in one naturally aligned 16-byte word: written and read by single 128-bit accesses that bypass L1
__device__ __forceinline__ void st_word(double2 *p, double v, unsigned tag) {
    asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(__hiloint2double(0, (int)tag)) : "memory");
}
__device__ __forceinline__ unsigned ld_word(const double2 *p, double &v) {
    double t;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(t) : "l"(p) : "memory");
    return (unsigned)__double2loint(t);
}
// b = p with this step's overrides on my 8 cells (matbuilder.py:45-76: boundary rows first, sources in list order).
// tile_has_src: some source lies in this CTA's tile (almost never: the scan over the source list is skipped otherwise)
__device__ __forceinline__ void apply_overrides(double (&x)[8], int64_t cell0, int64_t nx, int lbc, int rbc,
                                                const int64_t *__restrict__ src_idx, int n_src, const double *__restrict__ src_val,
                                                bool tile_has_src) {
    if (cell0 == 0 && lbc != FBR_BC_NONE) x[0] = 0.0;
    const int64_t jl = nx - 1 - cell0;
    if (jl >= 0 && jl < 8 && rbc != FBR_BC_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j == jl) x[j] = 0.0;
    }
    if (!tile_has_src) return;
#pragma unroll 1
    for (int k = 0; k < n_src; ++k) {
        const int64_t jj = __ldg(src_idx + k) - cell0;
        if (jj >= 0 && jj < 8 && __ldg(src_idx + k) < nx) {
            const double v = __ldg(src_val + k);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j == jj) x[j] = v;
        }
    }
}

// SETUP: only the tile products and the forward aggregates of step 0 are produced (p_out, look-back untouched).
template <bool SETUP, int W>
__global__ void __launch_bounds__(32 * W, 16 / W) pass_kernel(const PassArgs a) {
    constexpr int kPTile = 256 * W;  // cells per tile
    __shared__ double s_E[2][W];        // zero-carry warp totals of the current sweep
    __shared__ double s_T[2][W];        // products of the multipliers over each warp
    __shared__ double s_carry[2];       // carries into the tile: forward (from the left), backward (from the right)
    __shared__ unsigned s_ticket[2];
    extern __shared__ __align__(128) double s_stage[];  // [4 rows][2048] doubles: the NEXT tile's rows (bulk copies)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned base = a.ticket_base;
    const int64_t nx = a.nx;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.fl) | reinterpret_cast<uintptr_t>(a.fr) | reinterpret_cast<uintptr_t>(a.fg) |
                           reinterpret_cast<uintptr_t>(a.p_in)) & 15) == 0;
    // rows of tile `t` -> shared memory, linear (row q at s_stage + q * 2048 doubles): only whole, aligned tiles are staged
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned bar = smem_addr(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto staged = [&](int64_t t) { return !SETUP && aligned && t >= 0 && (t + 1) * kPTile <= nx; };
    // (callers: every thread, after everybody has read the previous tile out of the buffer; lane 0 of warps 0 .. min(W, 4) - 1
    //  issues — four copies from one thread kept warp 0 for ~470 cycles, scratch/k3p_prof.py)
    auto stage_tile = [&](int64_t t) {
        if (lane != 0 || !staged(t)) return;
        const int64_t c0 = t * kPTile;
        const double *src[4] = {a.fl + c0, a.fr + c0, a.fg + c0, a.p_in + c0};
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before async-proxy writes
        if (warp == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4u * kPTile * 8u) : "memory");
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (warp == (W >= 4 ? q : (q & 1))) bulk_g2s(smem_addr(s_stage) + q * kPTile * 8, src[q], kPTile * 8, bar);
    };
    unsigned n_staged = 0;  // staged tiles consumed so far: the mbarrier's phase
    // the source cells, once per CTA: which tiles hold one is then a shared-memory question
    __shared__ int64_t s_src[256];
    for (int k = tid; k < a.n_src && k < 256; k += 32 * W) s_src[k] = __ldg(a.src_idx + k);

    if (tid == 0) s_ticket[0] = atomicAdd(a.ticket, 1u) - base;
    __syncthreads();
    if (s_ticket[0] < (unsigned)a.n_tiles) stage_tile(a.n_tiles - 1 - (int64_t)s_ticket[0]);
    // Forward carry into tile t: fold the finished aggregates of the tiles on its left (exact: it stops at tile 0 or when
    // the running product underflows to 0).  One warp's job — warp 1 does it for the NEXT tile while warp 0 looks back for
    // the current one, so nobody waits for it (at the head of a tile it was 17 % of the stall samples).
    auto forward_fold = [&](int64_t t) {
        double c = 0.0, prod = 1.0;
        for (int64_t j0 = t - 1; j0 >= 0 && prod != 0.0; j0 -= 32) {
            const int64_t j = j0 - lane;  // lane 0 = nearest tile
            double acc = j >= 0 ? __ldg(a.Ef_now + j) : 0.0;  // value carried out of the nearest tile by lanes 0 .. lane
            double pm = j >= 0 ? __ldg(a.Mf + j) : 0.0;       // product of the multipliers of lanes 0 .. lane
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {  // affine maps composed nearest-first
                const double up_acc = __shfl_up_sync(kAllLanes, acc, s), up_pm = __shfl_up_sync(kAllLanes, pm, s);
                if (lane >= s) { acc = fma(up_pm, acc, up_acc); pm *= up_pm; }
            }
            c = fma(prod, __shfl_sync(kAllLanes, acc, 31), c);
            prod *= __shfl_sync(kAllLanes, pm, 31);
        }
        return c;
    };
    __shared__ double s_cf[2];  // forward carry of the current / the next tile
    if (!SETUP && warp == 1 && s_ticket[0] < (unsigned)a.n_tiles) {
        const double c = forward_fold(a.n_tiles - 1 - (int64_t)s_ticket[0]);
        if (lane == 0) s_cf[0] = c;
    }
    for (int it = 0;; ++it) {
        const unsigned ticket = s_ticket[it & 1];
        if (ticket >= (unsigned)a.n_tiles) break;
        if (tid == 0) s_ticket[(it & 1) ^ 1] = atomicAdd(a.ticket, 1u) - base;  // the next tile's ticket, a tile ahead
        const int64_t tile = a.n_tiles - 1 - (int64_t)ticket;  // right to left
        const int64_t cell0 = tile * kPTile + (int64_t)tid * 8;

        // (warp 0) products of the multipliers over each of the 32 tiles on my right: time-invariant, needed by the look-back
        double pm_near = 0.0;
        if (!SETUP && warp == 0 && tile + 1 + lane < a.n_tiles) pm_near = __ldg(a.Mb + tile + 1 + lane);
        // ---- my 8 cells: factors and state (identity rows behind the mesh) ----
        double l[8], r[8], g[8], x[8];
        K3P_STAMP(0);
        if (staged(tile)) {
            mbar_wait(bar, n_staged & 1u);
            K3P_STAMP(1);
            ++n_staged;
            // my 64 bytes of row q.  The bulk copies leave the rows linear, so the same 16-byte piece of every thread would
            // hit 2 of the 8 bank groups (4-way conflicts: 16 wavefronts per load — with 16 warps per SM a third of a tile's
            // time on the shared-memory pipe, which the sweeps' shuffles share).  Thread t therefore starts with piece
            // (t / 2) % 4 — every quarter warp then covers all 8 groups — and rotates the four pieces back into place with two
            // select stages (128 selects per tile on idle ALUs): 113 -> 95 us per step on C5's mesh.
            const int rot = (tid >> 1) & 3;
            auto take = [&](int q, double (&v)[8]) {
                double2 ld[4], mid[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    ld[j] = *reinterpret_cast<const double2 *>(s_stage + q * kPTile + tid * 8 + 2 * ((j + rot) & 3));
#pragma unroll
                for (int p = 0; p < 4; ++p) {  // ld[j] holds piece (j + rot) % 4
                    mid[p].x = (rot & 1) ? ld[(p + 3) & 3].x : ld[p].x;
                    mid[p].y = (rot & 1) ? ld[(p + 3) & 3].y : ld[p].y;
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    v[2 * p] = (rot & 2) ? mid[(p + 2) & 3].x : mid[p].x;
                    v[2 * p + 1] = (rot & 2) ? mid[(p + 2) & 3].y : mid[p].y;
                }
            };
            take(0, l); take(1, r); take(2, g); take(3, x);
        } else {
            auto load8 = [&](const double *src, double pad, double (&v)[8], bool ro) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = cell0 + j < nx ? (ro ? __ldg(src + cell0 + j) : src[cell0 + j]) : pad;
            };
            load8(a.fl, 0.0, l, true);
            load8(a.fr, 1.0, r, true);
            load8(a.fg, 0.0, g, true);
            load8(a.p_in, 0.0, x, false);  // written by the previous launch: plain (coherent) loads
        }
        K3P_STAMP(8);
        // the next ticket is visible; everybody has taken its rows out of the staging buffer; does a source lie in this tile?
        bool src_here = false;
        for (int k = tid; k < a.n_src; k += 32 * W) {
            const int64_t i = k < 256 ? s_src[k] : __ldg(a.src_idx + k);
            src_here = src_here || (i >= tile * kPTile && i < (tile + 1) * kPTile);
        }
        const bool tile_has_src = __syncthreads_or(src_here);
        K3P_STAMP(9);
        const unsigned nt = s_ticket[(it & 1) ^ 1];
        const int64_t next_tile = nt < (unsigned)a.n_tiles ? a.n_tiles - 1 - (int64_t)nt : -1;
        stage_tile(next_tile);  // (everybody passed the barrier above: the buffer is free)
        K3P_STAMP(2);
        const ScanMultipliers sm = scan_multipliers<double>(chunk_product<double, 8>(l), chunk_product<double, 8>(g), lane);
        if (lane == 31) s_T[0][warp] = sm.tot_f;
        if (lane == 0) s_T[1][warp] = sm.tot_b;

        // zero-carry forward sweep of the tile for right-hand side b: the warp-inclusive chunk-end values
        auto forward_ends = [&](const double (&b)[8]) {
            double y = chunk_end_fwd_plain<double, 8>(b, l);
#pragma unroll
            for (int s = 0; s < 5; ++s) y = fma(sm.f[s], __shfl_up_sync(kAllLanes, y, 1 << s), y);
            return y;
        };
        // value at the tile's last cell from the warps' zero-carry totals E (and a carry c into the tile)
        auto tile_forward_end = [&](double c) {
            double v = c;
            for (int w = 0; w < W; ++w) v = fma(s_T[0][w], v, s_E[0][w]);
            return v;
        };

        if (SETUP) {
            apply_overrides(x, cell0, nx, a.lbc, a.rbc, a.src_idx, a.n_src, a.src_now, tile_has_src);
            const double y = forward_ends(x);
            if (lane == 31) s_E[0][warp] = y;
            __syncthreads();
            if (tid == 0) {
                double mf = 1.0, mb = 1.0;
                for (int w = 0; w < W; ++w) { mf *= s_T[0][w]; mb *= s_T[1][w]; }
                a.Mf[tile] = mf;
                a.Mb[tile] = mb;
                a.Ef_next[tile] = tile_forward_end(0.0);
            }
            __syncthreads();  // the shared scratch is reused by the next tile
            continue;
        }

        apply_overrides(x, cell0, nx, a.lbc, a.rbc, a.src_idx, a.n_src, a.src_now, tile_has_src);
        // ---- forward sweep: y_i = b_i - l_i y_{i-1}, t_i = r_i y_i ----
        {
            const double y = forward_ends(x);
            double c = and_mask(__shfl_up_sync(kAllLanes, y, 1), lane == 0 ? 0 : -1);
            if (lane == 31) s_E[0][warp] = y;
            __syncthreads();
            double cw = s_cf[it & 1];  // y in front of my warp: the tile's carry folded through the warps on my left
            for (int w = 0; w < warp; ++w) cw = fma(s_T[0][w], cw, s_E[0][w]);
            c = fma(sm.ex_f, cw, c);
            chunk_apply_fwd_plain<double, 8>(x, l, r, c);
        }
        K3P_STAMP(3);
        // ---- backward sweep, zero carry from the right: publish the tile's aggregate, look back, apply ----
        {
            double z = chunk_end_bwd_plain<double, 8>(x, g);
#pragma unroll
            for (int s = 0; s < 5; ++s) z = fma(sm.b[s], __shfl_down_sync(kAllLanes, z, 1 << s), z);
            double c = and_mask(__shfl_down_sync(kAllLanes, z, 1), lane == 31 ? 0 : -1);
            if (lane == 0) s_E[1][warp] = z;
            __syncthreads();
            if (warp == 0) {
                if (lane == 0) {
                    double loc = 0.0;  // value at the tile's first cell for a zero carry from the right
                    for (int w = W - 1; w >= 0; --w) loc = fma(s_T[1][w], loc, s_E[1][w]);
                    st_word(a.Lw + tile, loc, a.epoch * 4u + 1u);
                }
                K3P_STAMP(4);
                // Carry from the right: fold the zero-carry aggregates of the tiles on the right, nearest first, 32 per
                // iteration, until the running product of the tile multipliers is exactly 0 (one tile on ordinary meshes) or
                // the mesh ends.  Always the same association — the result does not depend on which tiles happen to be done
                // (an inclusive value picked up when available would make the last bits a matter of timing).  The tiles on
                // the right hold earlier tickets: they are running or done, and publish before they wait for anybody.
                double cb = 0.0, prod = 1.0;
                if (tile + 1 < a.n_tiles) {
                    const unsigned want = a.epoch * 4u + 1u;
                    if (__shfl_sync(kAllLanes, pm_near, 0) == 0.0) {
                        // ordinary mesh, wide tile: the product over the nearest tile is already 0 — its aggregate is the carry
                        double v = 0.0;
                        if (lane == 0)
                            while (ld_word(a.Lw + tile + 1, v) != want) { }
                        cb = __shfl_sync(kAllLanes, v, 0);
                        prod = 0.0;
                    } else {
                        // the 32 nearest tiles at once: every lane polls its tile's word FIRST (one round trip for all of them),
                        // the time-invariant products — loaded at the head of the tile — say meanwhile which lanes matter
                        const int64_t j = tile + 1 + lane;  // lane 0 = nearest tile
                        const bool in = j < a.n_tiles;
                        double acc = 0.0, pm = pm_near;
                        unsigned tag = in ? ld_word(a.Lw + j, acc) : want;
                        double reach = pm;  // product of the multipliers of lanes 0 .. lane
#pragma unroll
                        for (int s = 1; s < 32; s <<= 1) {
                            const double up = __shfl_up_sync(kAllLanes, reach, s);
                            if (lane >= s) reach *= up;
                        }
                        double before = __shfl_up_sync(kAllLanes, reach, 1);
                        if (lane == 0) before = 1.0;
                        // something of a tile reaches me: wait for it.  ONE lane spins at a time (the nearest tile still
                        // missing), the others look again when it has arrived: 32 lanes of every CTA hammering the words the
                        // publishers are writing slowed everybody down on stiff meshes (4e6 cells, alpha ~ 3e4: 87 -> 110 us)
                        const bool need = in && before != 0.0;
                        unsigned waiting = __ballot_sync(kAllLanes, need && tag != want);
                        while (waiting) {
                            if (lane == __ffs(waiting) - 1)
                                while (tag != want) tag = ld_word(a.Lw + j, acc);
                            __syncwarp();
                            if (need && tag != want) tag = ld_word(a.Lw + j, acc);
                            waiting = __ballot_sync(kAllLanes, need && tag != want);
                        }
                        if (!need) acc = 0.0;
#pragma unroll
                        for (int s = 1; s < 32; s <<= 1) {  // affine maps composed nearest-first
                            const double up_acc = __shfl_up_sync(kAllLanes, acc, s), up_pm = __shfl_up_sync(kAllLanes, pm, s);
                            if (lane >= s) { acc = fma(up_pm, acc, up_acc); pm *= up_pm; }
                        }
                        cb = __shfl_sync(kAllLanes, acc, 31);
                        prod = __shfl_sync(kAllLanes, pm, 31);
                    }
                }
                for (int64_t j0 = tile + 33; j0 < a.n_tiles && prod != 0.0; j0 += 32) {  // very stiff meshes: further blocks of 32
                    const int64_t j = j0 + lane;
                    double pm = j < a.n_tiles ? __ldg(a.Mb + j) : 0.0;
                    double reach = pm;
#pragma unroll
                    for (int s = 1; s < 32; s <<= 1) {
                        const double up = __shfl_up_sync(kAllLanes, reach, s);
                        if (lane >= s) reach *= up;
                    }
                    double before = __shfl_up_sync(kAllLanes, reach, 1);
                    if (lane == 0) before = 1.0;
                    double acc = 0.0;  // (these tiles published long ago: a plain spin per lane; restructuring this loop like the
                    // block above made the whole kernel slower — 57 -> 82 us on a mesh that never enters it — and was dropped)
                    if (j < a.n_tiles && prod * before != 0.0)
                        while (ld_word(a.Lw + j, acc) != a.epoch * 4u + 1u) { }
#pragma unroll
                    for (int s = 1; s < 32; s <<= 1) {
                        const double up_acc = __shfl_up_sync(kAllLanes, acc, s), up_pm = __shfl_up_sync(kAllLanes, pm, s);
                        if (lane >= s) { acc = fma(up_pm, acc, up_acc); pm *= up_pm; }
                    }
                    cb = fma(prod, __shfl_sync(kAllLanes, acc, 31), cb);
                    prod *= __shfl_sync(kAllLanes, pm, 31);
                }
                if (lane == 0) s_carry[1] = cb;
                K3P_STAMP(5);
            } else if (warp == 1 && next_tile >= 0) {  // meanwhile: the NEXT tile's forward carry
                const double c = forward_fold(next_tile);
                if (lane == 0) s_cf[(it & 1) ^ 1] = c;
            }
            __syncthreads();
            double cw = s_carry[1];  // x behind my warp: the tile's carry folded through the warps on my right
            for (int w = W - 1; w > warp; --w) cw = fma(s_T[1][w], cw, s_E[1][w]);
            c = fma(sm.ex_b, cw, c);
            chunk_apply_bwd_plain<double, 8>(x, g, c);
        }
        // ---- p^{n+1} out ----
        if (cell0 + 8 <= nx && (reinterpret_cast<uintptr_t>(a.p_out + cell0) & 15) == 0) {
            double2 *q = reinterpret_cast<double2 *>(a.p_out + cell0);
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = make_double2(x[2 * j], x[2 * j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (cell0 + j < nx) a.p_out[cell0 + j] = x[j];
        }
        K3P_STAMP(6);
        // ---- the next step's forward aggregate of this tile (its right-hand side is p^{n+1} with that step's overrides) ----
        if (a.src_next || a.n_src == 0) {
            apply_overrides(x, cell0, nx, a.lbc, a.rbc, a.src_idx, a.n_src, a.src_next, tile_has_src);
            const double y = forward_ends(x);
            if (lane == 31) s_E[0][warp] = y;  // (this step's forward totals were last read before the previous barrier)
            __syncthreads();
            if (tid == 0) a.Ef_next[tile] = tile_forward_end(0.0);
        }
        __syncthreads();  // the shared scratch is reused by the next tile
        K3P_STAMP(7);
    }
}

constexpr int64_t kPassTileWords = 8;  // Mf, Mb, Ef[2], two 16-byte status words per tile

}  // namespace fbr

using namespace fbr;

extern "C" int64_t fbr_run_pass_work_bytes(int64_t nx) {
    const int64_t tiles = (nx + 256 * kPWMin - 1) / (256 * kPWMin);
    return (2 * nx + kPassTileWords * tiles + 16) * (int64_t)sizeof(double);
}

template <int W>
static int run_pass_impl(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                         int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                         const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                         double *snap, double *p_final, double *work, void *stream) {
    constexpr int kPTile = 256 * W;
    DeviceFacts f;
    if (int rc = device_facts(&f)) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t tiles = (nx + kPTile - 1) / kPTile;
    FBR_REQUIRE((tiles + 4096) * (n_steps + 2) < 0xffffffffll, "tiles x steps overflows the ticket counter: split the run");
    // ---- carve the workspace ----
    double *w = work;
    double *buf[2] = {w, w + nx};
    w += 2 * nx;
    double *Mf = w; w += tiles;
    double *Mb = w; w += tiles;
    double *Ef[2] = {w, w + tiles};
    w += 2 * tiles;
    if ((reinterpret_cast<uintptr_t>(w) & 15) != 0) ++w;  // 16-byte words
    double2 *Lw = reinterpret_cast<double2 *>(w); w += 2 * tiles;
    unsigned *ticket = reinterpret_cast<unsigned *>(w);
    FBR_CUDA(cudaMemsetAsync(Lw, 0, (size_t)tiles * 16 + 16, st));  // status words and the ticket counter
    const size_t smem = 4 * (size_t)kPTile * sizeof(double);  // staging of one tile's four rows
    FBR_CUDA(cudaFuncSetAttribute(pass_kernel<false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FBR_CUDA(cudaFuncSetAttribute(pass_kernel<true, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    FBR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pass_kernel<false, W>, 32 * W, smem));
    if (occ < 1) return fail(FBR_ERR_UNSUPPORTED, "pass kernel does not fit on an SM");
    int64_t grid = (int64_t)occ * f.sm_count;
    if (const char *e = getenv("FBR_PASS_GRID")) {  // testing aid: other interleavings of the look-back (never more CTAs than fit)
        const int64_t v = atoll(e);
        if (v >= 1 && v < grid) grid = v;
    }
    if (grid > tiles) grid = tiles;

    PassArgs a;
    a.fl = fl; a.fr = fr; a.fg = fg; a.nx = nx; a.n_tiles = tiles; a.lbc = lbc; a.rbc = rbc; a.src_idx = src_idx; a.n_src = n_src;
    a.Mf = Mf; a.Mb = Mb; a.Lw = Lw; a.ticket = ticket;
    // setup launch (epoch 0): tile products and the forward aggregates of step 0
    a.p_in = p0; a.p_out = nullptr; a.src_now = src_table; a.src_next = nullptr; a.Ef_now = nullptr; a.Ef_next = Ef[0]; a.epoch = 0;
    a.ticket_base = 0;
    if (n_steps > 0) {
        pass_kernel<true, W><<<(unsigned)grid, 32 * W, smem, st>>>(a);
        if (int rc = check_launch("fbr_run_pass (setup)")) return rc;
    }
    const double *cur = p0;
    int64_t snap_row = 0;
    for (int64_t n = 0; n < n_steps; ++n) {
        const bool row_due = snap && (n + 1) % snapshot_every == 0;
        const bool last = n + 1 == n_steps;
        double *out = row_due ? snap + snap_row * snap_row_stride : (last && p_final) ? p_final : buf[n & 1];
        if (out == cur) out = buf[(n & 1) ^ 1];  // (never the case: rows and work buffers alternate) keep in != out
        a.p_in = cur; a.p_out = out;
        a.src_now = n_src ? src_table + n * n_src : nullptr;
        a.src_next = (n_src && !last) ? src_table + (n + 1) * n_src : nullptr;
        a.Ef_now = Ef[n & 1]; a.Ef_next = Ef[(n & 1) ^ 1];
        a.epoch = (unsigned)(n + 1);
        a.ticket_base = (unsigned)((n + 1) * (tiles + grid));
        pass_kernel<false, W><<<(unsigned)grid, 32 * W, smem, st>>>(a);
        if (int rc = check_launch("fbr_run_pass")) return rc;
        if (row_due) ++snap_row;
        cur = out;
    }
    if (p_final && cur != p_final)
        FBR_CUDA(cudaMemcpyAsync(p_final, cur, (size_t)nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return FBR_OK;
}

extern "C" int fbr_run_pass_tiled_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                      int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                      const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                                      double *snap, double *p_final, int tile_warps, double *work, void *stream) {
    FBR_REQUIRE(nx >= 1 && n_steps >= 0, "bad nx / n_steps");
    FBR_REQUIRE(fl && fr && fg && p0 && work, "NULL factor / initial-state / work pointer");
    FBR_REQUIRE(lbc >= 0 && lbc <= 2 && rbc >= 0 && rbc <= 2, "bad boundary code");
    FBR_REQUIRE(n_src >= 0 && (n_src == 0 || (src_idx && src_table)), "sources given without src_idx / src_table");
    FBR_REQUIRE(!snap || (snapshot_every >= 1 && snap_row_stride >= nx), "bad snapshot layout");
    FBR_REQUIRE(n_steps < (1 << 29), "n_steps per call is limited to 2^29 - 1");
    FBR_REQUIRE(tile_warps == 0 || tile_warps == 2 || tile_warps == 4 || tile_warps == 8, "tile_warps must be 0 (default), 2, 4 or 8");
    if (const char *e = getenv("FBR_PASS_WARPS")) {  // tuning aid
        const int v = atoi(e);
        if (v == 2 || v == 4 || v == 8) tile_warps = v;
    }
    switch (tile_warps) {
        case 2: return run_pass_impl<2>(fl, fr, fg, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every,
                                        snap_row_stride, snap, p_final, work, stream);
        case 4: return run_pass_impl<4>(fl, fr, fg, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every,
                                        snap_row_stride, snap, p_final, work, stream);
        default: return run_pass_impl<8>(fl, fr, fg, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every,
                                         snap_row_stride, snap, p_final, work, stream);
    }
}

extern "C" int fbr_run_pass_f64(const double *fl, const double *fr, const double *fg, const double *p0, int64_t nx,
                                int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                                const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                                double *snap, double *p_final, double *work, void *stream) {
    return fbr_run_pass_tiled_f64(fl, fr, fg, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every, snap_row_stride,
                                  snap, p_final, 0, work, stream);
}

#ifdef FBR_K3P_PROF
extern "C" int fbr_debug_pass_prof(long long *host) {
    return (int)cudaMemcpyFromSymbol(host, fbr::g_k3p_prof, sizeof(fbr::g_k3p_prof));
}
#endif
