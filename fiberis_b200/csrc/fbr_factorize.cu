// K1b  time-invariant factorisation  u_0 = d_0, l_i = dl_i r_{i-1}, u_i = d_i - l_i du_{i-1}, r_i = 1 / u_i,
//      g_i = du_i r_i.  One division per cell: fp64 division is a ~30-instruction sequence on the
//      sequential chain (three of them cost ~1000 cycles per cell on B200, one ~250).
//
// Two variants:
//   * sequential, one thread per model  -> ensembles of short systems (C3/C4: thousands of models);
//   * parallel in the cell direction    -> long meshes and single models.  The pivot recurrence
//       u_i = d_i - dl_i du_{i-1} / u_{i-1}
//     is a Moebius map of u_{i-1}; with u = p/q it is the linear map (p,q)_i = [[d_i, e_i],[1, 0]] (p,q)_{i-1},
//     e_i = -dl_i du_{i-1}, so pivots at arbitrary positions follow from a scan of 2x2 matrix products
//     (matrices are projective: each product is rescaled by a power of two, so nothing overflows).
//     The scan yields the pair (p_{i-1}, p_{i-2}) IN FRONT of every 8-cell chunk; inside its chunk a thread
//     runs the two-term recurrence and takes r_i = p_{i-1} / p_i with independent divisions
//     (fbr_factor_reg.cuh); the factors agree with the sequential variant to a few ulps.
#include "fbr_common.cuh"
#include "fbr_moebius.cuh"
#include "fbr_factor_reg.cuh"

#include <type_traits>

namespace fbr {

constexpr int kFacCPT = 8;

// ---- sequential variant ----
__global__ void __launch_bounds__(128) factorize_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                        const double *__restrict__ du, int64_t M, int64_t nx,
                                                        double *__restrict__ fl, double *__restrict__ fr,
                                                        double *__restrict__ fg, int32_t *__restrict__ singular) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t o = m * nx;
    double r_prev = 1.0, c_prev = 0.0;  // r = 1 / u: ONE division per cell (see the file header)
    int32_t bad = 0;
    for (int64_t i = 0; i < nx; ++i) {
        const double a = dl[o + i], b = d[o + i], c = du[o + i];
        const double l = (i == 0) ? 0.0 : a * r_prev;
        const double u = b - l * c_prev;
        if (!bad && (u == 0.0 || !isfinite(u))) bad = (int32_t)(i + 1);
        const double r = 1.0 / u;
        fl[o + i] = l;
        fr[o + i] = r;
        fg[o + i] = c * r;
        r_prev = r;
        c_prev = c;
    }
    if (singular) singular[m] = bad;
}

// ---- sequential recurrence, coalesced through shared-memory tiles ----
// Same arithmetic per model as factorize_kernel (bit-identical factors), but a warp owns 32 models
// and walks the mesh in tiles of 32 cells: a tile of every array is read with coalesced 256-byte
// requests (lane = cell) into shared memory, transposed by indexing (lane = model) for the
// recurrence, and written back the same way.  The plain kernel reads 32 different cache lines per
// request (0.5 - 1 TB/s on B200); this one streams (see profiles/).
constexpr int kTileCells = 32, kTileWarps = 4;

__global__ void __launch_bounds__(32 * kTileWarps) factorize_tiled_kernel(
    const double *__restrict__ dl, const double *__restrict__ d, const double *__restrict__ du, int64_t M, int64_t nx,
    double *__restrict__ fl, double *__restrict__ fr, double *__restrict__ fg, int32_t *__restrict__ singular) {
    extern __shared__ double s_tile[];  // [warp][3][cell][33]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double(*sa)[33] = reinterpret_cast<double(*)[33]>(s_tile + (size_t)warp * 3 * kTileCells * 33);
    double(*sb)[33] = sa + kTileCells;
    double(*sc)[33] = sb + kTileCells;
    const int64_t m0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    if (m0 >= M) return;
    const int rows = M - m0 < 32 ? (int)(M - m0) : 32;  // models of this group
    double r_prev = 1.0, c_prev = 0.0;
    int32_t bad = 0;
    for (int64_t c0 = 0; c0 < nx; c0 += kTileCells) {
        const int cells = nx - c0 < kTileCells ? (int)(nx - c0) : kTileCells;
        __syncwarp();
        for (int r0 = 0; r0 < rows; r0 += 8) {  // lane = cell; 24 independent loads in flight per thread
            double va[8], vb[8], vc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool ok = r0 + k < rows && lane < cells;
                const int64_t o = (m0 + (ok ? r0 + k : 0)) * nx + c0 + (ok ? lane : 0);
                va[k] = __ldg(dl + o); vb[k] = __ldg(d + o); vc[k] = __ldg(du + o);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (r0 + k < rows && lane < cells) { sa[lane][r0 + k] = va[k]; sb[lane][r0 + k] = vb[k]; sc[lane][r0 + k] = vc[k]; }
        }
        __syncwarp();
        if (lane < rows) {  // lane = model
            for (int c = 0; c < cells; ++c) {
                const double a = sa[c][lane], b = sb[c][lane], cc = sc[c][lane];
                const double l = (c0 + c == 0) ? 0.0 : a * r_prev;
                const double u = b - l * c_prev;
                if (!bad && (u == 0.0 || !isfinite(u))) bad = (int32_t)(c0 + c + 1);
                const double r = 1.0 / u;
                sa[c][lane] = l;
                sb[c][lane] = r;
                sc[c][lane] = cc * r;
                r_prev = r;
                c_prev = cc;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; ++r) {
            const int64_t o = (m0 + r) * nx + c0 + lane;
            if (lane < cells) { fl[o] = sa[lane][r]; fr[o] = sb[lane][r]; fg[o] = sc[lane][r]; }
        }
    }
    if (singular && lane < rows) singular[m0 + lane] = bad;
}

// ---- parallel variant ----
// Two passes over tiles of 256 * W cells (a thread owns 8 consecutive cells, fbr_factor_reg.cuh):
//   pass 1  every tile composes the 2x2 matrix that carries (p_{i-1}, p_{i-2}) across it,
//   scan    one CTA per model scans the tile matrices: the (p, q) pair in front of every tile,
//   pass 2  every tile factorises its rows from that pair and writes l, r, g with 16-byte stores.
// 72 B of traffic per cell (rows read twice, factors written once).
template <int W>
__device__ __forceinline__ void load_tile_rows(const double *__restrict__ dl, const double *__restrict__ d,
                                               const double *__restrict__ du, int64_t o, int64_t nx, int64_t gi, int nvalid,
                                               double (&l)[kFacCPT], double (&r)[kFacCPT], double (&g)[kFacCPT]) {
    const bool vec = nvalid == kFacCPT && ((o + gi) & 1) == 0 &&
                     ((reinterpret_cast<uintptr_t>(dl) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(du)) & 15) == 0;
    if (vec) {
        const double2 *ql = reinterpret_cast<const double2 *>(dl + o + gi), *qd = reinterpret_cast<const double2 *>(d + o + gi),
                      *qu = reinterpret_cast<const double2 *>(du + o + gi);
#pragma unroll
        for (int j = 0; j < kFacCPT / 2; ++j) {
            const double2 a = __ldg(ql + j), b = __ldg(qd + j), c = __ldg(qu + j);
            l[2 * j] = a.x; l[2 * j + 1] = a.y; r[2 * j] = b.x; r[2 * j + 1] = b.y; g[2 * j] = c.x; g[2 * j + 1] = c.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kFacCPT; ++j) {
            const bool in = j < nvalid;
            l[j] = in ? __ldg(dl + o + gi + j) : 0.0;
            r[j] = in ? __ldg(d + o + gi + j) : 1.0;
            g[j] = in ? __ldg(du + o + gi + j) : 0.0;
        }
    }
    if (gi == 0) l[0] = 0.0;  // outside the matrix
#pragma unroll
    for (int j = 0; j < kFacCPT; ++j)
        if (gi + j == nx - 1) g[j] = 0.0;
}

template <int W>
__global__ void __launch_bounds__(32 * W) factor_tile_total_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                                   const double *__restrict__ du, int64_t nx, int n_tiles,
                                                                   double *__restrict__ tile_mat) {
    __shared__ Mat2 s_w[W], s_total;
    __shared__ double s_c[W];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t o = (int64_t)blockIdx.y * nx, first = (int64_t)blockIdx.x * (256 * W), gi = first + (int64_t)tid * kFacCPT;
    const int64_t left = nx - gi;
    const int nvalid = left < 0 ? 0 : (left > kFacCPT ? kFacCPT : (int)left);
    double l[kFacCPT], r[kFacCPT], g[kFacCPT];
    load_tile_rows<W>(dl, d, du, o, nx, gi, nvalid, l, r, g);
    TileLink link;
    link.first_row = first;
    link.c_front = (tid == 0 && first > 0) ? __ldg(du + o + first - 1) : 0.0;
    link.total = &s_total;
    factor_rows_in_registers<kFacCPT, W>(l, r, g, tid * kFacCPT, nvalid, lane, warp, s_w, s_c, link);
    __syncthreads();
    if (tid < 4) {
        const double v = tid == 0 ? s_total.a : tid == 1 ? s_total.b : tid == 2 ? s_total.c : s_total.d;
        tile_mat[((int64_t)blockIdx.y * n_tiles + blockIdx.x) * 4 + tid] = v;
    }
}

// one CTA per model: (p, q) in front of every tile = (exclusive prefix of the tile matrices) applied to (1, 1)
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) factor_tile_scan_kernel(const double *__restrict__ tile_mat, int n_tiles,
                                                                        double *__restrict__ tile_vec) {
    __shared__ Mat2 s_warp[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *mats = tile_mat + (int64_t)blockIdx.x * n_tiles * 4;
    double *vecs = tile_vec + (int64_t)blockIdx.x * n_tiles * 2;
    const int per = (n_tiles + kScanThreads - 1) / kScanThreads, t0 = tid * per, t1 = min(n_tiles, t0 + per);
    Mat2 mine{1.0, 0.0, 0.0, 1.0};
    for (int t = t0; t < t1; ++t) mine = rescale_fast(compose_raw(Mat2{mats[4 * t], mats[4 * t + 1], mats[4 * t + 2], mats[4 * t + 3]}, mine));
    Mat2 inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const Mat2 up = shfl_up_mat(inc, s);
        if (lane >= s) inc = rescale_fast(compose_raw(inc, up));
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    Mat2 before{1.0, 0.0, 0.0, 1.0};
    for (int q = 0; q < warp; ++q) before = rescale_fast(compose_raw(s_warp[q], before));
    Mat2 ex = shfl_up_mat(inc, 1);
    if (lane == 0) ex = Mat2{1.0, 0.0, 0.0, 1.0};
    ex = compose_raw(ex, before);
    double p = ex.a + ex.b, q = ex.c + ex.d;
    for (int t = t0; t < t1; ++t) {
        const double s = pow2_scale_for(max(exp_field(p), exp_field(q)));
        p *= s; q *= s;
        vecs[2 * t] = p; vecs[2 * t + 1] = q;
        const double np = fma(mats[4 * t], p, mats[4 * t + 1] * q), nq = fma(mats[4 * t + 2], p, mats[4 * t + 3] * q);
        p = np; q = nq;
    }
}

template <int W>
__global__ void __launch_bounds__(32 * W) factor_tile_apply_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                                                                   const double *__restrict__ du, int64_t nx, int n_tiles,
                                                                   const double *__restrict__ tile_vec, double *__restrict__ fl,
                                                                   double *__restrict__ fr, double *__restrict__ fg,
                                                                   int32_t *__restrict__ singular) {
    __shared__ Mat2 s_w[W];
    __shared__ double s_c[W];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t o = (int64_t)blockIdx.y * nx, first = (int64_t)blockIdx.x * (256 * W), gi = first + (int64_t)tid * kFacCPT;
    const int64_t left = nx - gi;
    const int nvalid = left < 0 ? 0 : (left > kFacCPT ? kFacCPT : (int)left);
    double l[kFacCPT], r[kFacCPT], g[kFacCPT];
    load_tile_rows<W>(dl, d, du, o, nx, gi, nvalid, l, r, g);
    TileLink link;
    link.first_row = first;
    link.c_front = (tid == 0 && first > 0) ? __ldg(du + o + first - 1) : 0.0;
    const double *v = tile_vec + ((int64_t)blockIdx.y * n_tiles + blockIdx.x) * 2;
    link.p = __ldg(v);
    link.q = __ldg(v + 1);
    const int bad = factor_rows_in_registers<kFacCPT, W>(l, r, g, tid * kFacCPT, nvalid, lane, warp, s_w, s_c, link);
    const bool vec = nvalid == kFacCPT && ((o + gi) & 1) == 0 &&
                     ((reinterpret_cast<uintptr_t>(fl) | reinterpret_cast<uintptr_t>(fr) | reinterpret_cast<uintptr_t>(fg)) & 15) == 0;
    if (vec) {
        double2 *ql = reinterpret_cast<double2 *>(fl + o + gi), *qr = reinterpret_cast<double2 *>(fr + o + gi),
                *qg = reinterpret_cast<double2 *>(fg + o + gi);
#pragma unroll
        for (int j = 0; j < kFacCPT / 2; ++j) {
            ql[j] = make_double2(l[2 * j], l[2 * j + 1]);
            qr[j] = make_double2(r[2 * j], r[2 * j + 1]);
            qg[j] = make_double2(g[2 * j], g[2 * j + 1]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kFacCPT; ++j)
            if (j < nvalid) { fl[o + gi + j] = l[j]; fr[o + gi + j] = r[j]; fg[o + gi + j] = g[j]; }
    }
    if (bad && singular) atomicMin(singular + blockIdx.y, bad);  // the caller preset the flag to INT_MAX
}

__global__ void flag_preset_kernel(int32_t *flag, int64_t M, int32_t v) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m < M) flag[m] = v;
}
__global__ void flag_finish_kernel(int32_t *flag, int64_t M) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m < M && flag[m] == 0x7fffffff) flag[m] = 0;
}

}  // namespace fbr

using namespace fbr;

// warps per tile team of the parallel variant: the shortest team that covers the system, 16 (4096 cells) at most
static int fac_team_warps(int64_t nx) { return nx <= 512 ? 2 : nx <= 1024 ? 4 : nx <= 2048 ? 8 : 16; }

extern "C" int64_t fbr_factorize_work_bytes(int64_t M, int64_t nx) {
    const int W = fac_team_warps(nx);
    const int64_t tiles = (nx + 256 * W - 1) / (256 * W);
    return M * tiles * 6 * (int64_t)sizeof(double);
}

extern "C" int fbr_factorize_f64(const double *dl, const double *d, const double *du, int64_t M, int64_t nx,
                                 double *fl, double *fr, double *fg, int32_t *singular, double *work,
                                 void *stream) {
    FBR_REQUIRE(M >= 1 && nx >= 1, "M and nx must be >= 1");
    FBR_REQUIRE(dl && d && du && fl && fr && fg, "NULL array argument");
    cudaStream_t st = as_stream(stream);
    // thread-per-model is the better shape for many short systems; the cell-parallel variant needs
    // the caller's work buffer
    const bool parallel = work && nx >= 512 && M * 8 < nx;
    if (!parallel) {
        if (M >= 64 && !getenv("FBR_FACTORIZE_PLAIN")) {  // many models: stream through transposing tiles
            const int64_t groups = (M + 31) / 32;
            const int warps = groups >= 8 * 148 ? kTileWarps : 1;  // few groups: one warp per CTA spreads them over the SMs
            const size_t smem = (size_t)warps * 3 * kTileCells * 33 * sizeof(double);
            FBR_CUDA(cudaFuncSetAttribute(factorize_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)((size_t)kTileWarps * 3 * kTileCells * 33 * sizeof(double))));
            factorize_tiled_kernel<<<(unsigned)((groups + warps - 1) / warps), 32 * warps, smem, st>>>(
                dl, d, du, M, nx, fl, fr, fg, singular);
            return check_launch("fbr_factorize_f64 (tiled)");
        }
        const int block = M >= 128 * 148 ? 128 : 32;
        factorize_kernel<<<(unsigned)((M + block - 1) / block), block, 0, st>>>(dl, d, du, M, nx, fl, fr, fg, singular);
        return check_launch("fbr_factorize_f64");
    }
    const int W = fac_team_warps(nx);
    const int64_t tiles = (nx + 256 * W - 1) / (256 * W);
    FBR_REQUIRE(M <= 65535 && tiles <= 0x7fffffff, "too many models / tiles for the parallel factorisation");
    double *tile_mat = work, *tile_vec = work + M * tiles * 4;
    const dim3 grid((unsigned)tiles, (unsigned)M);
    if (singular) {
        flag_preset_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(singular, M, 0x7fffffff);
        if (int rc = check_launch("fbr_factorize_f64 (flag preset)")) return rc;
    }
    auto go = [&](auto wtag) -> int {
        constexpr int WW = decltype(wtag)::value;
        factor_tile_total_kernel<WW><<<grid, 32 * WW, 0, st>>>(dl, d, du, nx, (int)tiles, tile_mat);
        if (int rc = check_launch("fbr_factorize_f64 (tile totals)")) return rc;
        factor_tile_scan_kernel<<<(unsigned)M, kScanThreads, 0, st>>>(tile_mat, (int)tiles, tile_vec);
        if (int rc = check_launch("fbr_factorize_f64 (tile scan)")) return rc;
        factor_tile_apply_kernel<WW><<<grid, 32 * WW, 0, st>>>(dl, d, du, nx, (int)tiles, tile_vec, fl, fr, fg, singular);
        return check_launch("fbr_factorize_f64 (apply)");
    };
    int rc;
    switch (W) {
        case 2: rc = go(std::integral_constant<int, 2>{}); break;
        case 4: rc = go(std::integral_constant<int, 4>{}); break;
        case 8: rc = go(std::integral_constant<int, 8>{}); break;
        default: rc = go(std::integral_constant<int, 16>{}); break;
    }
    if (rc) return rc;
    if (singular) {
        flag_finish_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(singular, M);
        if (int rc = check_launch("fbr_factorize_f64 (flag finish)")) return rc;
    }
    return FBR_OK;
}
