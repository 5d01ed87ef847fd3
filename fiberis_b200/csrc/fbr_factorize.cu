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
//     The scan is only used to obtain the pivot IN FRONT of every 8-cell chunk; inside its chunk every
//     thread then runs the plain sequential recurrence, so all factors come out of the same
//     formulas as in the sequential variant (the chunk start differs from it by a few ulps at most).
#include "fbr_common.cuh"
#include "fbr_moebius.cuh"

namespace fbr {

constexpr int kFacCPT = 8;
constexpr int kFacW = 4;
constexpr int kFacTile = 32 * kFacW * kFacCPT;

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
// Composite of the thread's chunk, then inclusive scan over the tile.  Returns the EXCLUSIVE prefix of
// this thread within the tile (identity for thread 0); *tile_total receives the whole-tile composite
// (valid in every thread).
__device__ Mat2 tile_prefix(const double *__restrict__ dl, const double *__restrict__ d,
                            const double *__restrict__ du, int64_t nx, int64_t cell0, Mat2 *tile_total,
                            Mat2 (*s_w)[kFacW]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Mat2 g{1.0, 0.0, 0.0, 1.0};
#pragma unroll
    for (int j = 0; j < kFacCPT; ++j) {
        const int64_t i = cell0 + j;
        if (i < nx) {
            const double di = d[i];
            const double e = i > 0 ? -dl[i] * du[i - 1] : 0.0;
            // [[di, e],[1, 0]] * g
            const Mat2 n{di * g.a + e * g.c, di * g.b + e * g.d, g.a, g.b};
            g = (j == kFacCPT / 2 - 1) ? rescale(n) : n;  // keep extreme dt / dx^2 ratios in range
        }
    }
    g = rescale(g);
    Mat2 inc = g;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const Mat2 up = shfl_up_mat(inc, s);
        if (lane >= s) inc = compose(inc, up);
    }
    if (lane == 31) (*s_w)[warp] = inc;
    __syncthreads();
    Mat2 before{1.0, 0.0, 0.0, 1.0};  // composite of the warps before mine
    Mat2 total{1.0, 0.0, 0.0, 1.0};
    for (int v = 0; v < kFacW; ++v) {
        if (v == warp) before = total;
        total = compose((*s_w)[v], total);
    }
    *tile_total = total;
    Mat2 ex = shfl_up_mat(inc, 1);
    if (lane == 0) ex = Mat2{1.0, 0.0, 0.0, 1.0};
    return compose(ex, before);
}

__global__ void __launch_bounds__(32 * kFacW) factor_compose_kernel(const double *__restrict__ dl,
                                                                    const double *__restrict__ d,
                                                                    const double *__restrict__ du, int64_t nx,
                                                                    int n_tiles, double *__restrict__ tile_mat) {
    __shared__ Mat2 s_w[kFacW];
    const int64_t o = (int64_t)blockIdx.y * nx;
    const int64_t cell0 = ((int64_t)blockIdx.x * (32 * kFacW) + threadIdx.x) * kFacCPT;
    Mat2 total;
    tile_prefix(dl + o, d + o, du + o, nx, cell0, &total, &s_w);
    if (threadIdx.x == 0) {
        double *dst = tile_mat + ((int64_t)blockIdx.y * n_tiles + blockIdx.x) * 4;
        dst[0] = total.a; dst[1] = total.b; dst[2] = total.c; dst[3] = total.d;
    }
}

// one thread per model walks the tile composites: (p, q) in front of every tile
__global__ void factor_tile_walk_kernel(const double *__restrict__ tile_mat, int n_tiles, int64_t M,
                                        double *__restrict__ tile_vec) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    double p = 1.0, q = 1.0;
    for (int t = 0; t < n_tiles; ++t) {
        tile_vec[(m * n_tiles + t) * 2 + 0] = p;
        tile_vec[(m * n_tiles + t) * 2 + 1] = q;
        const double *g = tile_mat + (m * n_tiles + t) * 4;
        const double np = g[0] * p + g[1] * q, nq = g[2] * p + g[3] * q;
        const double big = fmax(fabs(np), fabs(nq));
        const int e = (big > 0.0 && isfinite(big)) ? -ilogb(big) : 0;
        p = scalbn(np, e);
        q = scalbn(nq, e);
    }
}

__global__ void __launch_bounds__(32 * kFacW) factor_apply_kernel(const double *__restrict__ dl,
                                                                  const double *__restrict__ d,
                                                                  const double *__restrict__ du, int64_t nx,
                                                                  int n_tiles, const double *__restrict__ tile_vec,
                                                                  double *__restrict__ fl, double *__restrict__ fr,
                                                                  double *__restrict__ fg,
                                                                  int32_t *__restrict__ singular) {
    __shared__ Mat2 s_w[kFacW];
    const int64_t o = (int64_t)blockIdx.y * nx;
    const int64_t cell0 = ((int64_t)blockIdx.x * (32 * kFacW) + threadIdx.x) * kFacCPT;
    Mat2 total;
    const Mat2 ex = tile_prefix(dl + o, d + o, du + o, nx, cell0, &total, &s_w);
    if (cell0 >= nx) return;
    const double *v = tile_vec + ((int64_t)blockIdx.y * n_tiles + blockIdx.x) * 2;
    const double p = ex.a * v[0] + ex.b * v[1], q = ex.c * v[0] + ex.d * v[1];
    double r_prev = q / p;  // reciprocal of the pivot in front of my chunk (unused for the very first cell)
    double c_prev = cell0 > 0 ? du[o + cell0 - 1] : 0.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < kFacCPT; ++j) {
        const int64_t i = cell0 + j;
        if (i < nx) {
            const double a = dl[o + i], b = d[o + i], c = du[o + i];
            const double l = (i == 0) ? 0.0 : a * r_prev;
            const double u = b - l * c_prev;
            if (!bad && (u == 0.0 || !isfinite(u))) bad = (int)(i + 1);
            const double r = 1.0 / u;
            fl[o + i] = l;
            fr[o + i] = r;
            fg[o + i] = c * r;
            r_prev = r;
            c_prev = c;
        }
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

extern "C" int64_t fbr_factorize_work_bytes(int64_t M, int64_t nx) {
    const int64_t tiles = (nx + kFacTile - 1) / kFacTile;
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
    const int64_t tiles = (nx + kFacTile - 1) / kFacTile;
    FBR_REQUIRE(M <= 65535 && tiles <= 0x7fffffff, "too many models / tiles for the parallel factorisation");
    double *tile_mat = work, *tile_vec = work + M * tiles * 4;
    const dim3 grid((unsigned)tiles, (unsigned)M);
    if (singular) {
        flag_preset_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(singular, M, 0x7fffffff);
        if (int rc = check_launch("fbr_factorize_f64 (flag preset)")) return rc;
    }
    factor_compose_kernel<<<grid, 32 * kFacW, 0, st>>>(dl, d, du, nx, (int)tiles, tile_mat);
    if (int rc = check_launch("fbr_factorize_f64 (compose)")) return rc;
    factor_tile_walk_kernel<<<(unsigned)((M + 31) / 32), 32, 0, st>>>(tile_mat, (int)tiles, M, tile_vec);
    if (int rc = check_launch("fbr_factorize_f64 (tile walk)")) return rc;
    factor_apply_kernel<<<grid, 32 * kFacW, 0, st>>>(dl, d, du, nx, (int)tiles, tile_vec, fl, fr, fg, singular);
    if (int rc = check_launch("fbr_factorize_f64 (apply)")) return rc;
    if (singular) {
        flag_finish_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(singular, M);
        if (int rc = check_launch("fbr_factorize_f64 (flag finish)")) return rc;
    }
    return FBR_OK;
}
