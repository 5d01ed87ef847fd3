// 2x2 projective matrices for the pivot recurrence of a tridiagonal LU (shared by K1b and the adaptive loop).
//   u_i = d_i - dl_i du_{i-1} / u_{i-1}  is a Moebius map of u_{i-1}; with u = p / q it is the linear map
//   (p, q)_i = [[d_i, e_i], [1, 0]] (p, q)_{i-1},  e_i = -dl_i du_{i-1},
// so pivots at arbitrary positions follow from a scan of matrix products.  Matrices are projective: every
// product is rescaled by a power of two (exact), so nothing overflows.
#pragma once
#include <cuda_runtime.h>

namespace fbr {

constexpr unsigned kAll = 0xffffffffu;

struct Mat2 {  // [[a, b], [c, d]]
    double a, b, c, d;
};

__device__ __forceinline__ Mat2 rescale(Mat2 m) {
    const double big = fmax(fmax(fabs(m.a), fabs(m.b)), fmax(fabs(m.c), fabs(m.d)));
    if (big > 0.0 && isfinite(big)) {
        const int e = -ilogb(big);
        m.a = scalbn(m.a, e); m.b = scalbn(m.b, e); m.c = scalbn(m.c, e); m.d = scalbn(m.d, e);
    }
    return m;
}
// later * earlier (apply `earlier` first)
__device__ __forceinline__ Mat2 compose(const Mat2 &later, const Mat2 &earlier) {
    Mat2 r;
    r.a = later.a * earlier.a + later.b * earlier.c;
    r.b = later.a * earlier.b + later.b * earlier.d;
    r.c = later.c * earlier.a + later.d * earlier.c;
    r.d = later.c * earlier.b + later.d * earlier.d;
    return rescale(r);
}
__device__ __forceinline__ Mat2 shfl_up_mat(const Mat2 &m, int delta) {
    Mat2 r;
    r.a = __shfl_up_sync(kAll, m.a, delta); r.b = __shfl_up_sync(kAll, m.b, delta);
    r.c = __shfl_up_sync(kAll, m.c, delta); r.d = __shfl_up_sync(kAll, m.d, delta);
    return r;
}


}  // namespace fbr
