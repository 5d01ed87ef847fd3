// Register-resident sweep building blocks shared by the run kernels (K3 / K3W), fp64, 8 cells per thread.
//
// A thread owns 8 consecutive cells: multipliers m[0..7] (= -l going forward, -g going backward) and
// values v[0..7].  One sweep of the first-order recurrence  y_i = v_i + m_i y_{i-1}  is
//   (1) chunk_end:   the chunk's last value for a zero carry-in — two independent 4-cell chains
//                    joined by one FMA (depth 5 instead of 8),
//   (2) a Kogge-Stone scan of the chunk-end values across the warp, NS stages (5 = exact; fewer
//       when the multipliers' decay horizon makes the far stages vanish below 2^-80),
//   (3) the carry from the neighbouring warp(s) through shared memory,
//   (4) chunk_apply: the second local pass with the true carry-in, again as two parallel half
//       chains (the second half starts from  y_3 = half_end_a + P_a * carry).
// The products P_a (first half) and P_b (second half) are time-invariant and live in registers.
// Shared-memory traffic uses precomputed 32-bit addresses (ld/st.shared) so that the time loop
// holds no address arithmetic, and lanes that must not publish write to a dummy slot instead of
// being predicated off.  (Point-to-point flagged messages between neighbouring warps were measured
// against the CTA barrier on B200: 0.59 vs 0.48 us per step for 16 warps — the barrier stays.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fbr {

constexpr unsigned kAllLanes = 0xffffffffu;

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
// bulk copy global -> shared (1-D TMA) whose bytes complete on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void *gsrc, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// v with all bits cleared when keep == 0 (keep is 0 or -1): a select without a predicate
__device__ __forceinline__ double and_mask(double v, int keep) {
    const long long b = __double_as_longlong(v) & (long long)keep;
    return __longlong_as_double(b);
}

// Forward direction: cells 0 -> 7.  m = -l.
struct HalfProducts {
    double a, b;  // product of the multipliers of the first / second half of the chunk
};

__device__ __forceinline__ HalfProducts half_products_fwd(const double (&l)[8]) {
    HalfProducts p;
    p.a = (l[0] * l[1]) * (l[2] * l[3]);  // four sign flips cancel
    p.b = (l[4] * l[5]) * (l[6] * l[7]);
    return p;
}
__device__ __forceinline__ HalfProducts half_products_bwd(const double (&g)[8]) {
    HalfProducts p;
    p.a = (g[7] * g[6]) * (g[5] * g[4]);  // backward: the "first" half is cells 7..4
    p.b = (g[3] * g[2]) * (g[1] * g[0]);
    return p;
}

// (1) forward: returns the zero-carry chunk-end value; ya = zero-carry value at cell 3
__device__ __forceinline__ double chunk_end_fwd(const double (&x)[8], const double (&l)[8], double pb, double &ya) {
    double a = x[0], b = x[4];
    a = fma(-l[1], a, x[1]);  b = fma(-l[5], b, x[5]);
    a = fma(-l[2], a, x[2]);  b = fma(-l[6], b, x[6]);
    a = fma(-l[3], a, x[3]);  b = fma(-l[7], b, x[7]);
    ya = a;
    return fma(pb, a, b);
}
// (4) forward: x_j <- r_j * y_j with the true carry-in c (y at the cell before the chunk)
__device__ __forceinline__ void chunk_apply_fwd(double (&x)[8], const double (&l)[8], const double (&r)[8], double pa,
                                                double ya, double c) {
    double b = fma(pa, c, ya);     // true y at cell 3
    double a = fma(-l[0], c, x[0]);
    double t;
    t = fma(-l[4], b, x[4]); b = t; x[4] = r[4] * t;
    x[0] = r[0] * a;
    t = fma(-l[1], a, x[1]); a = t; x[1] = r[1] * t;
    t = fma(-l[5], b, x[5]); b = t; x[5] = r[5] * t;
    t = fma(-l[2], a, x[2]); a = t; x[2] = r[2] * t;
    t = fma(-l[6], b, x[6]); b = t; x[6] = r[6] * t;
    t = fma(-l[3], a, x[3]); x[3] = r[3] * t;
    t = fma(-l[7], b, x[7]); x[7] = r[7] * t;
}
// (1) backward (cells 7 -> 0): zero-carry value at cell 0; za = zero-carry value at cell 4
__device__ __forceinline__ double chunk_end_bwd(const double (&x)[8], const double (&g)[8], double pb, double &za) {
    double a = x[7], b = x[3];
    a = fma(-g[6], a, x[6]);  b = fma(-g[2], b, x[2]);
    a = fma(-g[5], a, x[5]);  b = fma(-g[1], b, x[1]);
    a = fma(-g[4], a, x[4]);  b = fma(-g[0], b, x[0]);
    za = a;
    return fma(pb, a, b);
}
// (4) backward: x_j <- z_j with the true carry-in c (z at the cell after the chunk)
__device__ __forceinline__ void chunk_apply_bwd(double (&x)[8], const double (&g)[8], double pa, double za, double c) {
    double b = fma(pa, c, za);  // true z at cell 4
    x[7] = fma(-g[7], c, x[7]);
    x[3] = fma(-g[3], b, x[3]);
    x[6] = fma(-g[6], x[7], x[6]);
    x[2] = fma(-g[2], x[3], x[2]);
    x[5] = fma(-g[5], x[6], x[5]);
    x[1] = fma(-g[1], x[2], x[1]);
    x[4] = fma(-g[4], x[5], x[4]);
    x[0] = fma(-g[0], x[1], x[0]);
}

// Plain (single-chain) versions of the local passes: 16 instead of 22 operations per 8-cell sweep,
// CPT-deep dependent chains.  Preferred where the SM is shared by many warps (ensembles: K3), where
// throughput counts more than the depth of one warp's chain.  T = double (8 cells per thread) or
// float (16 cells per thread: the fp32 mode has the same register footprint and half the shuffles).
template <typename T> __device__ __forceinline__ T fma_any(T a, T b, T c);
template <> __device__ __forceinline__ double fma_any<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_any<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T, int CPT>
__device__ __forceinline__ T chunk_end_fwd_plain(const T (&x)[CPT], const T (&l)[CPT]) {
    T y = x[0];
#pragma unroll
    for (int j = 1; j < CPT; ++j) y = fma_any<T>(-l[j], y, x[j]);
    return y;
}
template <typename T, int CPT>
__device__ __forceinline__ void chunk_apply_fwd_plain(T (&x)[CPT], const T (&l)[CPT], const T (&r)[CPT], T c) {
    T y = c;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        y = fma_any<T>(-l[j], y, x[j]);
        x[j] = r[j] * y;
    }
}
template <typename T, int CPT>
__device__ __forceinline__ T chunk_end_bwd_plain(const T (&x)[CPT], const T (&g)[CPT]) {
    T z = x[CPT - 1];
#pragma unroll
    for (int j = CPT - 2; j >= 0; --j) z = fma_any<T>(-g[j], z, x[j]);
    return z;
}
template <typename T, int CPT>
__device__ __forceinline__ void chunk_apply_bwd_plain(T (&x)[CPT], const T (&g)[CPT], T c) {
    T z = c;
#pragma unroll
    for (int j = CPT - 1; j >= 0; --j) {
        z = fma_any<T>(-g[j], z, x[j]);
        x[j] = z;
    }
}
// product of an even number of multipliers (the sign flips of -m cancel)
template <typename T, int CPT>
__device__ __forceinline__ T chunk_product(const T (&m)[CPT]) {
    static_assert(CPT % 2 == 0, "an even chunk keeps the product's sign");
    T p = m[0];
#pragma unroll
    for (int j = 1; j < CPT; ++j) p *= m[j];
    return p;
}
__device__ __forceinline__ void sts_any(unsigned addr, double v) { sts_f64(addr, v); }
__device__ __forceinline__ void sts_any(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float and_mask(float v, int keep) { return __int_as_float(__float_as_int(v) & keep); }

// Time-invariant stage multipliers of the warp scan: stage s adds P[s] * (value of lane -/+ 2^s);
// lanes a stage does not reach get a zero multiplier, so the scan needs no predicates.
// tot_f / tot_b = products over lanes 0..lane / lane..31; ex_f / ex_b = the exclusive versions.
template <typename T>
struct ScanMultipliersT {
    T f[5], b[5], ex_f, ex_b, tot_f, tot_b;
};
using ScanMultipliers = ScanMultipliersT<double>;
template <typename T>
__device__ __forceinline__ ScanMultipliersT<T> scan_multipliers(T chunk_f, T chunk_b, int lane) {
    ScanMultipliersT<T> m;
    T pf = chunk_f, pb = chunk_b;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const bool fa = lane >= (1 << s), ba = lane + (1 << s) < 32;
        m.f[s] = fa ? pf : (T)0;
        m.b[s] = ba ? pb : (T)0;
        const T upf = __shfl_up_sync(kAllLanes, pf, 1 << s);
        const T dnb = __shfl_down_sync(kAllLanes, pb, 1 << s);
        if (fa) pf *= upf;
        if (ba) pb *= dnb;
    }
    m.ex_f = __shfl_up_sync(kAllLanes, pf, 1);
    m.ex_b = __shfl_down_sync(kAllLanes, pb, 1);
    if (lane == 0) m.ex_f = (T)1;
    if (lane == 31) m.ex_b = (T)1;
    m.tot_f = pf;
    m.tot_b = pb;
    return m;
}

}  // namespace fbr
