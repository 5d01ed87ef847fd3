// Shared helpers for the fiberis_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

#include <string>

#include "../../include/fiberis_b200.h"

namespace fbr {

std::string &last_error_ref();
int &launch_count_ref();

inline int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FBR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    launch_count_ref()++;
    return FBR_OK;
}

#define FBR_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return fbr::fail(FBR_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

#define FBR_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fbr::fail(FBR_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));          \
    } while (0)

struct DeviceFacts {
    int sm_count = 0, cc_major = 0, cc_minor = 0, device = -1;
    int64_t smem_optin = 0, l2_bytes = 0, hbm_bytes = 0;
};
int device_facts(DeviceFacts *out);

// K2 REG (fbr_solve_reg.cu): register-resident bare solve, 1 <= nx <= kSolveRegMaxNx
constexpr int kSolveRegMaxNx = 4096;
int step_factors_dispatch(const double *fl, const double *fr, const double *fg, const double *p, double *p_next, int64_t M, int64_t nx,
                          int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val, cudaStream_t stream);
int solve_reg_dispatch(const double *dl, const double *d, const double *du, const double *b, double *x, int64_t M, int64_t nx,
                       int32_t *singular, cudaStream_t stream);

int step_reg_dispatch(const double *dl, const double *d, const double *du, const double *p, double *p_next, int64_t M, int64_t nx,
                      int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val, int64_t src_val_stride,
                      int32_t *singular, cudaStream_t stream);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace fbr
