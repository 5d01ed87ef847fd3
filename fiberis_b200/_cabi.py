"""ctypes binding of the C-ABI library ``libfiberis_b200.so`` (declared in ``include/fiberis_b200.h``).

There is no CPU fallback: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C fiberis_b200/csrc``) every compute entry point
raises ``RuntimeError``.  PyTorch tensors are used only as device buffers; this module passes
their ``data_ptr()`` and the current CUDA stream to the library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libfiberis_b200.so")

FBR_OK, FBR_ERR_BAD_ARG, FBR_ERR_UNSUPPORTED, FBR_ERR_CUDA = 0, -1, -2, -3
BC_CODES = {"None": 0, "Dirichlet": 1, "Neumann": 2}
SOLVER_AUTO, SOLVER_THOMAS, SOLVER_PCR, SOLVER_REG, SOLVER_PIVOT = 0, 1, 2, 3, 4
WS_FACTORIZE, WS_SOLVE, WS_SOLVE_THOMAS, WS_RUN, WS_RUN_LONG, WS_RUN_STREAM, WS_RUN_WINDOW, WS_RUN_PASS = range(8)

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f64 = C.c_double

class RunOpts(C.Structure):
    """``fbr_run_opts`` of include/fiberis_b200.h (options of the fbr_run_* family; all zero = defaults)."""
    _fields_ = [("struct_bytes", C.c_int32), ("reserve_ctas", C.c_int32), ("snap_cell_stride", C.c_int64),
                ("scan_log2_eps", C.c_double), ("n_samples", C.c_int64), ("sample_row", C.c_void_p),
                ("sample_w", C.c_void_p), ("row_first_sample", C.c_void_p), ("base_sample", C.c_int32),
                ("first_sample", C.c_int32), ("last_sample", C.c_int32), ("snap_f64", C.c_int32),
                ("scan_mode", C.c_void_p), ("member_order", C.c_void_p)]


def run_opts(reserve_ctas=0, snap_cell_stride=0, scan_log2_eps=0.0, sampling=None, scan_mode=None, member_order=None, snap_f64=False):
    """Build an ``fbr_run_opts``.  ``sampling``: None, or a dict with device tensors ``sample_row`` (int32),
    ``sample_w`` (float64), ``row_first_sample`` (int32) and ints ``base_sample``, ``first_sample``, ``last_sample``."""
    o = RunOpts()
    o.struct_bytes = C.sizeof(RunOpts)
    o.reserve_ctas = int(reserve_ctas)
    o.snap_cell_stride = int(snap_cell_stride)
    o.scan_log2_eps = float(scan_log2_eps)
    if scan_mode is not None:
        o.scan_mode = scan_mode.data_ptr()
    o.snap_f64 = 1 if snap_f64 else 0
    if member_order is not None:  # device int32 permutation (a scheduling hint, see the header)
        o.member_order = member_order.data_ptr()
    if sampling is not None:
        o.n_samples = int(sampling["sample_row"].numel())
        o.sample_row = sampling["sample_row"].data_ptr()
        o.sample_w = sampling["sample_w"].data_ptr()
        o.row_first_sample = sampling["row_first_sample"].data_ptr()
        o.base_sample = int(sampling["base_sample"])
        o.first_sample = int(sampling["first_sample"])
        o.last_sample = int(sampling["last_sample"])
    return o


# name -> (restype, argtypes); mirrors include/fiberis_b200.h one to one
SIGNATURES = {
    "fbr_version": (_i32, []),
    "fbr_last_error": (C.c_char_p, []),
    "fbr_last_launch_count": (_i32, []),
    "fbr_nccl_allgather_f64": (_i32, [_p, _p, _i64, _p, _p]),
    "fbr_probe_fp64_peak": (_i32, [C.POINTER(_f64), _p, _p]),
    "fbr_transpose_f64": (_i32, [_p, _i64, _i64, _p, _p]),
    "fbr_device_info": (_i32, [C.POINTER(_i32)] * 3 + [C.POINTER(_i64)] * 3),
    "fbr_assemble_f64": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _f64, _i32, _i32, _p, _i32, _f64, _f64, _p, _p, _p, _p]),
    "fbr_assemble_zonal_f64": (_i32, [_p, _p, _i32, _p, _i64, _i64, _f64, _i32, _i32, _p, _i32, _f64, _f64, _p, _p, _p, _p]),
    "fbr_pml_sigma_f64": (_i32, [_p, _i64, _f64, _f64, _p, _p]),
    "fbr_factorize_work_bytes": (_i64, [_i64, _i64]),
    "fbr_factorize_f64": (_i32, [_p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "fbr_run_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _i32, _p, _p, _p,
                           C.POINTER(RunOpts), _p]),
    "fbr_run_f32": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _i32, _p, _p, _p,
                           C.POINTER(RunOpts), _p]),
    "fbr_run_zonal_f64": (_i32, [_p, _p, _i32, _p, _f64, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _p,
                                 _i64, _i64, _i64, _p, _p, _p, _i32, _p, _p, _p, C.POINTER(RunOpts), _p]),
    "fbr_run_zonal_f32": (_i32, [_p, _p, _i32, _p, _f64, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _p,
                                 _i64, _i64, _i64, _p, _p, _p, _i32, _p, _p, _p, C.POINTER(RunOpts), _p]),
    "fbr_rank_members_f64": (_i32, [_p, _i64, _p, _p]),
    "fbr_run_max_nx": (_i64, []),
    "fbr_run_long_max_nx": (_i64, []),
    "fbr_run_long_work_bytes": (_i64, [_i64]),
    "fbr_run_long_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "fbr_run_stream_work_bytes": (_i64, [_i64]),
    "fbr_run_stream_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _p, _p, _p, _p]),
    "fbr_run_pass_work_bytes": (_i64, [_i64]),
    "fbr_run_pass_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _p, _p, _p, _p]),
    "fbr_run_pass_tiled_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _p, _p, _i32, _p, _p]),
    "fbr_decay_horizon_f64": (_i32, [_p, _p, _i64, _f64, _i32, _p, _p, _p]),
    "fbr_run_window_plan": (_i32, [_i64, _i64, _i64, _i64, C.POINTER(_i64)]),
    "fbr_run_window_work_bytes": (_i64, [_i64]),
    "fbr_run_window_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _p, _p, _i64, _i64, _p, _p]),
    "fbr_run_window_plan_f32": (_i32, [_i64, _i64, _i64, _i64, C.POINTER(_i64)]),
    "fbr_run_window_f32_work_bytes": (_i64, [_i64]),
    "fbr_run_window_f32": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _i64, _p, _p, _i64, _i64, _p, _p]),
    "fbr_run_adaptive_max_nx": (_i64, []),
    "fbr_run_adaptive_f64": (_i32, [_p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _p, _p, _i32, _f64, _f64,
                                    _p, _f64, _f64, _f64, _f64, _f64, _f64, _i64, _i64, _p, _p, _p, _i64, _p, _p]),
    "fbr_solve_tridiag_work_bytes": (_i64, [_i64, _i64, _i32]),
    "fbr_solve_tridiag_f64": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p]),
    "fbr_solve_dense_f64": (_i32, [_p, _i64, _p, _p, _p]),
    "fbr_jacobi_dense_f64": (_i32, [_p, _p, _p, _i64, _f64, _i32, _p, _p, _p]),
    "fbr_step_f64": (_i32, [_p, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _i64, _p, _i32, _p, _p, _p]),
    "fbr_misfit_f64": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i32, _p, _p, _i64, _p, _p]),
    "fbr_workspace_bytes": (_i64, [_i32, _i64, _i64]),
    "fbr_rhs_f64": (_i32, [_p, _i64, _i64, _i32, _i32, _p, _i32, _p, _p, _p]),
    "fbr_error_norms_f64": (_i32, [_p, _p, _i64, _i64, _p, _p]),
    "fbr_jacobi_f64": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _f64, _i32, _p, _p]),
}

_lib = None


class FbrError(RuntimeError):
    pass


def load():
    """Load the library once; fail loudly when it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fiberis_b200: CUDA library not built ({LIB_PATH} missing). Build it with "
            "`make -C fiberis_b200/csrc` (needs nvcc, targets sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    """Map an fbr_status to the exception type the reference raises for the same condition."""
    if rc == FBR_OK:
        return
    msg = load().fbr_last_error().decode("utf-8", "replace")
    if rc == FBR_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == FBR_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise FbrError(msg)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device buffers must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def current_stream():
    """cudaStream_t of torch's current stream on the current device.  ``torch.cuda.current_stream()`` builds a Stream
    object through four Python layers (18 us a call, six calls per ensemble solve); the raw handle is one C call."""
    import torch
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_cuda_ok = False


def require_cuda():
    """Fail loudly without a CUDA device or without the library.  The positive answer is remembered: the check sits on
    every upload, and ``torch.cuda.is_available()`` costs ~5 us a call (16 calls = 9 % of C1's whole solve)."""
    global _cuda_ok
    if _cuda_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("fiberis_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    load()
    _cuda_ok = True


def launch_count() -> int:
    return int(load().fbr_last_launch_count())


def device_info() -> dict:
    lib = load()
    a, b, c = _i32(), _i32(), _i32()
    d, e, f = _i64(), _i64(), _i64()
    check(lib.fbr_device_info(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), C.byref(f)))
    return dict(sm_count=a.value, cc=(b.value, c.value), smem_optin=d.value, l2_bytes=e.value, hbm_bytes=f.value)


def as_f64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))
