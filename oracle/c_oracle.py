"""ctypes loader for the C oracle (oracle/pds_oracle.c).  TEST INFRASTRUCTURE — see that file."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libpds_oracle.so")
_BC = {"None": 0, "Dirichlet": 1, "Neumann": 2}
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.fbo_run.restype = C.c_int64
        _lib.fbo_run_ensemble.restype = C.c_int64
        _lib.fbo_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return None if a is None else np.ascontiguousarray(np.atleast_1d(a), dtype=np.int64)


def max_threads():
    return int(load().fbo_max_threads())


def assemble(mesh, D, dt, lbc, rbc, src_idx, pml=0.0, sigma_max=2.0):
    mesh, D, s = _f(mesh), _f(D), _i(src_idx)
    nx = len(mesh)
    out = np.empty((3, nx))
    load().fbo_assemble(_p(mesh), _p(D), C.c_int64(nx), C.c_double(dt), _BC[lbc], _BC[rbc], _p(s),
                        C.c_int(len(s)), C.c_double(pml), C.c_double(sigma_max), _p(out[0]), _p(out[1]), _p(out[2]))
    return out[0], out[1], out[2]


def run(mesh, D, p0, n_steps, dt, lbc, rbc, src_idx, src_table, every=1, obs_idx=None, obs=None, obs_w=None):
    """Returns (snap (n_steps // every, nx), final (nx), misfit)."""
    mesh, D, p0, s, tab = _f(mesh), _f(D), _f(p0), _i(src_idx), _f(src_table)
    oi, ob, ow = _i(obs_idx), _f(obs), _f(obs_w)
    nx = len(mesh)
    snap = np.empty((n_steps // every if every else 0, nx))
    final = np.empty(nx)
    mis = C.c_double(0.0)
    info = load().fbo_run(_p(mesh), _p(D), _p(p0), C.c_int64(nx), C.c_int64(n_steps), C.c_double(dt), _BC[lbc],
                          _BC[rbc], _p(s), C.c_int(len(s)), _p(tab), C.c_int64(every), _p(snap), _p(final), _p(oi),
                          C.c_int(0 if oi is None else len(oi)), _p(ob), _p(ow), C.byref(mis))
    if info:
        raise np.linalg.LinAlgError(f"singular matrix (row {info - 1})")
    return snap, final, mis.value


def run_ensemble(mesh, D, zone_of_cell, p0, n_steps, dt, lbc, rbc, src_idx, src_table, obs_idx=None, obs=None,
                 obs_w=None, want_final=False, threads=0):
    """D: (M, nx) dense, or (M, Z) zone values with zone_of_cell (nx).  Returns (misfit (M), final | None)."""
    mesh, D, p0, s, tab = _f(mesh), _f(D), _f(p0), _i(src_idx), _f(src_table)
    oi, ob, ow = _i(obs_idx), _f(obs), _f(obs_w)
    zc = None if zone_of_cell is None else np.ascontiguousarray(zone_of_cell, dtype=np.int32)
    M, nx = D.shape[0], len(mesh)
    mis = np.zeros(M)
    final = np.empty((M, nx)) if want_final else None
    bad = load().fbo_run_ensemble(_p(mesh), _p(D), C.c_int(D.shape[1]), _p(zc), _p(p0), C.c_int64(M), C.c_int64(nx),
                                  C.c_int64(n_steps), C.c_double(dt), _BC[lbc], _BC[rbc], _p(s), C.c_int(len(s)),
                                  _p(tab), _p(oi), C.c_int(0 if oi is None else len(oi)), _p(ob), _p(ow), _p(mis),
                                  _p(final), C.c_int(threads))
    if bad:
        raise np.linalg.LinAlgError(f"{bad} singular members")
    return mis, final
