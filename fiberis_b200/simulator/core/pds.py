"""1D pressure-diffusion models: the reference's ``PDS1D_SingleSource`` / ``PDS1D_MultiSource`` API
(/root/reference/src/fiberis/simulator/core/pds.py:17-610) on a B200-native engine.

What stays on the host (cheap, sequential, and it has to be bit-identical):
  * validation and ``history`` bookkeeping;
  * the time axis — ``while taxis[-1] < t_total: taxis.append(taxis[-1] + dt)`` with fp64
    accumulation (reference :287,309), so the number of steps and every ``taxis`` entry are equal
    bit for bit;
  * sampling the sources at the start of every step (``np.interp``, reference core1D.py:247).

What runs on the GPU through the C ABI (``include/fiberis_b200.h``):
  * K1 fused coefficient assembly + K1b one-off factorisation (A is time-invariant for fixed dt);
  * K3 the whole time loop in ONE persistent kernel (pressure and factors resident on chip,
    snapshots streamed out as they are produced);
  * for ``optimizer=True``: per iteration assembly, three batched tridiagonal solves (K2) and the
    step-doubling error reduction (K5); only the 8-byte error crosses to the host, where the
    reference's scalar accept/reject logic runs unchanged.

Additive keyword arguments of ``solve`` (defaults keep the reference behaviour):
  ``dtype='float64'|'float32'``, ``snapshot_every=1``, ``pml_thickness=0.0``, ``sigma_max``.
"""
from __future__ import annotations

from datetime import datetime

import numpy as np

from fiberis_b200 import _cabi, _engine
from fiberis_b200.simulator.optimizer import tso

_ALLOWED_BC = ["Dirichlet", "Neumann", "None"]


def _in_mesh(idx, n) -> bool:
    """``idx in np.arange(n)`` (the reference's membership test, pds.py:179-181, 478-490) without
    materialising and scanning an n-element array per source: on a 1e7-cell mesh with 16 sources the
    literal form costs 130 ms per solve.  Same truth value for every scalar (integral floats count,
    NaN / negative / fractional do not); anything else falls back to the literal expression."""
    if np.ndim(idx) == 0:
        try:
            return bool(0 <= idx < n and idx == int(idx))
        except (TypeError, ValueError, OverflowError):
            pass
    return idx in np.arange(n)


def _uses_stock_interp(src) -> bool:
    """True when ``src`` samples with plain np.interp on (taxis, data), so a whole step table can
    be produced by one vectorised np.interp call with identical results."""
    from fiberis_b200.analyzer.Data1D.core1D import Data1D
    fn = getattr(type(src), "get_value_by_time", None)
    if fn is Data1D.get_value_by_time:
        return src.data is not None and src.taxis is not None and np.size(src.data) > 0
    # the reference's own Data1D (same np.interp semantics) when scripts mix the two packages
    mod = getattr(type(src), "__module__", "")
    return mod.startswith("fiberis.analyzer.Data1D") and isinstance(getattr(src, "taxis", None), np.ndarray) \
        and np.size(src.taxis) > 1



def fixed_time_axis_array(t0, dt, t_total):
    """``fixed_time_axis`` as a float64 array, without the detour through a Python list (5000 steps: 0.02 instead of
    0.1 ms, and ``np.interp`` of the source table no longer converts a list) — the same sequential ``np.cumsum``
    accumulation, bit for bit (tests/test_host_logic.py)."""
    try:
        f0, fdt, ft = float(t0), float(dt), float(t_total)
        plain = isinstance(t0, (int, float)) and isinstance(dt, (int, float)) and isinstance(t_total, (int, float))
    except (TypeError, ValueError):
        plain = False
    if not plain or not (fdt > 0.0) or not np.isfinite(ft - f0) or (ft - f0) / fdt < 64:
        return np.array(fixed_time_axis(t0, dt, t_total), dtype=np.float64)
    parts, last = [np.array([f0])], f0
    while last < ft:
        guess = int(min(max((ft - last) / fdt, 1.0), 1e6)) + 2
        block = np.cumsum(np.concatenate(([last], np.full(guess, fdt))))[1:]  # ((t + dt) + dt) + ... in order
        stop = int(np.searchsorted(block, ft, side="left"))  # first entry >= t_total (the axis increases)
        block = block[: stop + 1]
        parts.append(block)
        last = float(block[-1])
    return np.concatenate(parts)


def fixed_time_axis_pair(t0, dt, t_total):
    """``(list, array)``: the reference's time axis as the list its loop builds (what log lines print, ``t0`` kept as
    given) and, for long plain axes, the same values as a float64 array built without the list detour (``None`` for
    short or unusual axes, where the list is literally the reference's loop and ``np.array(list)`` keeps its dtype)."""
    try:
        f0, fdt, ft = float(t0), float(dt), float(t_total)
        plain = isinstance(t0, (int, float)) and isinstance(dt, (int, float)) and isinstance(t_total, (int, float))
    except (TypeError, ValueError):
        plain = False
    if not plain or not (fdt > 0.0) or not np.isfinite(ft - f0) or (ft - f0) / fdt < 64:
        return fixed_time_axis(t0, dt, t_total), None
    arr = fixed_time_axis_array(t0, dt, t_total)
    lst = arr.tolist()
    lst[0] = t0
    return lst, arr


def fixed_time_axis(t0, dt, t_total):
    """The time axis the reference's loop builds (``while taxis[-1] < t_total: taxis.append(taxis[-1] + dt)``,
    src/fiberis/simulator/core/pds.py:287,309), as a list.  The fp64 accumulation decides the step count (dt = 0.1,
    t_total = 1 gives 11 steps), so it is reproduced bit for bit: ``np.cumsum`` adds sequentially in the same order;
    long axes are built in chunks instead of one interpreter iteration per step."""
    taxis = [t0]
    try:
        f0, fdt, ft = float(t0), float(dt), float(t_total)
        plain = isinstance(t0, (int, float)) and isinstance(dt, (int, float)) and isinstance(t_total, (int, float))
    except (TypeError, ValueError):
        plain = False
    if not plain or not (fdt > 0.0) or not np.isfinite(ft - f0) or (ft - f0) / fdt < 64:
        while taxis[-1] < t_total:  # short or unusual axes: literally the reference's loop
            taxis.append(taxis[-1] + dt)
        return taxis
    taxis = [f0]
    while taxis[-1] < ft:
        guess = int(min(max((ft - taxis[-1]) / fdt, 1.0), 1e6)) + 2
        block = np.cumsum(np.concatenate(([taxis[-1]], np.full(guess, fdt))))[1:]  # ((t + dt) + dt) + ... in order
        stop = np.nonzero(block >= ft)[0]
        taxis.extend(block[: stop[0] + 1].tolist() if len(stop) else block.tolist())
    taxis[0] = t0
    return taxis

_K3_MAX_SOURCES = 253  # fbr_run_*: a source column is one descriptor byte


class PDS1D_SingleSource:
    """One mesh, one pressure source pinned at ``sourceidx`` (a Dirichlet-in-time row)."""

    _multi = False

    def __init__(self):
        self.mesh = None
        self.source = None
        self.lbc = None
        self.rbc = None
        self.initial = None
        self.diffusivity = None
        self.t0 = 0
        self.taxis = None
        self.sourceidx = None
        self.snapshot = None
        self.history = []

    # ------------------------------------------------------------------ configuration
    def _log_check(self, result):
        self.history.append(result[1])

    def set_mesh(self, mesh):
        self.mesh = mesh
        self._log_check(self._check_mesh())

    def set_source(self, source):
        self.source = source
        self._log_check(self._check_source())

    def set_bcs(self, lbc="Neumann", rbc="Neumann"):
        self.lbc, self.rbc = lbc, rbc
        self._log_check(self._check_bc())

    def set_initial(self, initial):
        self.initial = initial  # kept by reference, like the original
        self._log_check(self._check_initial())

    def set_diffusivity(self, diffusivity):
        """Scalar (broadcast over the mesh, which must be set first) or a per-node array."""
        if isinstance(diffusivity, (int, float)):
            self.diffusivity = np.full(len(self.mesh), diffusivity)
            print("Diffusivity is a single scalar value, broadcasted to the mesh length.")
        elif isinstance(diffusivity, (list, np.ndarray)) and len(diffusivity) == len(self.mesh):
            self.diffusivity = np.array(diffusivity)
        else:
            raise ValueError(
                "Diffusivity must be either a single scalar value or an array of the same length as the mesh.")
        self._log_check(self._check_diffusivity())

    def set_t0(self, t0):
        self.t0 = t0
        self.history.append(f"Initial time set done. \nThe simulation starts at {t0}.")

    def set_sourceidx(self, sourceidx):
        self.sourceidx = sourceidx
        self._log_check(self._check_sourceidx())

    def print(self):
        bar = "=" * 41
        print(bar)
        print(f"{type(self).__name__ + ' Info':^41}")
        print(bar)
        for key, value in self.__dict__.items():
            if isinstance(value, np.ndarray):
                print(f"{key:>15}: shape={value.shape}, unique={np.unique(value)}")
            else:
                print(f"{key:>15}: {value}")
        print(bar)

    # ------------------------------------------------------------------ checks
    def _check_mesh(self):
        if self.mesh is None:
            return False, "Mesh is not set."
        if not isinstance(self.mesh, np.ndarray):
            return False, "Mesh must be a numpy array."
        if len(self.mesh) < 2:
            return False, "Mesh must have at least 2 points."
        return True, "Mesh is properly set."

    def _check_source(self):
        if self.source is None:
            return False, "Source term is not set."
        return True, "Source term is properly set."

    def _check_bc(self):
        if self.lbc is None or self.rbc is None:
            return False, "Boundary condition(s) (lbc/rbc) not set."
        if self.lbc not in _ALLOWED_BC:
            return False, f"Invalid left BC: {self.lbc}. Must be {_ALLOWED_BC}."
        if self.rbc not in _ALLOWED_BC:
            return False, f"Invalid right BC: {self.rbc}. Must be {_ALLOWED_BC}."
        return True, "Boundary conditions are properly set."

    def _check_initial(self):
        if self.initial is None:
            return False, "Initial condition is not set."
        if self.mesh is None:
            return False, "Mesh is not set. Please set the mesh first."
        if len(self.initial) != len(self.mesh):
            return False, "Length of initial condition array must match mesh length."
        return True, "Initial condition is properly set."

    def _check_diffusivity(self):
        if self.diffusivity is None:
            return False, "Diffusivity is not set."
        if not isinstance(self.diffusivity, np.ndarray):
            return False, "Diffusivity must be a numpy array or the pds frame is failed to init the diffusivity."
        if len(self.diffusivity) != len(self.mesh):
            return False, "Diffusivity array length must match mesh length."
        return True, "Diffusivity is properly set."

    def _check_sourceidx(self):
        if self.sourceidx is None:
            return False, "Source index is not set."
        if not _in_mesh(self.sourceidx, len(self.mesh)):
            return False, "Source index is not in the mesh."
        return True, "Source index is properly set."

    def self_check(self, params=None):
        """True when every requested check passes; failures are printed (never raised)."""
        checks = {"mesh": self._check_mesh, "source": self._check_source, "bc": self._check_bc,
                  "initial": self._check_initial, "diffusivity": self._check_diffusivity,
                  "sourceidx": self._check_sourceidx}
        if params is None:
            params = list(checks)
        elif isinstance(params, str):
            params = [params]
        ok_all = True
        for name in params:
            if name not in checks:
                print(f"Parameter '{name}' is not recognized. Skipping.")
                ok_all = False
                continue
            ok, msg = checks[name]()
            if not ok:
                print(msg)
                ok_all = False
        return ok_all

    # ------------------------------------------------------------------ source helpers
    def _sources(self):
        return list(self.source) if self._multi else [self.source]

    def _source_indices(self):
        return np.atleast_1d(np.asarray(self.sourceidx)).astype(np.int64)

    def _source_values_at(self, t_rel):
        return [s.get_value_by_time(t_rel) for s in self._sources()]

    def _source_table(self, step_starts):
        """(n_steps, n_src): value of every source at the start of every step."""
        rel = np.asarray(step_starts, dtype=np.float64) - self.t0
        cols = []
        for s in self._sources():
            if _uses_stock_interp(s):
                cols.append(np.interp(rel, s.taxis, s.data))
            else:
                cols.append(np.array([s.get_value_by_time(float(t)) for t in rel], dtype=np.float64))
        return np.ascontiguousarray(np.stack(cols, axis=1)) if cols else np.zeros((len(rel), 0))

    def _end_time(self, kwargs):
        if "t_total" in kwargs:
            print("Time array generated using t_total.")
            return kwargs["t_total"], "Time array generated using t_total."
        first = self.source[0] if self._multi else self.source
        print("Time array generated using the source term.")
        return first.taxis[-1] + self.t0, "Time array generated using the source term."

    # ------------------------------------------------------------------ solve
    def solve(self, optimizer=False, **kwargs):
        """Integrate to ``t_total`` (default: end of the source record).

        ``optimizer=False``: fixed ``dt`` (required).  ``optimizer=True``: adaptive step doubling
        starting at ``dt_init`` with the ``tso`` keyword arguments.  Returns ``None``; results are
        left in ``snapshot`` ((n_t, nx) ndarray, row 0 = initial) and ``taxis``.
        """
        if not self.self_check():
            print("Self check failed. Please check the parameters.")
            return None
        if optimizer:
            dt_init = kwargs["dt_init"]
            t_total, _ = self._end_time(kwargs)
            self._solve_adaptive(dt_init, t_total, kwargs)
        else:
            dt = kwargs["dt"]  # required for fixed stepping (KeyError otherwise, as in the reference)
            t_total, note = self._end_time(kwargs)
            if not self._multi:
                self.history.append(note)
            self._solve_fixed(dt, t_total, kwargs)
        self.record_log("Problem solved")
        print("Problem solved.")
        return None

    def _device_model(self, kwargs, initial=None, table=None):
        """Model arrays on the device; with ``initial`` / ``table`` (per-step source values) their tensors are appended
        to the result (long meshes upload the three big arrays on parallel host threads, ``_engine.to_device_many``)."""
        nx = len(self.mesh)
        sidx = self._source_indices()
        sidx = np.where(sidx < 0, sidx + nx, sidx)
        pml = float(kwargs.get("pml_thickness", 0.0))
        smax = float(kwargs.get("sigma_max", 10 if self._multi else 2))
        if nx <= (1 << 16):
            # short models: everything goes up in ONE copy through a cached pinned buffer (four or five pageable copies
            # were 0.2 ms of C1's 0.9 ms end to end)
            sidx = np.atleast_1d(np.asarray(sidx)).astype(np.int64)
            n_src = int(sidx.size)
            dev = _engine.to_device_packed([(self.mesh, np.float64), (self.diffusivity, np.float64),
                                            (sidx, np.int64) if n_src else None,
                                            None if initial is None else (initial, np.float64),
                                            None if table is None else (table, np.float64)])
            out = (nx, dev[0], dev[1], dev[2], n_src, pml, smax)
            return out + (() if initial is None else (dev[3],)) + (() if table is None else (dev[4],))
        big = [self.mesh, self.diffusivity] + ([] if initial is None else [initial])
        dev = _engine.to_device_many(big)
        mesh_d, diff_d = dev[0], dev[1]
        sidx_d, n_src = _engine.index_tensor(sidx)
        out = (nx, mesh_d, diff_d, sidx_d, n_src, pml, smax)
        return out + (() if initial is None else (dev[2],)) + (() if table is None else (_engine.to_device(table),))

    def _solve_fixed(self, dt, t_total, kwargs):
        mode = kwargs.get("mode", "implicit")
        dtype = kwargs.get("dtype", "float64")
        every = int(kwargs.get("snapshot_every", 1))
        layout = kwargs.get("result_layout", "time_major")
        if dtype not in ("float64", "float32"):
            raise ValueError("dtype must be 'float64' or 'float32'.")
        if layout not in ("time_major", "cell_major"):
            raise ValueError("result_layout must be 'time_major' or 'cell_major'.")
        cell_major = layout == "cell_major" and mode == "implicit"
        self._data_cell_major = None
        if every < 1:
            raise ValueError("snapshot_every must be >= 1.")
        # exact replica of the reference's time axis (fp64 accumulation decides the step count)
        taxis, taxis_arr = fixed_time_axis_pair(self.t0, dt, t_total)  # the reference's list + (long axes) its float64 array
        n_steps = len(taxis) - 1
        self.record_log("Start to solve the problem.")
        initial = np.asarray(self.initial, dtype=np.float64)
        if n_steps == 0:
            self.snapshot = np.array([self.initial])
            self.taxis = np.array(taxis)
            return
        table = self._source_table(taxis[:-1] if taxis_arr is None else taxis_arr[:-1])
        taxis_out = np.array(taxis) if taxis_arr is None else taxis_arr
        nx, mesh_d, diff_d, sidx_d, n_src, pml, smax, p0_d, table_d = self._device_model(kwargs, initial[None, :], table)
        one_launch = (mode == "implicit" and dtype == "float64" and 2 <= nx <= _engine.run_max_nx()
                      and n_src <= _K3_MAX_SOURCES and kwargs.get("on_chip_assembly", True))
        if one_launch:
            # the whole solve is ONE kernel: the matrix is assembled (K1's arithmetic) and factorised in registers by the
            # run kernel itself (fbr_run_zonal_f64 with one diffusivity per cell) — no (nx) coefficient arrays, no
            # separate assembly / factorisation launches (C1: 3 launches and a 44 us one-thread factorisation before)
            sigma_d = _engine.pml_sigma(mesh_d, pml, smax) if pml != 0.0 else None
            try:
                snap_d, _, _ = _engine.run(None, p0_d, 1, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                                           snapshot_every=every, cell_major=cell_major,
                                           zonal=(mesh_d, diff_d.reshape(1, nx), None, dt, sigma_d))
                if cell_major and snap_d is not None:
                    snap_d = _CellMajor(snap_d[0])
            except np.linalg.LinAlgError:
                # a zero pivot WITHOUT row interchanges (PML on a Neumann row with dt * sigma = 1 ...): the reference's
                # LAPACK solve pivots (PDESolver_IMP.py:27-28) — so does this fallback, one pivoted solve per step
                # (DGTSV algorithm); a truly singular matrix raises LinAlgError from there, as in the reference
                diag = _engine.assemble(mesh_d, diff_d, 1, nx, dt, self.lbc, self.rbc, sidx_d, n_src, pml, smax)
                snap_d = self._step_loop(*diag, p0_d, sidx_d, n_src, table_d, n_steps, every, variant=_cabi.SOLVER_PIVOT)
                if cell_major and snap_d is not None:
                    snap_d = _CellMajor(_engine.transpose(snap_d[0]))
            if snap_d is None:
                self.snapshot = initial[None, :].copy()
            elif isinstance(snap_d, _CellMajor):
                self._data_cell_major = _engine.to_host(snap_d.t)
                self.snapshot = self._data_cell_major.T
            else:
                self.snapshot = _engine.to_host(snap_d[0])
            kept = np.arange(0, n_steps + 1, every)[: 1 + n_steps // every]
            self.taxis = taxis_out[kept] if every > 1 else taxis_out
            self._log_steps(taxis, table, kwargs)
            self.record_log(f"{n_steps} full steps solved on device (dt={dt}, mode={mode}, dtype={dtype}). "
                            f"taxis[-1]=", taxis[-1])
            if kwargs.get("print_progress"):
                for n in range(n_steps):
                    vals = table[n].tolist()
                    print("Time:", taxis[n + 1], "Source term:", vals if self._multi else vals[0])
            return
        dl, d, du = _engine.assemble(mesh_d, diff_d, 1, nx, dt, self.lbc, self.rbc, sidx_d, n_src, pml, smax)
        if mode == "implicit":
            try:
                factors = _engine.factorize(dl, d, du)
            except np.linalg.LinAlgError:
                if nx > (1 << 22):
                    raise
                factors = None  # zero pivot without row interchanges: pivoted solve per step, as above
            if factors is None:
                snap_d = self._step_loop(dl, d, du, p0_d, sidx_d, n_src, table_d, n_steps, every, variant=_cabi.SOLVER_PIVOT)
                if dtype == "float32" and snap_d is not None:
                    snap_d = snap_d.float()
            elif nx <= _engine.run_max_nx() and n_src > _K3_MAX_SOURCES:
                # more sources than the on-chip kernel's descriptors hold (the reference takes any number): one
                # right-hand side + one register-resident solve per step (fbr_rhs_f64 + fbr_solve_tridiag_f64)
                snap_d = self._step_loop(dl, d, du, p0_d, sidx_d, n_src, table_d, n_steps, every)
            elif nx <= _engine.run_max_nx():  # K3: one CTA keeps the whole model in registers
                snap_d, _, _ = _engine.run(factors, p0_d, 1, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                                           snapshot_every=every, dtype=dtype, cell_major=cell_major)
                if cell_major and snap_d is not None:  # (1, nx, rows + 1): the kernel wrote Data2D's layout
                    snap_d = _CellMajor(snap_d[0])
            else:
                if cell_major or dtype == "float32":  # keep the rows on the device: they are transposed there, then copied home once
                    kwargs = dict(kwargs, stream_rows=False)
                snap_d = self._run_long_mesh(factors, p0_d, nx, n_steps, sidx_d, n_src, table_d, every, kwargs)
                if dtype == "float32" and snap_d is not None and snap_d.dtype.itemsize == 8:
                    # stiff long meshes in the fp32 mode: the exact kernels integrate in fp64, the stored rows are rounded
                    self.long_mesh_path["arithmetic"] = "fp64, rows rounded to fp32"
                    snap_d = snap_d.float()
            if cell_major and snap_d is not None and not isinstance(snap_d, _CellMajor):
                snap_d = _CellMajor(_engine.transpose(_engine.as_f64(snap_d[0])))  # (rows + 1, nx) -> (nx, rows + 1) on the device
            if snap_d is None:  # snapshot_every > n_steps: only the initial state is kept
                self.snapshot = initial[None, :].copy()
            elif isinstance(snap_d, _CellMajor):
                # result_layout='cell_major': the host array IS Data2D's (cell, time) matrix; `snapshot` stays the
                # reference's (time, cell) object as a transposed view of it (no copy either way)
                self._data_cell_major = _engine.to_host(_engine.as_f64(snap_d.t))
                self.snapshot = self._data_cell_major.T
            elif isinstance(snap_d, np.ndarray):  # rows already streamed to the host
                self.snapshot = snap_d
            else:  # (1, rows + 1, nx) with row 0 = initial, written in place by the kernel
                self.snapshot = _engine.to_host(_engine.as_f64(snap_d[0]))  # fp32 mode: widened on the device
        else:
            rows = self._jacobi_loop(dl, d, du, p0_d, sidx_d, n_src, table, n_steps, every)
            self.snapshot = np.concatenate([initial[None, :], rows], axis=0)
        kept = np.arange(0, n_steps + 1, every)[: 1 + n_steps // every]
        self.taxis = taxis_out[kept] if every > 1 else taxis_out
        self._log_steps(taxis, table, kwargs)
        self.record_log(f"{n_steps} full steps solved on device (dt={dt}, mode={mode}, dtype={dtype}). "
                        f"taxis[-1]=", taxis[-1])
        if kwargs.get("print_progress"):
            for n in range(n_steps):
                vals = table[n].tolist()
                print("Time:", taxis[n + 1], "Source term:", vals if self._multi else vals[0])

    def _log_steps(self, taxis, table, kwargs):
        """``log_steps=True`` (additive kwarg): the two ``history`` lines per time step the reference records inside its
        loop (pds.py:290,310 / 556,575) — "Time: t Source term: v" and "Full step solution done. taxis[-1]= t'".  Off by
        default: the device integrates all steps in one launch, and 2 x n_steps strings cost more host time than C1's
        whole solve; one summary line is recorded instead."""
        if not kwargs.get("log_steps"):
            return
        for n in range(len(taxis) - 1):
            vals = table[n].tolist()
            self.record_log("Time:", taxis[n], "Source term:", vals if self._multi else vals[0])
            self.record_log("Full step solution done. taxis[-1]=", taxis[n + 1])

    def _run_long_mesh(self, factors, p0_d, nx, n_steps, sidx_d, n_src, table_d, every, kwargs):
        """Meshes that do not fit one CTA.  ``long_mesh`` (additive kwarg): 'auto' (default) takes the
        time-tiled window kernel K3W when the measured decay horizon of the factors makes it pay,
        else the exact kernels; 'windowed' insists on K3W (NotImplementedError when the horizon is too
        long); 'exact' uses only the exact kernels — K3L (cooperative, register-resident) up to its
        capacity, K3P (one pass over HBM per step) beyond.  ``dtype='float32'``: K3W's fp32 instantiation
        (16 cells per thread); the exact kernels stay fp64 and their rows are rounded."""
        which = kwargs.get("long_mesh", "auto")
        if which not in ("auto", "windowed", "exact"):
            raise ValueError("long_mesh must be 'auto', 'windowed' or 'exact'.")
        self.long_mesh_path = None
        if which != "exact":
            horizon = _engine.decay_horizon(factors, log2_eps=float(kwargs.get("window_log2_eps", -80.0)))
            f32 = kwargs.get("dtype", "float64") == "float32"
            plan = _engine.window_plan(nx, horizon, every, dtype="float32" if f32 else "float64")
            if plan is None and which == "windowed":
                raise NotImplementedError(f"decay horizon {horizon} is too long for the windowed path")
            if plan is not None:
                self.long_mesh_path = dict(kernel="K3W", horizon=horizon, **plan)
                # rows of >= 8 MB — or >= 512 KB each when the stored rows add up to >= 32 MB (C2: 101 rows of 0.8 MB, whose
                # 80 MB copy otherwise trails the 7.7 ms run by 3 ms): stream them to the host while later blocks are integrated
                n_rows = n_steps // every if every else 0
                big_rows = nx * 8 >= (8 << 20) or (nx * 8 >= (512 << 10) and n_rows * nx * 8 >= (32 << 20))
                if big_rows and kwargs.get("stream_rows", True):
                    self.long_mesh_path["rows"] = "streamed to host"
                    return _engine.run_window_to_host(factors, p0_d, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src,
                                                      table_d, horizon, snapshot_every=every)
                snap_d, _ = _engine.run_window(factors, p0_d, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src,
                                               table_d, horizon, snapshot_every=every, dtype="float32" if f32 else "float64")
                return snap_d
        if nx <= min(_engine.run_long_max_nx(), int(kwargs.get("_stream_above", 1 << 62))):
            self.long_mesh_path = dict(kernel="K3L")  # tiled over one cooperative grid
            snap_d, _ = _engine.run_long(factors, p0_d, 1, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src,
                                         table_d, snapshot_every=every)
        else:  # state and factors stream from HBM every step: one pass per step (K3P; 'levels' = K3S of round 1)
            which_exact = kwargs.get("_exact_kernel", "pass")
            self.long_mesh_path = dict(kernel="K3P" if which_exact == "pass" else "K3S")
            tile_warps = None
            if which_exact == "pass":  # narrow tiles on ordinary meshes, wide ones on stiff meshes (K3P's own domain)
                tile_warps = kwargs.get("_tile_warps") or _engine.pass_tile_warps(factors, n_steps,
                                                                                  horizon if which != "exact" else None)
                self.long_mesh_path["tile_warps"] = tile_warps
            snap_d, _ = _engine.run_stream(factors, p0_d, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                                           snapshot_every=every, kernel=which_exact, tile_warps=tile_warps)
        return snap_d

    def _step_loop(self, dl, d, du, p_d, sidx_d, n_src, table_d, n_steps, every, variant=_cabi.SOLVER_AUTO):
        """Fixed-dt loop with one library call pair per step (models with more than 253 sources; matrices that need the
        row interchanges of the reference's LAPACK solve: ``variant=SOLVER_PIVOT``)."""
        import torch
        rows = [p_d[0]]
        for n in range(n_steps):
            b = _engine.rhs(p_d, self.lbc, self.rbc, sidx_d, n_src, table_d[n])
            p_d = _engine.solve_tridiag(dl, d, du, b, variant)
            if (n + 1) % every == 0:
                rows.append(p_d[0])
        return torch.stack(rows)[None] if len(rows) > 1 else None

    def _jacobi_loop(self, dl, d, du, p_d, sidx_d, n_src, table, n_steps, every):
        """mode != 'implicit': one Jacobi solve per step (reference pds.py:296-300)."""
        import torch
        table_d = _engine.to_device(table)
        rows = []
        bad = 0
        for n in range(n_steps):
            b = _engine.rhs(p_d, self.lbc, self.rbc, sidx_d, n_src, table_d[n])
            p_d, iters = _engine.jacobi(dl, d, du, b)
            bad += int((iters <= 0).sum().item())
            if (n + 1) % every == 0:
                rows.append(p_d[0])
        if bad:
            print("Warning: Maximum iterations reached without convergence.")
        return torch.stack(rows).cpu().numpy() if rows else np.zeros((0, dl.shape[1]))

    def _solve_adaptive(self, dt_init, t_total, kwargs):
        """Step-doubling loop of reference pds.py:312-336.

        Default: the whole loop runs in one persistent kernel per chunk of accepted steps
        (``fbr_run_adaptive_f64``); only the accepted states come back.  Meshes above its size
        limit, PML runs, ``print_progress`` and duck-typed sources use the host-driven loop below,
        which keeps the states on the device and moves one 8-byte error per iteration."""
        nx = len(self.mesh)
        if kwargs.get("mode", "implicit") != "implicit":
            # the reference solves the FULL step of every iteration with Jacobi in this mode (pds.py:296-300, 562-569;
            # the two half steps stay implicit): host-driven loop, Jacobi kernel for the full step
            return self._solve_adaptive_host(dt_init, t_total, kwargs)
        device_loop = (nx <= _engine.adaptive_max_nx() and not kwargs.get("print_progress")
                       and float(kwargs.get("pml_thickness", 0.0)) == 0.0
                       and all(_uses_stock_interp(s) for s in self._sources())
                       and kwargs.get("adaptive_on_device", True))
        if device_loop:
            return self._solve_adaptive_device(dt_init, t_total, kwargs)
        return self._solve_adaptive_host(dt_init, t_total, kwargs)

    def _solve_adaptive_device(self, dt_init, t_total, kwargs):
        import torch
        nx, mesh_d, diff_d, sidx_d, n_src, _, _ = self._device_model(kwargs)
        srcs = self._sources()
        width = max(len(s.taxis) for s in srcs)
        st = np.zeros((len(srcs), width))
        sd = np.zeros((len(srcs), width))
        for k, s in enumerate(srcs):
            st[k, :len(s.taxis)] = np.asarray(s.taxis, dtype=np.float64)
            sd[k, :len(s.data)] = np.asarray(s.data, dtype=np.float64)
        st_d, sd_d = _engine.to_device(st), _engine.to_device(sd)
        len_d = _engine.to_device(np.array([len(s.taxis) for s in srcs], dtype=np.int32), np.int32)
        initial = np.asarray(self.initial, dtype=np.float64)
        state = _engine.to_device(initial[None, :])
        ctrl = _engine.to_device(np.array([[float(self.t0), float(dt_init)]]))
        tol = kwargs.get("tol", 1e-3)
        adj = dict(safety=kwargs.get("safety_factor", 0.9), p_order=kwargs.get("p", 2), max_dt=kwargs.get("max_dt", 40),
                   min_dt=kwargs.get("min_dt", 1e-4))
        budget = kwargs.get("max_iterations")
        capacity = int(max(16, min(4096, (64 << 20) // (8 * nx))))
        self.record_log("Start to solve the problem.")
        rows, times, used, accepted = [initial[None, :]], [np.array([float(self.t0)])], 0, 0
        while True:
            per_launch = 200000 if budget is None else max(0, min(200000, budget - used))
            snap, tax, counts = _engine.run_adaptive(mesh_d, diff_d, state, ctrl, self.lbc, self.rbc, sidx_d, n_src,
                                                     st_d, sd_d, len_d, self.t0, t_total, tol, 1e-3, adj["safety"],
                                                     adj["p_order"], adj["max_dt"], adj["min_dt"], per_launch, capacity)
            n_rows, n_it, done, bad = (int(v) for v in counts[0].tolist())
            if bad:
                raise np.linalg.LinAlgError(f"Matrix is singular (zero pivot at row {bad - 1}).")
            if n_rows:
                rows.append(snap[0, :n_rows].cpu().numpy())
                times.append(tax[0, :n_rows].cpu().numpy())
            used += n_it
            accepted += n_rows
            if done or (budget is not None and used >= budget):
                break
            if n_rows == 0 and n_it >= per_launch:
                raise RuntimeError("Adaptive time stepping made no progress in 200000 iterations (the step-doubling "
                                   "error is NaN or above tol at min_dt); the reference loops forever in this case.")
        self.snapshot = np.concatenate(rows, axis=0)
        self.taxis = np.concatenate(times)
        self.record_log(f"Dynamic time sampling on device: {accepted} steps accepted, {used - accepted} rejected, "
                        f"final dt = {float(ctrl[0, 1])}")

    def _solve_adaptive_host(self, dt_init, t_total, kwargs):
        import torch
        nx, mesh_d, diff_d, sidx_d, n_src, pml, smax = self._device_model(kwargs)
        self.snapshot = [self.initial]
        self.taxis = [self.t0]
        self.record_log("Start to solve the problem.")
        accepted = []  # device rows of the accepted states
        p_d = _engine.to_device(np.asarray(self.initial, dtype=np.float64)[None, :])
        dt = dt_init
        max_iter = kwargs.get("max_iterations")  # additive safety valve (the reference can spin forever)
        it = 0
        while self.taxis[-1] < t_total:
            if max_iter is not None and it >= max_iter:
                break
            it += 1
            vals = self._source_values_at(self.taxis[-1] - self.t0)
            self.record_log("Time:", self.taxis[-1], "Source term:", vals if self._multi else vals[0])
            vals_d = _engine.to_device(np.asarray(vals, dtype=np.float64))
            bc = (self.lbc, self.rbc, sidx_d, n_src, vals_d)  # fbr_step_f64: rhs overrides fused into the solve
            full_sys = _engine.assemble(mesh_d, diff_d, 1, nx, dt, self.lbc, self.rbc, sidx_d, n_src, pml, smax)
            if kwargs.get("mode", "implicit") == "implicit":
                full = _engine.step(*full_sys, p_d, *bc)
            else:  # reference pds.py:296-300: anything but 'implicit' is the Jacobi solver (warns when it does not converge)
                full, iters = _engine.jacobi(*full_sys, _engine.rhs(p_d, self.lbc, self.rbc, sidx_d, n_src, vals_d))
                if int((iters <= 0).sum().item()):
                    print("Warning: Maximum iterations reached without convergence.")
            half_sys = _engine.assemble(mesh_d, diff_d, 1, nx, dt / 2, self.lbc, self.rbc, sidx_d, n_src, pml, smax)
            mid = _engine.step(*half_sys, p_d, *bc)
            half = _engine.step(*half_sys, mid, *bc)
            error = tso.step_error(full, half)
            n_before = len(self.taxis)
            marker = _DeviceRow(len(accepted))
            dt = tso.decide(self, marker, error, dt, **kwargs)
            if len(self.taxis) > n_before:  # accepted: the FULL-step state becomes the new state
                accepted.append(full[0])
                p_d = full
            if kwargs.get("print_progress"):
                print("Time:", self.taxis[-1], "Source term:", vals if self._multi else vals[0])
        rows = torch.stack(accepted).cpu().numpy() if accepted else np.zeros((0, nx))
        self.snapshot = np.concatenate([np.asarray(self.initial, dtype=np.float64)[None, :], rows], axis=0)
        self.taxis = np.array(self.taxis)

    # ------------------------------------------------------------------ logging / results
    def record_log(self, *args):
        self.history.append(" ".join(map(str, args)) + f" | Time: {datetime.now()}")

    def print_log(self):
        for msg in self.history:
            print(msg)

    def get_solution(self):
        return self.snapshot, self.taxis

    def get_val_at_source_idx(self):
        return self.snapshot[:, self.sourceidx]

    def get_val_at_idx(self, idx, as_data1d=False):
        """Pressure history of cell ``idx`` (reference pds.py:364-366: the raw column).  ``as_data1d=True`` returns it
        as a ``Data1D`` channel (taxis of the run, the source's start_time) — the object the analyzer side works with."""
        if not as_data1d:
            return self.snapshot[:, idx]
        from fiberis_b200.analyzer.Data1D.core1D import Data1D
        first = self.source[0] if isinstance(self.source, list) else self.source
        return Data1D(data=np.array(self.snapshot[:, idx], dtype=np.float64), taxis=np.asarray(self.taxis, dtype=float).copy(),
                      start_time=getattr(first, "start_time", None), name=f"cell {int(idx)}")

    def get_val_at_time(self, time):
        k = np.argmin(np.abs(self.taxis - time))
        print("Closest time = ", self.taxis[k])
        return self.snapshot[k]

    def plot_solution(self, **kwargs):
        method = kwargs.get("method", "imshow")
        if method not in ("imshow", "pcolormesh"):
            raise ValueError("Method must be 'imshow' or 'pcolormesh'.")
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:  # plotting is optional tooling, not part of the compute path
            raise RuntimeError("plot_solution needs matplotlib, which is not installed.") from exc
        cmap = kwargs.get("cmap", "bwr")
        plt.figure()
        if method == "imshow":
            plt.imshow(self.snapshot.T, aspect="auto", cmap=cmap,
                       extent=[self.taxis[0], self.taxis[-1], self.mesh[0], self.mesh[-1]])
            plt.gca().invert_yaxis()
        else:
            plt.pcolormesh(self.taxis, self.mesh, self.snapshot.T, cmap=cmap)
            plt.colorbar()
        plt.xlabel("Time/s")
        plt.ylabel("Distance/ft")
        plt.show()

    def pack_result(self, **kwargs):
        """npz in the layout ``Data2D.load_npz`` reads: data is (distance, time)."""
        filename = kwargs.get("filename", "result.npz")
        if kwargs.get("mode", "fiberis") != "fiberis":
            raise ValueError("Mode must be 'fiberis' data format.")
        first = self.source[0] if isinstance(self.source, list) else self.source
        np.savez(filename, daxis=self.mesh, taxis=self.taxis, data=self.snapshot.T, start_time=first.start_time)

    def to_data2d(self, name=None):
        """The result as a ``Data2D`` object without the npz round trip (additive helper)."""
        from fiberis_b200.analyzer.Data2D.core2D import Data2D
        first = self.source[0] if isinstance(self.source, list) else self.source
        # after solve(result_layout='cell_major') snapshot.T is the C-contiguous array the kernel wrote: no copy here
        return Data2D(data=np.ascontiguousarray(self.snapshot.T), taxis=np.asarray(self.taxis, dtype=float),
                      daxis=np.asarray(self.mesh, dtype=float), start_time=getattr(first, "start_time", None),
                      name=name)

    def reset(self):
        self.taxis = None
        self.snapshot = None
        self._data_cell_major = None
        self.history = []


class _CellMajor:
    """Marks a device tensor that already has the (cell, time) layout."""

    def __init__(self, t):
        self.t = t


class _DeviceRow:
    """Placeholder appended to ``snapshot`` by ``tso.decide`` while the accepted state stays on the
    device; replaced by the real rows when the loop ends."""

    def __init__(self, k):
        self.k = k


class PDS1D_MultiSource(PDS1D_SingleSource):
    """Several independent pressure sources: ``source`` is a list, ``sourceidx`` a list of nodes."""

    _multi = True

    def set_sourceidx(self, sourceidx):
        super().set_sourceidx(sourceidx)
        self.sourceidx = np.array(self.sourceidx)

    def _check_source(self):
        if self.source is None:
            return False, "Source term is not set."
        if not isinstance(self.source, list):
            return False, "Source term must be a list."
        return True, "Source term is properly set."

    def _check_sourceidx(self):
        if self.sourceidx is None:
            return False, "Source index is not set."
        if isinstance(self.sourceidx, int):
            return False, ("Source index must be a list of integers. If you want to set a single source index, "
                           "please use PDS1D_SingleSource.")
        for idx in self.sourceidx:  # reference quirk: a raw list of Python ints is rejected; the setter
            if isinstance(idx, int):  # converts to an ndarray afterwards, whose entries are np.int64
                return False, "Source index must be a list of integers."
        for idx in self.sourceidx:
            if not _in_mesh(idx, len(self.mesh)):
                return False, "Source index is not in the mesh."
        if self.source is not None and len(self.sourceidx) != len(self.source):
            return False, "Source index and source term length do not match."
        return True, "Source index is properly set."
