"""``PDS1D_Ensemble`` — many independent diffusion models stepped on one GPU (new capability).

The reference has no batched driver (``fiberis.simulator.optimizer`` only holds the adaptive-dt
controller); ensembles are what BASELINE.json's configs C3/C4 call "optimizer parameter sweep" /
"ensemble sweep".  Each member is exactly a ``PDS1D_MultiSource`` problem
(/root/reference/src/fiberis/simulator/core/pds.py:432-610) and its pressures are pinned by looping
the reference over members.  Members share mesh, sources, boundary conditions and time axis; they
differ in diffusivity (dense ``(M, nx)`` or zonal ``zone_value[m, zone_of_cell]``) and optionally in
the initial state.

Outputs: ``misfit`` (M,) — sum over steps and observation cells of w_k (p - obs)^2, accumulated on
the fly inside the time-stepping kernel so member snapshots never leave the chip — and full
snapshots of the few ``audit_models`` asked for.  The misfit definition is this package's own
(parity unpinned; the reference has none in ``fiberis.simulator``).

Multi-GPU: members are sharded contiguously over the ranks of a ``torch.distributed`` process
group (``shard_bounds``); the only collective is an all-gather of the per-member misfits.
"""
from __future__ import annotations

import numpy as np

from fiberis_b200 import _engine
from fiberis_b200.simulator.core.pds import PDS1D_MultiSource, _ALLOWED_BC, fixed_time_axis, fixed_time_axis_array


def shard_bounds(n_models: int, world_size: int, rank: int):
    """Contiguous block [lo, hi) of members owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_models, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class PDS1D_Ensemble:
    def __init__(self):
        self.mesh = None
        self.source = None           # list of Data1D-like, shared by all members
        self.sourceidx = None
        self.lbc = self.rbc = None
        self.t0 = 0
        self.initial = None          # (nx,) shared or (M, nx)
        self.diffusivity = None      # (M, nx) or None when zonal
        self.zone_value = None       # (M, Z)
        self.zone_of_cell = None     # (nx,) int
        self.obs_idx = None
        self.obs = None              # (n_steps + 1, n_obs), row 0 unused
        self.obs_weights = None
        self.obs_taxis = None        # observations on their own time axis (None: on the step times)
        self.obs_baseline = None
        self.obs_window = None
        self.misfit = None
        self.snapshot = None         # (n_audit, n_t, nx)
        self.taxis = None
        self.snapshot_taxis = None
        self.audit_models = None
        self.history = []

    # -- configuration (same vocabulary as PDS1D_*) ------------------------------------------------
    def set_mesh(self, mesh):
        if not isinstance(mesh, np.ndarray) or mesh.ndim != 1 or len(mesh) < 2:
            raise ValueError("Mesh must be a 1D numpy array with at least 2 points.")
        self.mesh = mesh

    def set_source(self, source):
        if not isinstance(source, list):
            raise ValueError("Source term must be a list.")
        self.source = source

    def set_sourceidx(self, sourceidx):
        self.sourceidx = np.atleast_1d(np.asarray(sourceidx)).astype(np.int64)

    def set_bcs(self, lbc="Neumann", rbc="Neumann"):
        if lbc not in _ALLOWED_BC or rbc not in _ALLOWED_BC:
            raise ValueError(f"Boundary conditions must be in {_ALLOWED_BC}.")
        self.lbc, self.rbc = lbc, rbc

    def set_t0(self, t0):
        self.t0 = t0

    def set_initial(self, initial):
        self.initial = np.asarray(initial, dtype=np.float64)

    def set_diffusivity(self, diffusivity):
        """Dense per-member diffusivity (M, nx)."""
        d = np.asarray(diffusivity, dtype=np.float64)
        if d.ndim != 2 or self.mesh is None or d.shape[1] != len(self.mesh):
            raise ValueError("Ensemble diffusivity must be an (M, nx) array; set the mesh first.")
        self.diffusivity, self.zone_value, self.zone_of_cell = d, None, None

    def set_zonal_diffusivity(self, zone_value, zone_of_cell):
        """D_m(x) = zone_value[m, zone_of_cell[x]] (e.g. zone_value = 10 ** theta)."""
        zv = np.asarray(zone_value, dtype=np.float64)
        zc = np.asarray(zone_of_cell)
        if zv.ndim != 2 or self.mesh is None or zc.shape != (len(self.mesh),):
            raise ValueError("zone_value must be (M, Z) and zone_of_cell (nx,); set the mesh first.")
        if zc.min() < 0 or zc.max() >= zv.shape[1]:
            raise ValueError("zone_of_cell refers to a zone outside zone_value.")
        self.zone_value, self.zone_of_cell, self.diffusivity = zv, zc.astype(np.int32), None

    def set_observations(self, obs_idx, obs, weights=None, taxis=None, baseline_index=None, window=None):
        """Observed channels for the fused misfit.

        Without ``taxis``: ``obs[n, k]`` is the observed pressure at cell ``obs_idx[k]`` at the step time
        ``taxis[n]`` of the run (row 0 is not used) and ``misfit = sum_n sum_k w_k (p - obs)^2``.

        With ``taxis`` (seconds on the run's own time axis ``self.taxis``, i.e. including ``t0``; non-decreasing): the observations keep their OWN time
        axis, ``obs`` is ``(len(taxis), n_obs)`` and the misfit is the reference's
        (src/fiberis/moose/templates/baseline_model_generator.py:394-403): every simulated channel is interpolated
        linearly onto ``taxis`` (``Data1D.interpolate``), the sample ``baseline_index`` is subtracted from it (the
        reference uses 1; ``None``: nothing), and ``w_k (sim - obs)^2`` is summed over the samples of ``window =
        (first, last)`` (the reference uses ``(1, -1)``; default: all samples).  Observed times outside the simulated
        range inside the window make the misfit NaN, as the reference's interpolation (fill = NaN) does."""
        idx = np.atleast_1d(np.asarray(obs_idx)).astype(np.int64)
        if len(np.unique(idx)) != len(idx):
            raise ValueError("Observation cells must be distinct.")
        if self.mesh is not None and (idx.min() < 0 or idx.max() >= len(self.mesh)):
            raise ValueError("Observation cells must lie in the mesh.")
        obs = np.asarray(obs, dtype=np.float64)
        if obs.ndim != 2 or obs.shape[1] != len(idx):
            raise ValueError("obs must have shape (n_samples, n_obs).")
        if taxis is not None:
            taxis = np.asarray(taxis, dtype=np.float64)
            if taxis.shape != (obs.shape[0],) or np.any(np.diff(taxis) < 0):
                raise ValueError("Observed taxis must be non-decreasing with one entry per row of obs.")
        elif baseline_index is not None or window is not None:
            raise ValueError("baseline_index / window need the observations' own taxis.")
        self.obs_idx, self.obs = idx, obs
        self.obs_weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.obs_taxis, self.obs_baseline, self.obs_window = taxis, baseline_index, window

    def set_observed_data(self, observed, obs_idx, channels=None, weights=None, baseline_index=1, window=(1, -1)):
        """Observations from a ``Data2D``-like object (``data (n_channels, n_time)``, ``taxis`` in seconds after
        ``start_time``) with the reference's misfit conventions as defaults.  ``channels``: rows of ``observed.data``
        matched with the cells ``obs_idx`` (default: the first ``len(obs_idx)``).  The observed ``start_time`` is
        aligned with the first source's ``start_time`` when both are set (``Data1D.interpolate``'s time offset)."""
        idx = np.atleast_1d(np.asarray(obs_idx)).astype(np.int64)
        rows = np.arange(len(idx)) if channels is None else np.atleast_1d(np.asarray(channels)).astype(np.int64)
        data = np.asarray(observed.data, dtype=np.float64)[rows]
        taxis = np.asarray(observed.taxis, dtype=np.float64).copy()
        sim_start = getattr(self.source[0], "start_time", None) if self.source else None
        obs_start = getattr(observed, "start_time", None)
        if sim_start is not None and obs_start is not None:
            taxis = taxis + (obs_start - sim_start).total_seconds()
        self.set_observations(idx, np.ascontiguousarray(data.T), weights=weights, taxis=taxis,
                              baseline_index=baseline_index, window=window)

    @property
    def n_models(self):
        if self.diffusivity is not None:
            return self.diffusivity.shape[0]
        if self.zone_value is not None:
            return self.zone_value.shape[0]
        return 0

    def member(self, m):
        """Member ``m`` as a stand-alone ``PDS1D_MultiSource`` (the parity cross-check uses this)."""
        p = PDS1D_MultiSource()
        p.set_mesh(self.mesh)
        p.set_source(self.source)
        p.set_bcs(self.lbc, self.rbc)
        init = self.initial if self.initial.ndim == 1 else self.initial[m]
        p.set_initial(init)
        D = self.diffusivity[m] if self.diffusivity is not None else self.zone_value[m][self.zone_of_cell]
        p.set_diffusivity(np.array(D))
        p.set_sourceidx(list(self.sourceidx))
        p.set_t0(self.t0)
        return p

    # -- solve ------------------------------------------------------------------------------------
    def solve(self, dt, t_total=None, audit_models=(), snapshot_every=1, dtype="float64", models=None,
              process_group=None, gather=True, split_audit=True, on_chip_assembly=True, scan_log2_eps=0.0,
              result_layout="time_major"):
        """Integrate every member with a fixed ``dt``.

        Ensembles in fp64 with 129..4096 cells assemble and factorise their matrices inside the run
        kernel (``fbr_run_zonal_f64``; ``on_chip_assembly=False`` selects the three-kernel path K1 + K1b + K3).

        ``result_layout='cell_major'``: the audited members' snapshots are written by the kernel as ``(member, cell,
        time)`` — ``Data2D``'s layout; ``snapshot`` stays ``(member, time, cell)`` as a transposed view of that array.

        ``scan_log2_eps``: threshold below which a member's far scan terms are skipped (``fbr_run_opts``; default
        2^-80, a positive value keeps every member's scans exact).

        ``models=(lo, hi)`` restricts the run to a contiguous block (multi-GPU sharding does this
        per rank when ``process_group`` / an initialised default group is given); ``misfit`` is then
        gathered over ranks unless ``gather=False``.
        """
        M_all = self.n_models
        if M_all == 0:
            raise ValueError("No diffusivity set.")
        nx = len(self.mesh)
        if t_total is None:
            t_total = self.source[0].taxis[-1] + self.t0
        taxis = fixed_time_axis_array(self.t0, dt, t_total)  # same fp64 accumulation as the single-model path
        n_steps = len(taxis) - 1

        dist = _dist(process_group)
        if models is None and dist is not None:
            models = shard_bounds(M_all, dist.get_world_size(process_group), dist.get_rank(process_group))
        sharded = models is None and dist is not None
        lo, hi = (0, M_all) if models is None else models
        M = hi - lo
        if dist is not None and gather and not sharded and (lo, hi) != shard_bounds(
                M_all, dist.get_world_size(process_group), dist.get_rank(process_group)):
            raise ValueError("gather=True assembles the misfits of the ranks' own shards (shard_bounds); with an explicit "
                             "`models` block that is not this rank's shard pass gather=False.")
        if M == 0:  # more ranks than members: nothing to integrate here, but the collective still needs this rank
            import torch
            self.taxis = np.array(taxis)
            self.snapshot, self.audit_models = None, []
            empty = torch.empty(0, dtype=torch.float64, device=_engine.device())
            self.misfit = (gather_misfit(empty, M_all, process_group) if (dist is not None and gather) else empty).cpu().numpy()
            return self.misfit

        proto = PDS1D_MultiSource()
        proto.source, proto.t0 = self.source, self.t0
        table = proto._source_table(taxis[:-1])

        # Every small input goes up in ONE copy (a dozen pageable copies cost 0.5 ms of host time in front of the first launch)
        audit = [int(a) for a in audit_models if lo <= int(a) < hi]
        # the kernel records every listed model once: repeated entries share one recorded slot
        audit_unique, audit_slot = np.unique(np.asarray(audit, dtype=np.int64), return_inverse=True)
        n_audit = len(audit_unique)
        sidx = np.atleast_1d(np.asarray(self.sourceidx)).astype(np.int64)
        n_src = int(sidx.size)
        init = self.initial if self.initial.ndim == 1 else self.initial[lo:hi]
        member_values = self.zone_value[lo:hi] if self.diffusivity is None else self.diffusivity[lo:hi]
        n_obs = 0
        sampling = None
        obs_items = [None, None, None]
        if self.obs_idx is not None:
            if self.obs_taxis is None and self.obs.shape[0] != n_steps + 1:
                raise ValueError(f"obs has {self.obs.shape[0]} rows, the run has {n_steps + 1} time samples.")
            oi = np.atleast_1d(np.asarray(self.obs_idx)).astype(np.int64)
            n_obs = int(oi.size)
            obs_items = [(oi, np.int64) if n_obs else None, (self.obs, np.float64),
                         None if self.obs_weights is None else (self.obs_weights, np.float64)]
        (mesh_d, sidx_d, values_d, zone_d, p0_d, table_d, snap_models_d, obs_idx_d, obs_d, w_d) = _engine.to_device_packed([
            (self.mesh, np.float64), (sidx, np.int64) if n_src else None, (member_values, np.float64),
            None if self.diffusivity is not None else (self.zone_of_cell, np.int32), (init, np.float64), (table, np.float64),
            (audit_unique - lo, np.int64) if n_audit else None] + obs_items)
        if self.obs_idx is not None and self.obs_taxis is not None:
            # interpolation plan of the observed times on this run's time axis
            sampling = _engine.sample_plan(np.asarray(taxis), self.obs_taxis, self.obs_baseline, self.obs_window)
        # fp64 ensembles on a shared mesh: the matrices are assembled and factorised inside the run kernel
        # (fbr_run_zonal_f64 / _f32), no (M, nx) coefficient array exists.  Short meshes: K1 + K1b + K3.
        zonal = flags_d = None
        on_chip = on_chip_assembly and nx <= _engine.run_max_nx() and (
            (dtype == "float64" and nx > 128 and (nx >= 256 or M >= 32)) or (dtype == "float32" and nx >= 512))
        if on_chip:
            # zone_d is None: one diffusivity per cell and member, the (M, nx) array is read once, 8 B per cell
            zonal = (mesh_d, values_d, zone_d, dt, None)
            flags_d = _engine.zonal_flags(M, mesh_d.device)
            factors = None
        else:
            if self.diffusivity is not None:
                diag = _engine.assemble(mesh_d, values_d, M, nx, dt, self.lbc, self.rbc, sidx_d, n_src)
            else:
                diag = _engine.assemble_zonal(mesh_d, values_d, zone_d, M, nx, dt, self.lbc, self.rbc, sidx_d, n_src)
            factors = _engine.factorize(*diag)
            del diag
        if result_layout not in ("time_major", "cell_major"):
            raise ValueError("result_layout must be 'time_major' or 'cell_major'.")
        cell_major = result_layout == "cell_major"
        run_kw = dict(obs_idx=obs_idx_d, n_obs=n_obs, obs=obs_d, obs_w=w_d, dtype=dtype, sampling=sampling,
                      scan_log2_eps=scan_log2_eps, cell_major=cell_major)
        n_rows = n_steps // snapshot_every if n_audit else 0
        host_snap = None
        # sweeps of a few members per CTA: the costly members (long decay horizons: large diffusivity) are handed out first
        order_d = _engine.costly_members_first(zonal[1], M, nx) if on_chip and n_steps >= 64 else None
        if n_rows and M >= 8 * n_audit and M >= 512 and split_audit:
            # Large sweep, few audited members: integrate the audited ones as their own small launch on a
            # side stream and bring their snapshots home WHILE the big launch is still running (the
            # device-to-host copy of a few hundred MB otherwise trails the run); the big launch then
            # records nothing.  The audited members' factor rows are gathered into a contiguous batch.
            import torch
            main, side = torch.cuda.current_stream(), _engine._copy_stream(mesh_d.device)
            pick = snap_models_d
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if on_chip:
                    fac_a, zon_a = None, (zonal[0], zonal[1].index_select(0, pick).contiguous(), zonal[2], dt, None)
                    flags_a = _engine.zonal_flags(n_audit, mesh_d.device)  # duplicates of the sweep's flags: not checked
                else:
                    fac_a, zon_a, flags_a = tuple(f.index_select(0, pick).contiguous() for f in factors), None, None
                p0_a = p0_d if p0_d.dim() == 1 else p0_d.index_select(0, pick).contiguous()
                # same observation arguments as the sweep: equal shared-memory footprints let CTAs of the two
                # launches share an SM (with different footprints the audited CTAs kept their SMs to
                # themselves and the sweep finished 2.5 ms later); the misfits computed here are discarded
                # fp32 mode: the rows leave the kernel as doubles (a widening pass here would have to squeeze past the
                # sweep's persistent grid: 4.6 ms for C3's 82 MB on the four free CTA slots)
                wide = dtype == "float32" and on_chip and n_audit <= 64 and nx >= 512
                snap_a, _, _ = _engine.run(fac_a, p0_a, n_audit, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                                           snapshot_every=snapshot_every, zonal=zon_a, singular_flags=flags_a,
                                           snap_f64=wide, **run_kw)
                host_snap = _engine.to_host(_engine.as_f64(snap_a), sync=False)
            # reserve_ctas: the audited members' CTAs stay resident next to the sweep's grid
            _, _, misfit_d = _engine.run(factors, p0_d, M, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                                         snapshot_every=0, zonal=zonal, singular_flags=flags_d, reserve_ctas=n_audit,
                                         member_order=order_d, **run_kw)
            main.wait_stream(side)
            snap_d = snap_a
        else:
            snap_d, _, misfit_d = _engine.run(
                factors, p0_d, M, nx, n_steps, self.lbc, self.rbc, sidx_d, n_src, table_d,
                snapshot_every=snapshot_every if n_audit else 0, snap_models=snap_models_d, n_snap_models=n_audit,
                zonal=zonal, singular_flags=flags_d, member_order=None if n_audit else order_d, **run_kw)
        if flags_d is not None:
            _engine.raise_if_singular_zonal(flags_d)

        self.taxis = np.array(taxis)
        # times of the recorded rows of ``snapshot`` (row 0 = initial state, then every ``snapshot_every`` steps)
        self.snapshot_taxis = self.taxis[np.arange(0, n_steps + 1, snapshot_every)[: 1 + n_steps // snapshot_every]]
        self.audit_models = audit
        if n_audit and snap_d is not None:  # (n_audit, n_t, nx), row 0 = initial, straight from the kernel
            if host_snap is not None:
                side.synchronize()
                self.snapshot = host_snap
            else:
                self.snapshot = _engine.to_host(_engine.as_f64(snap_d))  # fp32 mode: widened on the device, not on the host
            if len(audit_unique) != len(audit) or np.any(audit_slot != np.arange(len(audit))):
                self.snapshot = self.snapshot[audit_slot]  # back to the order (and repeats) the caller listed
            if cell_major:  # the array is (member, cell, time): expose the reference's (time, cell) order as a view
                self.snapshot = np.transpose(self.snapshot, (0, 2, 1))
        elif n_audit:  # snapshot_every > n_steps: only the initial state is kept, as in the single-model path
            rows = np.asarray(self.initial, dtype=np.float64)
            self.snapshot = np.stack([(rows if rows.ndim == 1 else rows[a]).copy() for a in audit])[:, None, :]
        else:
            self.snapshot = None
        if misfit_d is not None:
            if dist is not None and gather:
                misfit_d = gather_misfit(misfit_d, M_all, process_group)
            self.misfit = misfit_d.cpu().numpy()
        self.history.append(f"{M} members x {nx} cells x {n_steps} steps solved on device (dt={dt}).")
        return self.misfit


    # -- adaptive time stepping, every member with its own dt ------------------------------------------
    def solve_adaptive(self, dt_init, t_total=None, audit_models=(), models=None, max_iterations=None, **kwargs):
        """Every member runs the reference's step-doubling loop (``solve(optimizer=True)``,
        src/fiberis/simulator/core/pds.py:312-336 with src/fiberis/simulator/optimizer/tso.py:4-50)
        on the device, each with its OWN time step, accept/reject history and time axis — SURVEY.md
        8(f1).  One CTA per member (``fbr_run_adaptive_f64``), no host round trip per iteration.

        Keyword arguments are the tso ones (``tol``, ``safety_factor``, ``p``, ``max_dt``, ``min_dt``; the
        user's ``tol`` is not forwarded to ``adjust_dt``, as in the reference).  Returns, and keeps in
        ``self.adaptive``, a dict: ``final_state (M, nx)``, ``final_time (M,)``, ``final_dt (M,)``, ``accepted (M,)``,
        ``iterations (M,)``, and for the audited members the ragged ``taxis`` / ``snapshot`` lists (row 0 =
        initial state).  ``max_iterations`` bounds every member's loop (the reference can spin forever
        when the error estimate is NaN); a member that stalls without it raises ``RuntimeError``."""
        import torch
        M_all = self.n_models
        if M_all == 0:
            raise ValueError("No diffusivity set.")
        nx = len(self.mesh)
        if nx > _engine.adaptive_max_nx():
            raise NotImplementedError(f"device-side adaptive stepping covers nx <= {_engine.adaptive_max_nx()}")
        if t_total is None:
            t_total = self.source[0].taxis[-1] + self.t0
        lo, hi = (0, M_all) if models is None else models
        M = hi - lo
        mesh_d = _engine.to_device(self.mesh)
        if self.diffusivity is not None:
            diff_d = _engine.to_device(self.diffusivity[lo:hi])
        else:  # zonal members: gather the (M, nx) diffusivity once on the device (buffer plumbing)
            zone_d = _engine.to_device(self.zone_of_cell, np.int64)
            diff_d = _engine.to_device(self.zone_value[lo:hi])[:, zone_d].contiguous()
        sidx_d, n_src = _engine.index_tensor(self.sourceidx)
        width = max(len(s.taxis) for s in self.source)
        st, sd = np.zeros((len(self.source), width)), np.zeros((len(self.source), width))
        for k, src in enumerate(self.source):
            st[k, :len(src.taxis)] = np.asarray(src.taxis, dtype=np.float64)
            sd[k, :len(src.data)] = np.asarray(src.data, dtype=np.float64)
        st_d, sd_d = _engine.to_device(st), _engine.to_device(sd)
        len_d = _engine.to_device(np.array([len(s.taxis) for s in self.source], dtype=np.int32), np.int32)
        init = self.initial if self.initial.ndim == 2 else np.broadcast_to(self.initial, (M_all, nx))
        state = _engine.to_device(np.ascontiguousarray(init[lo:hi]))
        ctrl = _engine.to_device(np.tile(np.array([[float(self.t0), float(dt_init)]]), (M, 1)))
        tol = kwargs.get("tol", 1e-3)
        safety, p_order = kwargs.get("safety_factor", 0.9), kwargs.get("p", 2)
        max_dt, min_dt = kwargs.get("max_dt", 40), kwargs.get("min_dt", 1e-4)
        audit = [int(a) - lo for a in audit_models if lo <= int(a) < hi]
        audit_d, n_rec = _engine.index_tensor(np.asarray(sorted(set(audit)), dtype=np.int64)) if audit else (None, 0)
        if audit_d is None:
            audit_d = torch.zeros(1, dtype=torch.int64, device=mesh_d.device)  # empty list: nobody is recorded
        rec_slot = {a: k for k, a in enumerate(sorted(set(audit)))}
        capacity = int(max(4, min(4096, (256 << 20) // (8 * nx * max(n_rec, 1)))))
        rows = {a: [np.asarray(init[lo + a], dtype=np.float64)[None, :]] for a in audit}
        times = {a: [np.array([float(self.t0)])] for a in audit}
        accepted = np.zeros(M, dtype=np.int64)
        used = np.zeros(M, dtype=np.int64)
        done = np.zeros(M, dtype=bool)
        per_launch_cap = 200000
        while True:
            left = per_launch_cap if max_iterations is None else int(min(per_launch_cap, max(0, max_iterations - used.min())))
            snap, tax, counts = _engine.run_adaptive(mesh_d, diff_d, state, ctrl, self.lbc, self.rbc, sidx_d, n_src, st_d,
                                                     sd_d, len_d, self.t0, t_total, tol, 1e-3, safety, p_order, max_dt,
                                                     min_dt, left, capacity, snap_models=audit_d, n_snap_models=n_rec)
            c = counts.cpu().numpy()
            if np.any(c[:, 3]):
                m = int(np.nonzero(c[:, 3])[0][0])
                raise np.linalg.LinAlgError(f"Matrix is singular (model {lo + m}: zero pivot at row {int(c[m, 3]) - 1}).")
            for a in rec_slot:
                n = int(c[a, 0])
                if n:
                    rows[a].append(snap[rec_slot[a], :n].cpu().numpy())
                    times[a].append(tax[rec_slot[a], :n].cpu().numpy())
            accepted += c[:, 0]
            used += c[:, 1]
            done = c[:, 2] != 0
            if done.all() or (max_iterations is not None and used.min() >= max_iterations):
                break
            stalled = (~done) & (c[:, 0] == 0) & (c[:, 1] >= left)
            if stalled.any():
                raise RuntimeError(f"Adaptive time stepping made no progress in {left} iterations for member "
                                   f"{lo + int(np.nonzero(stalled)[0][0])} (the step-doubling error is NaN or above tol "
                                   "at min_dt); the reference loops forever in this case.")
        ctrl_h = ctrl.cpu().numpy()
        self.adaptive = dict(final_state=state.cpu().numpy(), final_time=ctrl_h[:, 0].copy(), final_dt=ctrl_h[:, 1].copy(),
                             accepted=accepted, iterations=used, audit_models=[lo + a for a in audit],
                             taxis=[np.concatenate(times[a]) for a in audit],
                             snapshot=[np.concatenate(rows[a], axis=0) for a in audit])
        self.history.append(f"{M} members x {nx} cells solved adaptively on device (dt_init={dt_init}, tol={tol}): "
                            f"{int(accepted.sum())} accepted / {int(used.sum())} iterations.")
        return self.adaptive


def _dist(process_group):
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def gather_misfit(local, n_models, process_group=None):
    """All-gather the per-member misfits of every rank's contiguous block into one (n_models,)
    tensor (NCCL over NVLink on GPUs; gloo in the CPU tests).  Block sizes may differ by one, so
    blocks are padded to the largest and trimmed after the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(process_group)
    sizes = [shard_bounds(n_models, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=process_group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
