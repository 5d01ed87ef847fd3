"""Device-side operations of the diffusion path: thin Python over the C ABI.

Every function takes / returns torch CUDA tensors (used purely as device buffers) and calls one
entry point of ``libfiberis_b200.so`` on the current CUDA stream.  Nothing here computes on the
host; without a CUDA device or without the built library the calls raise.
"""
from __future__ import annotations

import os

import numpy as np

from . import _cabi
from ._cabi import BC_CODES, check, current_stream, ptr


def _torch():
    import torch
    return torch


def device(dev=None):
    torch = _torch()
    _cabi.require_cuda()
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def to_device(x, dtype=np.float64, dev=None, pinned=False):
    """Host array -> device tensor (one H2D copy)."""
    torch = _torch()
    a = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    t = torch.from_numpy(a)
    if pinned:
        t = t.pin_memory()
    return t.to(device(dev), non_blocking=pinned)


import threading as _threading

_upload_cache = {}  # device -> ([side streams], [pinned staging buffers])
_upload_lock = _threading.Lock()  # one upload at a time per process: the lanes' staging buffers are shared state


def _upload_lanes(d, n, nbytes):
    torch = _torch()
    streams, staging = _upload_cache.setdefault(str(d), ([], []))
    while len(streams) < n:
        streams.append(torch.cuda.Stream(d))
        staging.append(None)
    for i in range(n):
        if staging[i] is None or staging[i].numel() < nbytes[i]:
            staging[i] = torch.empty(nbytes[i], dtype=torch.uint8, pin_memory=True)
    return streams[:n], staging[:n]


def to_device_many(arrays, dtype=np.float64, dev=None):
    """Several LARGE host arrays -> device tensors.  A copy from pageable memory is paced by a single-threaded
    host-side staging memcpy (~10 GB/s, a fifth of the PCIe rate; several threads calling cudaMemcpy do not
    overlap).  Here every array gets a host thread that copies it chunk by chunk into its own cached page-locked
    buffer (NumPy releases the GIL for the copy) and queues one DMA per chunk on its own stream, so the staging
    copies of the arrays run side by side and overlap the transfers (C5: three 80 MB arrays 23.5 -> ~9 ms).
    Returns when the data is on the device.  Small arrays: plain copies."""
    torch = _torch()
    d = device(dev)
    arrs = [np.ascontiguousarray(np.asarray(x, dtype=dtype)) for x in arrays]
    if len(arrs) < 2 or min(a.nbytes for a in arrs) < (8 << 20):
        return [to_device(a, dtype, dev) for a in arrs]
    with _upload_lock:
        return _to_device_many_locked(arrs, d)


def _to_device_many_locked(arrs, d):
    import threading
    torch = _torch()
    main = torch.cuda.current_stream(d)
    # destinations come from the current stream's allocator pool (a side stream would cudaMalloc afresh every call)
    outs = [torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=d) for a in arrs]
    # lanes: every array is cut into a few byte ranges, one host thread + stream + staging buffer per range
    parts = max(1, min(4, 12 // len(arrs)))
    lanes = []
    for i, a in enumerate(arrs):
        step = -(-a.nbytes // parts)
        step = (step + 4095) // 4096 * 4096
        lanes += [(i, o, min(o + step, a.nbytes)) for o in range(0, a.nbytes, step)]
    streams, staging = _upload_lanes(d, len(lanes), [e - o for _, o, e in lanes])
    for s in streams:
        s.wait_stream(main)
    errors = []
    chunk = 8 << 20

    def work(k):
        try:
            i, lo, hi = lanes[k]
            torch.cuda.set_device(d)
            src = arrs[i].reshape(-1).view(np.uint8)[lo:hi]
            pin = staging[k][: hi - lo]
            pin_np = pin.numpy()
            dst = outs[i].view(-1).view(torch.uint8)[lo:hi]
            with torch.cuda.stream(streams[k]):
                for o in range(0, hi - lo, chunk):
                    e = min(o + chunk, hi - lo)
                    np.copyto(pin_np[o:e], src[o:e])
                    dst[o:e].copy_(pin[o:e], non_blocking=True)
            streams[k].synchronize()  # the staging buffer is reused by the next call
        except Exception as ex:  # re-raised on the calling thread
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(lanes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return outs


_pack_cache = {}  # device -> [pinned staging buffer, event of the last copy out of it]
_pack_lock = _threading.Lock()


def to_device_packed(items, dev=None):
    """Several SMALL host arrays -> device tensors through ONE pinned staging buffer and ONE DMA (``items``: list of
    ``(array, numpy dtype)`` or ``None``; the results are views of one device buffer, 256-byte aligned).  A sweep's
    inputs are a dozen arrays of a few KB to a few hundred KB: one pageable ``cudaMemcpy`` each cost 0.05 ms of host
    time apiece in front of the first launch (C3: 0.56 of 20.7 ms end to end).  Arrays above 4 MB take ``to_device``."""
    torch = _torch()
    d = device(dev)
    arrs = [None if it is None else np.ascontiguousarray(np.asarray(it[0], dtype=it[1])) for it in items]
    packed = [a is not None and 0 < a.nbytes <= (4 << 20) for a in arrs]
    offs, total = [], 0
    for a, pk in zip(arrs, packed):
        offs.append(total if pk else None)
        if pk:
            total += (a.nbytes + 255) // 256 * 256
    outs = [None] * len(arrs)
    if total:
        with _pack_lock:
            ent = _pack_cache.get(str(d))
            if ent is None or ent[0].numel() < total:
                ent = [torch.empty(max(total, 2 << 20), dtype=torch.uint8, pin_memory=True), None]
                _pack_cache[str(d)] = ent
            if ent[1] is not None:
                ent[1].synchronize()  # the previous call's DMA has read the staging buffer
            pin_np = ent[0].numpy()
            for a, off, pk in zip(arrs, offs, packed):
                if pk:
                    pin_np[off:off + a.nbytes] = a.reshape(-1).view(np.uint8)
            buf = torch.empty(total, dtype=torch.uint8, device=d)
            buf.copy_(ent[0][:total], non_blocking=True)
            ent[1] = torch.cuda.Event()
            ent[1].record()
        typed = {}  # one typed view of the buffer per element type; the results are slices of those
        for k, (a, off, pk) in enumerate(zip(arrs, offs, packed)):
            if pk:
                base = typed.get(a.dtype)
                if base is None:
                    base = typed[a.dtype] = buf.view(torch.from_numpy(a.reshape(-1)[:0]).dtype)
                lo = off // a.itemsize
                t = base[lo:lo + a.size]
                outs[k] = t if a.ndim == 1 else t.view(a.shape)
    for k, (a, pk) in enumerate(zip(arrs, packed)):
        if a is not None and not pk:
            outs[k] = to_device(a, a.dtype, dev)
    return outs


def index_tensor(idx, dev=None):
    idx = np.atleast_1d(np.asarray(idx)).astype(np.int64)
    if idx.size == 0:
        return None, 0
    return to_device(idx, np.int64, dev), int(idx.size)


def raise_if_singular(flags, what="matrix"):
    """Zero pivot -> the exception scipy.linalg.solve raises in the reference
    (src/fiberis/simulator/solver/PDESolver_IMP.py:28)."""
    bad = flags.nonzero()
    if bad.numel():
        m = int(bad[0, 0])
        row = int(flags[m]) - 1
        raise np.linalg.LinAlgError(f"Matrix is singular ({what} {m}: zero pivot at row {row}).")


def assemble(mesh, diffusivity, M, nx, dt, lbc, rbc, src_idx, n_src, pml_thickness=0.0, sigma_max=2.0):
    """K1: (dl, d, du), each (M, nx).  mesh / diffusivity are (nx,) shared or (M, nx) tensors."""
    torch = _torch()
    lib = _cabi.load()
    out = torch.empty((3, M, nx), dtype=torch.float64, device=mesh.device)
    ms = 0 if mesh.dim() == 1 else nx
    ds = 0 if diffusivity.dim() == 1 else nx
    check(lib.fbr_assemble_f64(ptr(mesh), ms, ptr(diffusivity), ds, M, nx, float(dt), BC_CODES[lbc], BC_CODES[rbc],
                               ptr(src_idx), n_src, float(pml_thickness), float(sigma_max),
                               ptr(out[0]), ptr(out[1]), ptr(out[2]), current_stream()))
    return out[0], out[1], out[2]


def assemble_zonal(mesh, zone_value, zone_of_cell, M, nx, dt, lbc, rbc, src_idx, n_src, pml_thickness=0.0,
                   sigma_max=2.0):
    torch = _torch()
    lib = _cabi.load()
    out = torch.empty((3, M, nx), dtype=torch.float64, device=mesh.device)
    check(lib.fbr_assemble_zonal_f64(ptr(mesh), ptr(zone_value), int(zone_value.shape[1]), ptr(zone_of_cell), M, nx,
                                     float(dt), BC_CODES[lbc], BC_CODES[rbc], ptr(src_idx), n_src,
                                     float(pml_thickness), float(sigma_max), ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                     current_stream()))
    return out[0], out[1], out[2]


def pml_sigma(mesh, pml_thickness, sigma_max):
    torch = _torch()
    out = torch.empty_like(mesh)
    check(_cabi.load().fbr_pml_sigma_f64(ptr(mesh), mesh.numel(), float(pml_thickness), float(sigma_max), ptr(out),
                                         current_stream()))
    return out


def factorize(dl, d, du, check_singular=True, parallel=None):
    """K1b: time-invariant factors (fl, fr, fg); raises LinAlgError on a zero pivot.
    ``parallel``: None = pick by shape; False = always the sequential per-model recurrence (bit-stable
    across batch sizes); True = the cell-parallel variant (needs nx >= 512)."""
    torch = _torch()
    lib = _cabi.load()
    M, nx = d.shape
    out = torch.empty((3, M, nx), dtype=torch.float64, device=d.device)
    flags = torch.zeros(M, dtype=torch.int32, device=d.device)
    work = None
    if parallel is None:
        parallel = nx >= 512 and nx > 8 * M
    if parallel and nx >= 512:  # long systems: cell-parallel factorisation needs scratch
        work = torch.empty(int(lib.fbr_factorize_work_bytes(M, nx)) // 8, dtype=torch.float64, device=d.device)
    check(lib.fbr_factorize_f64(ptr(dl), ptr(d), ptr(du), M, nx, ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(flags),
                                ptr(work), current_stream()))
    if check_singular:
        raise_if_singular(flags, "model")
    return out[0], out[1], out[2]


_SM_COUNT = []


def _sm_count():
    if not _SM_COUNT:
        _SM_COUNT.append(int(_cabi.device_info()["sm_count"]))
    return _SM_COUNT[0]


def costly_members_first(diffusivity_d, M, nx):
    """Order in which a sweep should hand out its members (``fbr_run_opts.member_order``), or None when it cannot
    matter.  The cost of a member in K3 is set by its decay horizon (4 / 5 scan stages, neighbour carry or the warp
    table), which grows with the largest diffusivity of the member: members are ranked by it, largest first, so that
    the last round of the on-demand schedule holds only cheap members.  ``diffusivity_d``: the (M, n_zones) or (M, nx)
    device tensor the run kernel reads.  Worth its two small device kernels only while a CTA integrates a few dozen
    members at most; a hint — no result depends on it."""
    import torch
    if os.environ.get("FBR_NO_MEMBER_ORDER"):  # tuning aid: members in index order
        return None
    if M < 1024 or nx <= 128 or nx > run_max_nx():  # a handful of members / the member-group kernel: no on-demand schedule
        return None
    warps = 1
    while warps * 256 < nx:
        warps *= 2  # team width of the 64-byte-per-thread kernel: 1, 2, 4, 8 or 16 warps; 16 warps per SM are resident
    ctas = _sm_count() * (16 // warps)
    if M <= ctas or M > 64 * ctas:
        return None
    cost = diffusivity_d.amax(dim=1)  # buffer plumbing; the ranking itself is the library's one-CTA counting sort
    order = torch.empty(M, dtype=torch.int32, device=cost.device)
    check(_cabi.load().fbr_rank_members_f64(ptr(cost), M, ptr(order), current_stream()))
    return order


def run(factors, p0, M, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every=1, snap_models=None,
        n_snap_models=None, want_final=False, obs_idx=None, n_obs=0, obs=None, obs_w=None, dtype="float64",
        with_initial_row=True, misfit_out=None, snap_out=None, zonal=None, singular_flags=None, reserve_ctas=0,
        cell_major=False, scan_log2_eps=0.0, sampling=None, scan_mode_out=None, member_order=None, snap_f64=False):
    """K3: persistent on-chip integration.

    ``factors`` = (fl, fr, fg) from ``factorize``; or ``factors=None`` and ``zonal=(mesh, zone_value (M, n_zones),
    zone_of_cell int32 (nx), dt, sigma | None)``: assembly and factorisation happen inside the kernel
    (``fbr_run_zonal_f64``: 128 < nx <= 4096; ``fbr_run_zonal_f32``: 512 <= nx <= 4096).  Singular members raise ``LinAlgError`` here, unless the
    caller passes ``singular_flags`` (from ``zonal_flags(M)``) and checks them later with
    ``raise_if_singular_zonal`` — which keeps the launch asynchronous.

    Returns (snap | None, p_final | None, misfit | None).  ``snap`` has shape
    (n_recorded_models, rows, nx) — the reference's ``snapshot`` layout, written directly by the
    kernel — or, with ``cell_major``, (n_recorded_models, nx, rows): the (cell, time) layout of ``Data2D`` /
    ``pack_result`` (``fbr_run_opts.snap_cell_stride``).  rows = n_steps // snapshot_every (+ 1 leading row holding
    the initial state when ``with_initial_row``).  ``snapshot_every=0`` disables snapshots.

    ``sampling`` (from ``sample_plan``): the observations ``obs`` (n_samples, n_obs) sit on their own time axis.
    ``snap_f64``: fp32 run of at most 64 recorded members of >= 512 cells: the snapshot tensor is float64, the rows are
    widened when the kernel stores them (``fbr_run_opts.snap_f64``).
    ``member_order``: device int32 permutation, the order in which a sweep that records nothing hands out its members
    (``fbr_run_opts.member_order``; see ``costly_members_first``).
    ``reserve_ctas``: CTA slots the persistent grid leaves free.  ``scan_log2_eps``: see ``fbr_run_opts``.
    """
    torch = _torch()
    lib = _cabi.load()
    if zonal is not None:
        assert factors is None
        z_mesh, z_value, z_of_cell, z_dt, z_sigma = zonal
        dev = z_mesh.device
    else:
        fl, fr, fg = factors
        dev = fl.device
    tdt = torch.float64 if dtype == "float64" else torch.float32
    fn = lib.fbr_run_f64 if dtype == "float64" else lib.fbr_run_f32
    n_rows = n_steps // snapshot_every if snapshot_every else 0
    snap = snap_ptr = None
    row_stride = model_stride = cell_stride = 0
    if n_rows > 0:
        ns = M if snap_models is None else n_snap_models
        lead = 1 if with_initial_row else 0
        rows = n_rows + lead
        shape = (ns, nx, rows) if cell_major else (ns, rows, nx)
        sdt = torch.float64 if snap_f64 else tdt  # fbr_run_opts.snap_f64: fp32 run, rows widened at the store
        snap = snap_out if snap_out is not None else torch.empty(shape, dtype=sdt, device=dev)
        assert snap.shape == shape and snap.dtype == sdt
        if lead:
            first = p0 if p0.dim() == 1 else (p0 if snap_models is None else p0[snap_models])
            if cell_major:
                snap[:, :, 0] = first.to(sdt)
            else:
                snap[:, 0, :] = first.to(sdt)
        if cell_major:
            row_stride, model_stride, cell_stride, n_snap_models = 1, rows * nx, rows, ns
            snap_ptr = _cabi.C.c_void_p(snap.data_ptr() + lead * snap.element_size())
        else:
            row_stride, model_stride, n_snap_models = nx, rows * nx, ns
            snap_ptr = _cabi.C.c_void_p(snap.data_ptr() + lead * nx * snap.element_size())
    p_final = torch.empty((M, nx), dtype=tdt, device=dev) if want_final else None
    misfit = misfit_out
    if misfit is None and obs_idx is not None:
        misfit = torch.empty(M, dtype=torch.float64, device=dev)
    p0_stride = 0 if p0.dim() == 1 else nx
    opts = None
    snap_f64 = bool(snap_f64 and dtype == "float32" and n_rows > 0)
    if (reserve_ctas or cell_stride or scan_log2_eps or sampling is not None or scan_mode_out is not None
            or member_order is not None or snap_f64):
        opts = _cabi.C.byref(_cabi.run_opts(reserve_ctas, cell_stride, scan_log2_eps, sampling, scan_mode_out, member_order,
                                            snap_f64))
    tail = (ptr(p0), p0_stride, M, nx, n_steps, BC_CODES[lbc], BC_CODES[rbc],
            ptr(src_idx), n_src, ptr(src_table), max(int(snapshot_every), 1), ptr(snap_models),
            int(n_snap_models or 0), row_stride, model_stride, snap_ptr, ptr(p_final), ptr(obs_idx), int(n_obs),
            ptr(obs), ptr(obs_w), ptr(misfit), opts, current_stream())
    if zonal is not None:
        flags = zonal_flags(M, dev) if singular_flags is None else singular_flags
        zfn = lib.fbr_run_zonal_f64 if dtype == "float64" else lib.fbr_run_zonal_f32
        check(zfn(ptr(z_mesh), ptr(z_value), int(z_value.shape[1]), ptr(z_of_cell), float(z_dt), ptr(z_sigma), ptr(flags), *tail))
        if singular_flags is None:
            raise_if_singular_zonal(flags)
    else:
        check(fn(ptr(fl), ptr(fr), ptr(fg), *tail))
    return snap, p_final, misfit


def sample_plan(taxis, obs_taxis, baseline_index=None, window=None, dev=None):
    """Bracket every observed time between two simulated states (host, NumPy — the arithmetic of ``np.interp``'s
    index search, src/fiberis/analyzer/Data1D/core1D.py:791-800 in the reference) and upload the plan
    ``fbr_run_opts`` wants.  ``taxis``: (n_steps + 1) times of the simulated states (row 0 = initial);
    ``obs_taxis``: (n_samples) non-decreasing, in the same time frame; ``window`` = (first, last) sample range that
    is summed (default: all); ``baseline_index``: sample subtracted from the simulated channel (None: nothing).
    Times outside [taxis[0], taxis[-1]] give NaN misfits when they fall inside the window, as the reference's
    ``interpolate`` (left / right = NaN) does."""
    taxis = np.asarray(taxis, dtype=np.float64)
    t = np.asarray(obs_taxis, dtype=np.float64)
    n_steps, J = len(taxis) - 1, len(t)
    if J == 0 or np.any(np.diff(t) < 0):
        raise ValueError("Observed times must be a non-empty, non-decreasing sequence.")
    row = np.searchsorted(taxis, t, side="right") - 1          # taxis[row] <= t < taxis[row + 1]
    row = np.clip(row, 0, max(n_steps - 1, 0))
    if n_steps > 0:
        w = (t - taxis[row]) / (taxis[row + 1] - taxis[row])
    else:
        w = np.zeros(J)
    at_end = t == taxis[-1]                                     # the last state itself: row n_steps, nothing beyond it
    row = np.where(at_end, n_steps, row)
    w = np.where(at_end, 0.0, w)
    outside = (t < taxis[0]) | (t > taxis[-1])
    row = np.where(outside, -1, row).astype(np.int32)
    w = np.where(outside, 0.0, w)
    done = np.where(outside, 0, row + (w > 0))                  # last state a sample needs
    done = np.maximum.accumulate(done)                          # (outside samples ride along with their neighbours)
    row_first = np.searchsorted(done, np.arange(n_steps + 2), side="left").astype(np.int32)
    first, last = (0, J) if window is None else (int(window[0]), int(window[1]))
    if first < 0:
        first += J
    if last <= 0:
        last += J
    base = -1 if baseline_index is None else int(baseline_index) % J
    if not (0 <= first <= last <= J):
        raise ValueError("Observation window outside the samples.")
    if base >= 0 and base > first:
        raise ValueError("The baseline sample must not lie behind the first sample of the window.")
    return dict(sample_row=to_device(row, np.int32, dev), sample_w=to_device(w, np.float64, dev),
                row_first_sample=to_device(row_first, np.int32, dev), base_sample=base, first_sample=first,
                last_sample=last, host=dict(row=row, w=w, done=done))


def transpose(src):
    """(rows, cols) fp64 device matrix -> its transpose (cols, rows) as a new contiguous tensor (fbr_transpose_f64)."""
    torch = _torch()
    rows, cols = src.shape
    dst = torch.empty((cols, rows), dtype=torch.float64, device=src.device)
    check(_cabi.load().fbr_transpose_f64(ptr(src), rows, cols, ptr(dst), current_stream()))
    return dst


def fp64_peak():
    """Measured fp64 issue peak of the current device, lane-level DFMA per second (fbr_probe_fp64_peak)."""
    torch = _torch()
    out = _cabi.C.c_double(0.0)
    sink = torch.zeros(1, dtype=torch.float64, device=device())
    check(_cabi.load().fbr_probe_fp64_peak(_cabi.C.byref(out), ptr(sink), current_stream()))
    return float(out.value)


def zonal_flags(M, dev=None):
    """Per-member flag array for fbr_run_zonal_f64 (the kernel takes the minimum with 1 + zero-pivot row)."""
    return _torch().full((M,), 2**31 - 1, dtype=_torch().int32, device=device(dev))


def raise_if_singular_zonal(flags):
    """Flags of fbr_run_zonal_f64 (INT32_MAX = fine) -> the exception the reference's scipy.linalg.solve raises."""
    if flags.numel() <= 4096:  # a handful of models: one small copy instead of four launches and a sync (C1: -0.05 ms)
        f = flags.cpu().numpy()
        bad = np.flatnonzero(f != 2**31 - 1)
        bad = bad[f[bad] != 0]
        if bad.size:
            m = int(bad[0])
            raise np.linalg.LinAlgError(f"Matrix is singular (system {m}: zero pivot at row {int(f[m]) - 1}).")
        return
    raise_if_singular(flags * (flags != 2**31 - 1), "system")


def run_long_max_nx() -> int:
    return int(_cabi.load().fbr_run_long_max_nx())


def run_long(factors, p0, M, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every=1, want_final=False,
             with_initial_row=True):
    """K3L: cooperative persistent integration of long meshes.  p0: (M, nx).  Returns
    (snap (M, rows, nx) | None, p_final | None) with the same snapshot layout as ``run``."""
    torch = _torch()
    lib = _cabi.load()
    fl, fr, fg = factors
    dev = fl.device
    n_rows = n_steps // snapshot_every if snapshot_every else 0
    snap = snap_ptr = None
    row_stride = model_stride = 0
    lead = 1 if with_initial_row else 0
    if n_rows > 0:
        snap = torch.empty((M, n_rows + lead, nx), dtype=torch.float64, device=dev)
        if lead:
            snap[:, 0, :] = p0
        row_stride, model_stride = nx, (n_rows + lead) * nx
        snap_ptr = _cabi.C.c_void_p(snap.data_ptr() + lead * nx * 8)
    p_final = torch.empty((M, nx), dtype=torch.float64, device=dev) if want_final else None
    work = torch.empty(int(lib.fbr_run_long_work_bytes(nx)) // 8, dtype=torch.float64, device=dev)
    check(lib.fbr_run_long_f64(ptr(fl), ptr(fr), ptr(fg), ptr(p0), M, nx, n_steps, BC_CODES[lbc], BC_CODES[rbc],
                               ptr(src_idx), n_src, ptr(src_table), max(int(snapshot_every), 1), row_stride,
                               model_stride, snap_ptr, ptr(p_final), ptr(work), current_stream()))
    return snap, p_final


def run_stream(factors, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, snapshot_every=1, want_final=False,
               with_initial_row=True, kernel="pass", horizon=None, tile_warps=None):
    """Exact integration of one mesh of any length.  factors / p0: (1, nx).  Returns (snap (1, rows, nx) | None,
    p_final | None).  ``kernel='pass'``: K3P, one pass over HBM per step (40 B per cell-update, fbr_run_pass_tiled_f64);
    ``'levels'``: K3S, the level hierarchy of round 1 (88 B per cell-update, fbr_run_stream_f64).
    K3P's tile width follows the decay horizon of the factors (``horizon``: what decay_horizon returned, measured here
    with a small cap when not given and the run is long enough to pay for the 16-byte read): narrow tiles when the
    coupling ends at the neighbouring tile, wide ones on stiff meshes; ``tile_warps`` = 2 | 4 | 8 overrides."""
    torch = _torch()
    lib = _cabi.load()
    fl, fr, fg = factors
    if kernel == "pass":
        if tile_warps is None:
            tile_warps = pass_tile_warps(factors, n_steps, horizon)
        work_fn = lib.fbr_run_pass_work_bytes

        def run_fn(*args):
            return lib.fbr_run_pass_tiled_f64(*args[:-2], int(tile_warps), *args[-2:])
    else:
        work_fn, run_fn = lib.fbr_run_stream_work_bytes, lib.fbr_run_stream_f64
    dev = fl.device
    n_rows = n_steps // snapshot_every if snapshot_every else 0
    snap = snap_ptr = None
    lead = 1 if with_initial_row else 0
    if n_rows > 0:
        snap = torch.empty((1, n_rows + lead, nx), dtype=torch.float64, device=dev)
        if lead:
            snap[:, 0, :] = p0
        snap_ptr = _cabi.C.c_void_p(snap.data_ptr() + lead * nx * 8)
    p_final = torch.empty((1, nx), dtype=torch.float64, device=dev) if want_final else None
    work = torch.empty(int(work_fn(nx)) // 8, dtype=torch.float64, device=dev)
    check(run_fn(ptr(fl), ptr(fr), ptr(fg), ptr(p0), nx, n_steps, BC_CODES[lbc], BC_CODES[rbc],
                                 ptr(src_idx), n_src, ptr(src_table), max(int(snapshot_every), 1), nx, snap_ptr,
                                 ptr(p_final), ptr(work), current_stream()))
    return snap, p_final


def pass_tile_warps(factors, n_steps, horizon=None):
    """Tile width (warps of 256 cells) K3P should use for this factorisation: 2 when the decay horizon (``horizon``, or
    measured here — one 16-byte read, only for runs long enough to pay for it) is below ~100 cells, so that the coupling
    between tiles ends at the neighbouring tile; 8 on stiff meshes, whose carries fold over tens of tiles."""
    if horizon is None and n_steps >= 16:
        horizon = decay_horizon(factors, cap=512)
    return 2 if horizon is not None and max(horizon) <= 96 else 8


def decay_horizon(factors, log2_eps=-80.0, cap=2048):
    """(forward, backward) decay horizons of a factorisation (one model): how many consecutive
    multipliers it takes for their product to fall below 2**log2_eps.  One 16-byte D2H read."""
    torch = _torch()
    lib = _cabi.load()
    fl, _, fg = factors
    nx = fl.shape[-1]
    hor = torch.zeros(2, dtype=torch.int64, device=fl.device)
    work = torch.zeros(4, dtype=torch.float64, device=fl.device)
    check(lib.fbr_decay_horizon_f64(ptr(fl), ptr(fg), nx, float(log2_eps), int(cap), ptr(hor), ptr(work),
                                    current_stream()))
    hf, hb = (int(v) for v in hor.tolist())
    return hf, hb


def window_plan(nx, horizon, snapshot_every=0, dtype="float64", cap=2048):
    """What K3W would use for this mesh: dict(warps, steps_per_block, owned, windows), or None when
    the horizons are too long for windows to pay.  ``cap``: the cap decay_horizon ran with — a horizon of
    cap + 1 means "not established within cap cells" and is never a usable halo (the fp32 mode's 8192-cell
    windows would otherwise accept it)."""
    if max(int(horizon[0]), int(horizon[1])) > cap:
        return None
    lib = _cabi.load()
    out = (_cabi.C.c_int64 * 4)()
    fn = lib.fbr_run_window_plan if dtype == "float64" else lib.fbr_run_window_plan_f32
    rc = fn(int(nx), int(horizon[0]), int(horizon[1]), int(snapshot_every), out)
    if rc == _cabi.FBR_ERR_UNSUPPORTED:
        return None
    check(rc)
    return dict(warps=int(out[0]), steps_per_block=int(out[1]), owned=int(out[2]), windows=int(out[3]))


def run_window(factors, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, horizon, snapshot_every=1,
               want_final=False, with_initial_row=True, dtype="float64"):
    """K3W: time-tiled overlapping windows for one long mesh.  factors / p0: (1, nx), fp64.  Returns
    (snap (1, rows, nx) | None, p_final | None) — float tensors in the fp32 mode (``dtype='float32'``:
    fbr_run_window_f32, 16 cells per thread on the fp32 pipe)."""
    torch = _torch()
    lib = _cabi.load()
    fl, fr, fg = factors
    dev = fl.device
    f32 = dtype == "float32"
    tdt, esz = (torch.float32, 4) if f32 else (torch.float64, 8)
    n_rows = n_steps // snapshot_every if snapshot_every else 0
    snap = snap_ptr = None
    lead = 1 if with_initial_row else 0
    if n_rows > 0:
        snap = torch.empty((1, n_rows + lead, nx), dtype=tdt, device=dev)
        if lead:
            snap[:, 0, :] = p0
        snap_ptr = _cabi.C.c_void_p(snap.data_ptr() + lead * nx * esz)
    p_final = torch.empty((1, nx), dtype=tdt, device=dev) if want_final else None
    work_bytes = int((lib.fbr_run_window_f32_work_bytes if f32 else lib.fbr_run_window_work_bytes)(nx))
    work = torch.empty(work_bytes // 8 + 1, dtype=torch.float64, device=dev)
    fn = lib.fbr_run_window_f32 if f32 else lib.fbr_run_window_f64
    check(fn(ptr(fl), ptr(fr), ptr(fg), ptr(p0), nx, n_steps, BC_CODES[lbc], BC_CODES[rbc],
             ptr(src_idx), n_src, ptr(src_table), int(snapshot_every) if snap is not None else 0, nx, snap_ptr,
             ptr(p_final), int(horizon[0]), int(horizon[1]), ptr(work), current_stream()))
    return snap, p_final


def run_window_to_host(factors, p0, nx, n_steps, lbc, rbc, src_idx, n_src, src_table, horizon, snapshot_every=1):
    """K3W with the snapshot rows streamed to the host while later time blocks are still being
    integrated: one library call per snapshot interval, an event after each, and the
    row's device-to-host copy on a second stream into pooled pinned memory.  Returns the NumPy array
    (rows + 1, nx), row 0 = initial state; None when no row is due."""
    torch = _torch()
    lib = _cabi.load()
    fl, fr, fg = factors
    dev = fl.device
    n_rows = n_steps // snapshot_every if snapshot_every else 0
    if n_rows == 0:
        return None
    import weakref
    snap = torch.empty((n_rows + 1, nx), dtype=torch.float64, device=dev)
    snap[0] = p0.reshape(-1)
    entry = _pool.take(snap.numel() * 8)
    try:
        return _run_window_to_host_reserved(entry, snap, factors, nx, n_rows, lbc, rbc, src_idx, n_src, src_table, horizon,
                                            snapshot_every)
    except BaseException:
        _pool.release(entry)
        raise


def _run_window_to_host_reserved(entry, snap, factors, nx, n_rows, lbc, rbc, src_idx, n_src, src_table, horizon, snapshot_every):
    import weakref
    torch = _torch()
    lib = _cabi.load()
    fl, fr, fg = factors
    dev = fl.device
    host = entry[0][: snap.numel() * 8].view(torch.float64).view(snap.shape)
    work = torch.empty(int(lib.fbr_run_window_work_bytes(nx)) // 8, dtype=torch.float64, device=dev)
    main, side = torch.cuda.current_stream(), _copy_stream(dev)
    ev = torch.cuda.Event()
    ev.record(main)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        host[0].copy_(snap[0], non_blocking=True)
    for r in range(n_rows):
        check(lib.fbr_run_window_f64(ptr(fl), ptr(fr), ptr(fg), _cabi.C.c_void_p(snap[r].data_ptr()), nx, snapshot_every,
                                     BC_CODES[lbc], BC_CODES[rbc], ptr(src_idx), n_src,
                                     _cabi.C.c_void_p(src_table.data_ptr() + r * snapshot_every * n_src * 8) if n_src else None,
                                     int(snapshot_every), nx, None, _cabi.C.c_void_p(snap[r + 1].data_ptr()), int(horizon[0]),
                                     int(horizon[1]), ptr(work), current_stream()))
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            host[r + 1].copy_(snap[r + 1], non_blocking=True)
    side.synchronize()
    main.wait_stream(side)
    arr = host.numpy()
    entry[1] = weakref.ref(arr)
    return arr


_side_streams = {}


def _copy_stream(dev):
    torch = _torch()
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def adaptive_max_nx() -> int:
    return int(_cabi.load().fbr_run_adaptive_max_nx())


def run_adaptive(mesh, diffusivity, state, ctrl, lbc, rbc, src_idx, n_src, src_taxis, src_data, src_len, t0, t_total,
                 tol, adj_tol, safety, p_order, max_dt, min_dt, max_iters, capacity, snap_models=None, n_snap_models=None):
    """One launch of the device-side adaptive loop for M models (state (M, nx), ctrl (M, 2) updated in
    place).  ``snap_models`` (device int64 list, None = all): the models whose accepted states are
    recorded.  Returns (snap (recorded, capacity, nx), taxis (recorded, capacity), counts (M, 4) int64)
    on the device."""
    torch = _torch()
    lib = _cabi.load()
    M, nx = state.shape
    dev = state.device
    n_rec = M if snap_models is None else int(n_snap_models)
    snap = torch.empty((max(n_rec, 1), capacity, nx), dtype=torch.float64, device=dev)
    taxis = torch.empty((max(n_rec, 1), capacity), dtype=torch.float64, device=dev)
    counts = torch.zeros((M, 4), dtype=torch.int64, device=dev)
    if snap_models is None and n_rec == 0:
        raise ValueError("no models")
    rec_ptr = ptr(snap_models) if (snap_models is not None and n_rec > 0) else None
    if snap_models is not None and n_rec == 0:  # nothing recorded: an empty list needs a non-NULL pointer
        rec_ptr = ptr(torch.zeros(1, dtype=torch.int64, device=dev))
    ms = 0 if mesh.dim() == 1 else nx
    ds = 0 if diffusivity.dim() == 1 else nx
    width = int(src_taxis.shape[1]) if n_src else 0
    check(lib.fbr_run_adaptive_f64(ptr(mesh), ms, ptr(diffusivity), ds, ptr(state), M, nx, BC_CODES[lbc], BC_CODES[rbc],
                                   ptr(src_idx), n_src, ptr(src_taxis), ptr(src_data), ptr(src_len), width, float(t0),
                                   float(t_total), ptr(ctrl), float(tol), float(adj_tol), float(safety), float(p_order),
                                   float(max_dt), float(min_dt), int(max_iters), int(capacity), ptr(snap), ptr(taxis),
                                   rec_ptr, n_rec, ptr(counts), current_stream()))
    return snap, taxis, counts


class _PinnedPool:
    """Page-locked host buffers for device->host result copies.  cudaHostAlloc costs ~0.6 ms/MB, so
    buffers are kept and handed out again once the NumPy array that aliased them has been
    garbage-collected (tracked with a weak reference); a live result is never overwritten."""

    def __init__(self):
        self.entries = []  # [pinned uint8 tensor, weakref to the last ndarray handed out]
        self.lock = _threading.Lock()

    @staticmethod
    def _busy():  # stands in for the weak reference between take() and the caller's entry[1] = weakref.ref(array)
        return True

    def take(self, nbytes):
        """An entry [buffer, liveness] RESERVED for the caller: the reservation happens under the lock, so two threads
        never receive the same buffer.  The caller replaces entry[1] by a weak reference to the array it hands out, or
        calls release(entry) when the copy failed."""
        with self.lock:
            entry = self._take(nbytes)
            entry[1] = self._busy
            return entry

    def release(self, entry):
        entry[1] = lambda: None

    def _take(self, nbytes):
        torch = _torch()
        free = [e for e in self.entries if e[1]() is None]
        fits = [e for e in free if nbytes <= e[0].numel() <= 2 * nbytes + 4096]
        if fits:
            return min(fits, key=lambda e: e[0].numel())
        drop = free[: max(0, len(self.entries) - 7)]  # bound the pool: drop unused buffers first
        if drop:  # by identity: list.remove would compare the tensors inside the entries element by element
            self.entries = [x for x in self.entries if not any(x is e for e in drop)]
        entry = [torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True), lambda: None]
        self.entries.append(entry)
        return entry


_pool = _PinnedPool()


def as_f64(t):
    """Results are handed out as float64 arrays (reference convention); fp32-mode buffers are widened on
    the device — a host-side ``astype`` of a few hundred MB costs more than the run."""
    torch = _torch()
    return t if t.dtype == torch.float64 else t.to(torch.float64)


def to_host(t, sync=True):
    """Device tensor -> NumPy array through pooled pinned memory (one DMA, no extra host copy).
    ``sync=False`` only enqueues the copy on the current stream: the caller synchronises that stream
    before touching the array."""
    torch = _torch()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20) and sync:
        return t.cpu().numpy()
    import weakref
    entry = _pool.take(max(nbytes, 1))
    try:
        host = entry[0][:nbytes].view(t.dtype).view(t.shape)
        host.copy_(t, non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        arr = host.numpy()
    except BaseException:
        _pool.release(entry)
        raise
    entry[1] = weakref.ref(arr)
    return arr


def run_max_nx() -> int:
    return int(_cabi.load().fbr_run_max_nx())


def rhs(p, lbc, rbc, src_idx, n_src, src_val, out=None):
    torch = _torch()
    lib = _cabi.load()
    M, nx = p.shape
    b = torch.empty_like(p) if out is None else out
    check(lib.fbr_rhs_f64(ptr(p), M, nx, BC_CODES[lbc], BC_CODES[rbc], ptr(src_idx), n_src, ptr(src_val), ptr(b),
                          current_stream()))
    return b


def solve_tridiag(dl, d, du, b, variant=_cabi.SOLVER_AUTO, check_singular=True):
    """K2: bare batched solve (dl, d, du, b) -> x, all (M, nx)."""
    torch = _torch()
    lib = _cabi.load()
    M, nx = d.shape
    if variant == _cabi.SOLVER_AUTO and nx > 4096 and M <= 4:
        # a few LONG systems: the Thomas walk is one thread per system.  Parallel along the mesh instead: the tile-parallel
        # factorisation (K1b) and ONE pass of K3P with the right-hand side as "state" and nothing overridden
        # (fbr_run_pass_f64, n_steps = 1) — 1e7 unknowns in ~0.5 ms
        try:
            rows = []
            for m in range(M):
                fac = factorize(dl[m:m + 1], d[m:m + 1], du[m:m + 1], check_singular=check_singular, parallel=True)
                rows.append(run_stream(fac, b[m:m + 1], nx, 1, "None", "None", None, 0, None, snapshot_every=0,
                                       want_final=True, kernel="pass")[1])
            return torch.cat(rows, dim=0)
        except np.linalg.LinAlgError:  # zero pivot without row interchanges: the pivoted variant decides
            return solve_tridiag(dl, d, du, b, _cabi.SOLVER_PIVOT, True)
    x = torch.empty_like(b)
    flags = torch.zeros(M, dtype=torch.int32, device=d.device)
    work = None
    if variant == _cabi.SOLVER_THOMAS or (variant == _cabi.SOLVER_AUTO and nx > 4096):
        work = torch.empty((M, nx), dtype=torch.float64, device=d.device)
    elif variant == _cabi.SOLVER_PIVOT:
        work = torch.empty((4, M, nx), dtype=torch.float64, device=d.device)
    check(lib.fbr_solve_tridiag_f64(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(x), M, nx, int(variant), ptr(work),
                                    ptr(flags), current_stream()))
    if check_singular and bool(flags.any()):
        if variant == _cabi.SOLVER_AUTO and nx <= 4096:
            # the register-resident variant also flags systems whose leading minors leave the fp64 range between two
            # rescalings (rows scaled by ~1e77 and more): the elimination without products of pivots decides
            return solve_tridiag(dl, d, du, b, _cabi.SOLVER_THOMAS, True)
        if variant != _cabi.SOLVER_PIVOT:
            # a zero pivot WITHOUT row interchanges: the reference's LAPACK solve pivots (PDESolver_IMP.py:27-28) and
            # only a truly singular matrix raises there — same here (DGTSV algorithm, fbr_solve_tridiag_f64 PIVOT)
            return solve_tridiag(dl, d, du, b, _cabi.SOLVER_PIVOT, True)
        raise_if_singular(flags, "system")
    return x


def solve_dense(A, b):
    """General dense solve with partial pivoting (fbr_solve_dense_f64; API completeness of ``solver_implicit``)."""
    torch = _torch()
    n = A.shape[0]
    Ab = torch.cat([A, b.reshape(n, 1)], dim=1).contiguous()
    x = torch.empty(n, dtype=torch.float64, device=A.device)
    flag = torch.zeros(1, dtype=torch.int32, device=A.device)
    check(_cabi.load().fbr_solve_dense_f64(ptr(Ab), n, ptr(x), ptr(flag), current_stream()))
    if int(flag):
        raise np.linalg.LinAlgError("Matrix is singular.")
    return x


def jacobi_dense(A, b, tol=1e-10, max_iter=1000):
    """Jacobi iteration on a general dense system (fbr_jacobi_dense_f64; API completeness of ``solver_explicit``).
    Returns (x, iterations used | -max_iter)."""
    torch = _torch()
    n = A.shape[0]
    x = torch.empty(n, dtype=torch.float64, device=A.device)
    work = torch.empty(n, dtype=torch.float64, device=A.device)
    iters = torch.zeros(1, dtype=torch.int32, device=A.device)
    check(_cabi.load().fbr_jacobi_dense_f64(ptr(A), ptr(b), ptr(x), n, float(tol), int(max_iter), ptr(iters), ptr(work),
                                            current_stream()))
    return x, int(iters)


def step(dl, d, du, p, lbc, rbc, src_idx, n_src, src_val, variant=_cabi.SOLVER_AUTO, out=None, check_singular=True):
    """One implicit step of M models: p_next = A^-1 rhs(p) (fbr_step_f64; overrides fused into the
    register-resident solve up to 4096 cells).  src_val: (n_src) shared or (M, n_src) per model."""
    torch = _torch()
    lib = _cabi.load()
    M, nx = d.shape
    x = torch.empty_like(p) if out is None else out
    flags = torch.zeros(M, dtype=torch.int32, device=d.device)
    work = None
    if variant == _cabi.SOLVER_THOMAS or (variant == _cabi.SOLVER_AUTO and nx > 4096):
        work = torch.empty((M, nx), dtype=torch.float64, device=d.device)
    stride = 0 if (src_val is None or src_val.dim() == 1) else int(src_val.shape[1])
    check(lib.fbr_step_f64(ptr(dl), ptr(d), ptr(du), ptr(p), M, nx, BC_CODES[lbc], BC_CODES[rbc], ptr(src_idx), n_src,
                           ptr(src_val), stride, ptr(x), int(variant), ptr(work), ptr(flags), current_stream()))
    if check_singular:
        if variant == _cabi.SOLVER_AUTO and nx <= 4096 and stride == 0 and x is not p and bool(flags.any()):
            return step(dl, d, du, p, lbc, rbc, src_idx, n_src, src_val, _cabi.SOLVER_THOMAS, out, True)  # see solve_tridiag
        raise_if_singular(flags, "system")
    return x


def misfit_of_rows(snap, obs_idx, obs, obs_w=None, first_row=1):
    """Misfit of stored rows (n_models, n_rows, nx) against obs (n_rows, n_obs): fbr_misfit_f64."""
    torch = _torch()
    lib = _cabi.load()
    R, n_rows, nx = snap.shape
    out = torch.empty(R, dtype=torch.float64, device=snap.device)
    check(lib.fbr_misfit_f64(ptr(snap), R, n_rows, nx, snap.stride(1), snap.stride(0), ptr(obs_idx), int(obs.shape[1]),
                             ptr(obs), ptr(obs_w), int(first_row), ptr(out), current_stream()))
    return out


def error_norms(u1, u2):
    torch = _torch()
    lib = _cabi.load()
    M, nx = u1.shape
    err = torch.empty(M, dtype=torch.float64, device=u1.device)
    check(lib.fbr_error_norms_f64(ptr(u1), ptr(u2), M, nx, ptr(err), current_stream()))
    return err


def jacobi(dl, d, du, b, tol=1e-10, max_iter=1000):
    torch = _torch()
    lib = _cabi.load()
    M, nx = d.shape
    x = torch.empty_like(b)
    iters = torch.zeros(M, dtype=torch.int32, device=d.device)
    check(lib.fbr_jacobi_f64(ptr(dl), ptr(d), ptr(du), ptr(b), ptr(x), M, nx, float(tol), int(max_iter), ptr(iters),
                             current_stream()))
    return x, iters
