"""CPU-side behaviour of the reference-shaped API: validation, containers, controller scalars,
sharding arithmetic, the C-ABI export list, and the absence of any CPU fallback."""
import datetime
import os
import re

import numpy as np
import numpy.testing as npt
import pytest

from fiberis_b200 import _cabi
from fiberis_b200.analyzer.Data1D.core1D import Data1D
from fiberis_b200.analyzer.Data2D.core2D import Data2D
from fiberis_b200.simulator.core import pds
from fiberis_b200.simulator.core.bcs import BoundaryCondition
from fiberis_b200.simulator.core.ensemble import PDS1D_Ensemble, shard_bounds
from fiberis_b200.simulator.optimizer import tso
from fiberis_b200.utils import mesh_utils
from fiberis_b200.utils.history_utils import InfoManagementSystem
from tests import _golden as G
from tests import _models as Mo

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


# ---- C ABI -------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fiberis_b200.h")).read()
    declared = set(re.findall(r"\b(fbr_[a-z0-9_]+)\s*\(", header))
    assert declared >= {"fbr_assemble_f64", "fbr_factorize_f64", "fbr_run_f64", "fbr_run_f32",
                        "fbr_solve_tridiag_f64", "fbr_error_norms_f64", "fbr_jacobi_f64"}
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.fbr_version() == 201 and lib.fbr_run_max_nx() >= 4096


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = Mo.model(np.linspace(0, 4, 5), np.ones(5), np.zeros(5), "Dirichlet", "Dirichlet", [2],
                 [(np.array([0.0, 5.0]), np.array([0.0, 1.0]))], 0.0, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.solve(dt=0.5, t_total=2.0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fiberis_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


# ---- Data1D / Data2D ---------------------------------------------------------------------------
def test_data1d_interp_semantics():  # /root/reference/tests/test_core1D_extended.py:186-227
    d = Data1D(data=np.array([0.0, 10.0, 20.0, 30.0, 40.0]), taxis=np.arange(5.0),
               start_time=datetime.datetime(2023, 1, 1))
    assert d.get_value_by_time(1.5) == 15.0 and d.get_value_by_time(2) == 20.0
    assert d.get_value_by_time(datetime.datetime(2023, 1, 1, 0, 0, 1)) == 10.0
    assert d.get_value_by_time(-100.0) == 0.0 and d.get_value_by_time(100.0) == 40.0
    one = Data1D(data=np.array([5.0]), taxis=np.array([2.0]))
    assert one.get_value_by_time(2.0) == 5.0 and one.get_value_by_time(100.0) == 5.0
    with pytest.raises(ValueError):
        Data1D().get_value_by_time(1.0)
    with pytest.raises(ValueError):
        Data1D(data=np.array([1.0]), taxis=np.array([0.0])).get_value_by_time(datetime.datetime(2023, 1, 1))
    with pytest.raises(TypeError):
        d.get_value_by_time("nope")


def test_data1d_interp_matches_reference_fixture():
    z = G.load("data1d_interp_96")
    d = Data1D(data=z["data"], taxis=z["taxis"])
    npt.assert_array_equal([d.get_value_by_time(float(t)) for t in z["query"]], z["value"])


def test_source_table_is_bit_equal_to_scalar_sampling():
    z = G.load("data1d_interp_96")
    p = pds.PDS1D_MultiSource()
    p.source = [Data1D(data=z["data"], taxis=z["taxis"]), Data1D(data=z["data"][::-1].copy(), taxis=z["taxis"])]
    p.t0 = 3.25
    starts = [3.25]
    for _ in range(500):
        starts.append(starts[-1] + 0.3)
    table = p._source_table(starts)
    for k, s in enumerate(p.source):
        npt.assert_array_equal(table[:, k], [s.get_value_by_time(t - p.t0) for t in starts])


def test_data1d_npz_roundtrip(tmp_path):
    d = Data1D(data=np.array([1.0, 2.0]), taxis=np.array([0.0, 1.0]), start_time=datetime.datetime(2024, 5, 1))
    d.savez(str(tmp_path / "g"))
    e = Data1D()
    e.load_npz(str(tmp_path / "g.npz"))
    npt.assert_array_equal(e.data, d.data)
    assert e.start_time == d.start_time and e.name == "g.npz"
    with pytest.raises(FileNotFoundError):
        e.load_npz("/nonexistent/file.npz")
    np.savez(tmp_path / "bad.npz", data=np.array([1.0]))
    with pytest.raises(KeyError):
        e.load_npz(str(tmp_path / "bad.npz"))


def test_data2d_container(tmp_path):
    data = np.arange(12.0).reshape(3, 4)
    d = Data2D(data=data, taxis=np.arange(4.0), daxis=np.array([10.0, 20.0, 30.0]),
               start_time=datetime.datetime(2024, 1, 1), name="x")
    d.savez(str(tmp_path / "r"))
    e = Data2D()
    e.load_npz(str(tmp_path / "r"))
    npt.assert_array_equal(e.data, data)
    assert e.start_time == d.start_time and e.name == "r.npz"
    npt.assert_array_equal(e.get_value_by_depth(21.0), data[1])
    assert e.get_value_by_depth(99.0) is None
    trace, t = e.get_value_by_time(2.2)
    npt.assert_array_equal(trace, data[:, 2])
    assert t == 2.0 and e.get_max_daxis() == 30.0
    with pytest.raises(TypeError):
        e.set_data(np.zeros(3))
    with pytest.raises(ValueError):
        e.set_taxis(np.arange(5.0))
    with pytest.raises(ValueError):
        Data2D().savez(str(tmp_path / "none"))
    np.savez(tmp_path / "mism.npz", data=data, taxis=np.arange(3.0), daxis=np.arange(3.0), start_time="2024-01-01")
    with pytest.raises(ValueError):
        Data2D().load_npz(str(tmp_path / "mism.npz"))


def test_history_system():
    h = InfoManagementSystem()
    h.add_record("a")
    h.add_record(5, level="warning")
    assert [r["level"] for r in h.records] == ["INFO", "WARNING"] and h.records[1]["description"] == "5"
    assert len(h.get_records(lambda r: r["level"] == "INFO")) == 1
    assert "2 records" in str(h)
    h.clear_records()
    assert h.records == [] and str(h) == "No records in history."


def test_mesh_utils():
    x = np.linspace(0.0, 10.0, 11)
    fine = mesh_utils.refine_mesh(x, (2.0, 4.0), 4)
    assert len(fine) == 11 - 3 + 13 and np.all(np.diff(fine) > 0) and fine[0] == 0.0 and fine[-1] == 10.0
    assert mesh_utils.locate(x, 3.4) == (3, 3.0)
    with pytest.raises(ValueError):
        mesh_utils.refine_mesh(list(x), (2, 4), 2)
    with pytest.raises(ValueError):
        mesh_utils.refine_mesh(x, (4, 2), 2)
    with pytest.raises(ValueError):
        mesh_utils.locate(np.array([]), 1.0)


# ---- BoundaryCondition (host helper) -------------------------------------------------------------
def test_boundary_condition_rows():  # /root/reference/tests/test_simulator_core.py:53-112
    base = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]])
    for kind, idx, value, row, rhs in [("Dirichlet", 0, 5.0, [1, 0, 0], [5, 2, 3]),
                                       ("Neumann", 0, 0.0, [1, -1, 0], [0, 2, 3]),
                                       ("Neumann", 2, 0.0, [0, -1, 1], [1, 2, 0])]:
        A, b = base.copy(), np.array([1.0, 2.0, 3.0])
        bc = BoundaryCondition(type=kind, value=value, idx=idx)
        bc.set_matrix(A, b)
        bc.apply()
        npt.assert_array_equal(A[idx], row)
        npt.assert_array_equal(b, rhs)
    bc = BoundaryCondition()
    for A, b, msg in [(np.zeros((2, 3)), np.zeros(2), "not square"), (np.eye(2), np.zeros((2, 2)), "not a vector"),
                      (np.eye(3), np.zeros(2), "does not match")]:
        with pytest.raises(ValueError, match=msg):
            bc.set_matrix(A, b)
    bad = BoundaryCondition(type="foo", idx=0)
    bad.set_matrix(np.eye(2), np.zeros(2))
    with pytest.raises(ValueError, match="Unknown BC type"):
        bad.apply()
    pml = BoundaryCondition(type="PML", idx=0)
    A = np.eye(3)
    pml.set_matrix(A, np.zeros(3))
    assert pml.apply() is None
    npt.assert_array_equal(A, np.eye(3))


# ---- model configuration ---------------------------------------------------------------------------
def test_setters_and_checks(capsys):  # /root/reference/tests/test_simulator_core.py:118-174
    p = pds.PDS1D_SingleSource()
    assert p.self_check() is False
    assert p.solve(dt=0.1, t_total=1.0) is None  # prints, returns None, no exception
    p.set_mesh(np.linspace(0.0, 4.0, 5))
    p.set_diffusivity(2.0)
    npt.assert_array_equal(p.diffusivity, np.full(5, 2.0))
    p.set_diffusivity([1.0, 2.0, 3.0, 4.0, 5.0])
    npt.assert_array_equal(p.diffusivity, [1.0, 2.0, 3.0, 4.0, 5.0])
    with pytest.raises(ValueError, match="single scalar value or an array"):
        p.set_diffusivity([1.0, 2.0])
    q = pds.PDS1D_SingleSource()
    q.set_mesh(np.array([1.0]))
    assert q._check_mesh() == (False, "Mesh must have at least 2 points.")
    q.set_bcs("Dirichlet", "Bogus")
    ok, msg = q._check_bc()
    assert ok is False and "Invalid right BC" in msg
    good = Mo.model(np.linspace(0, 4, 5), np.ones(5), np.zeros(5), "Dirichlet", "Dirichlet", [2],
                    [(np.array([0.0, 5.0]), np.array([0.0, 1.0]))], 0.0, False)
    assert good.self_check() is True and good.self_check("does_not_exist") is False
    assert good.self_check(["mesh", "bc"]) is True
    assert len(good.history) == 7 and good.history[0] == "Mesh is properly set."
    good.snapshot, good.taxis = "x", "y"
    good.reset()
    assert good.snapshot is None and good.taxis is None and good.history == []
    capsys.readouterr()


def test_multi_source_configuration():
    srcs = [(np.array([0.0, 5.0]), np.array([0.0, 1.0]))] * 2
    p = Mo.model(np.linspace(0, 6, 7), np.ones(7), np.zeros(7), "Dirichlet", "Dirichlet", [2, 4], srcs, 0.0, True)
    assert isinstance(p.sourceidx, np.ndarray) and p.self_check() is True
    npt.assert_array_equal(p.sourceidx, [2, 4])
    assert isinstance(p, pds.PDS1D_SingleSource)
    q = pds.PDS1D_MultiSource()
    q.set_source(Mo.sources_of(srcs)[0])
    assert q._check_source() == (False, "Source term must be a list.")
    p.source = p.source[:1]
    assert p._check_sourceidx() == (False, "Source index and source term length do not match.")


def test_fixed_dt_requires_dt():
    p = Mo.model(np.linspace(0, 4, 5), np.ones(5), np.zeros(5), "Dirichlet", "Dirichlet", [2],
                 [(np.array([0.0, 5.0]), np.array([0.0, 1.0]))], 0.0, False)
    with pytest.raises(KeyError):
        p.solve(t_total=1.0)
    with pytest.raises(KeyError):
        p.solve(optimizer=True, t_total=1.0)


# ---- controller scalars ------------------------------------------------------------------------------
def test_adjust_dt_goldens():  # /root/reference/tests/test_simulator_optimizer.py:28-53
    assert tso.adjust_dt(1.0, 1e-3) == pytest.approx(0.9)
    assert tso.adjust_dt(1.0, 1e-2) == pytest.approx(0.28460498941515416, rel=1e-12)
    assert tso.adjust_dt(1.0, 1e-16) == pytest.approx(0.9)
    assert tso.adjust_dt(100.0, 1e-9) == pytest.approx(40.0)
    assert tso.adjust_dt(1e-6, 1.0) == pytest.approx(1e-4)
    assert tso.adjust_dt(10.0, 1e-9, safety_factor=0.5, max_dt=100.0) == pytest.approx(100.0)
    assert tso.adjust_dt(1.0, float("nan")) == pytest.approx(1e-4)  # nan error -> min_dt (reference quirk)


def test_decide_does_not_forward_tol():
    class Fake:
        snapshot, taxis = [0], [0.0]

        def record_log(self, *a):
            pass

    nxt = tso.decide(Fake(), "state", 5e-3, 1.0, tol=1e-2)  # accepted with the user tol ...
    assert Fake.taxis == [0.0, 1.0] and Fake.snapshot[-1] == "state"
    assert nxt == pytest.approx(tso.adjust_dt(1.0, 5e-3))  # ... but dt is adjusted with the default 1e-3


# ---- sharding ------------------------------------------------------------------------------------------
def test_shard_bounds_cover_everything():
    for n, w in [(10, 3), (4096, 8), (7, 8), (1_000_000, 8), (5, 1)]:
        blocks = [shard_bounds(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_ensemble_validation():
    e = PDS1D_Ensemble()
    with pytest.raises(ValueError):
        e.set_mesh(np.array([1.0]))
    e.set_mesh(np.linspace(0, 9, 10))
    with pytest.raises(ValueError):
        e.set_diffusivity(np.ones(10))
    with pytest.raises(ValueError):
        e.set_zonal_diffusivity(np.ones((3, 2)), np.full(10, 2))
    with pytest.raises(ValueError):
        e.set_observations([1, 1], np.zeros((5, 2)))
    with pytest.raises(ValueError):
        e.set_source("not a list")
    e.set_zonal_diffusivity(np.ones((3, 2)), np.arange(10) // 5)
    assert e.n_models == 3


# ---- sweep driver: candidate generators (host only) ---------------------------------------------
def test_sweep_candidate_generators():
    from fiberis_b200.simulator.optimizer import sweep
    g = sweep.grid_candidates([(-1, 1, 3), (0, 2, 2)])
    assert g.shape == (6, 2)
    np.testing.assert_array_equal(g[:, 0], [-1, -1, 0, 0, 1, 1])
    np.testing.assert_array_equal(g[:, 1], [0, 2, 0, 2, 0, 2])
    r = sweep.random_candidates(100, [-1, 0, 2], [1, 0.5, 3], seed=3)
    assert r.shape == (100, 3) and np.all(r >= [-1, 0, 2]) and np.all(r <= [1, 0.5, 3])
    np.testing.assert_array_equal(r, sweep.random_candidates(100, [-1, 0, 2], [1, 0.5, 3], seed=3))
    assert not np.array_equal(r, sweep.random_candidates(100, [-1, 0, 2], [1, 0.5, 3], seed=4))


def test_fixed_time_axis_is_the_reference_loop_bit_for_bit():
    """pds.py:287,309 accumulates taxis[-1] + dt in fp64 and stops at the first value >= t_total; the chunked
    np.cumsum version must give the same list (length and every entry) for dyadic and non-dyadic steps."""
    from fiberis_b200.simulator.core.pds import fixed_time_axis, fixed_time_axis_array, fixed_time_axis_pair
    rng = np.random.default_rng(0)

    def loop(t0, dt, tt):
        t = [t0]
        while t[-1] < tt:
            t.append(t[-1] + dt)
        return t
    for _ in range(400):
        t0 = float(rng.choice([0.0, rng.uniform(-5, 5)]))
        dt = float(rng.choice([0.1, 0.5, 1.0, 1 / 3, rng.uniform(1e-3, 2.0)]))
        n = int(rng.integers(1, 3000))
        tt = float(rng.choice([t0 + n * dt, t0 + rng.uniform(0, 1) * n * dt, round(t0 + n * dt, 3)]))
        assert fixed_time_axis(t0, dt, tt) == loop(t0, dt, tt)
        assert fixed_time_axis_array(t0, dt, tt).tolist() == loop(t0, dt, tt)  # the ensemble's array form
    assert fixed_time_axis_array(0, 0.1, 1).tolist() == loop(0, 0.1, 1) and fixed_time_axis_array(0.0, 1.0, 5000.0).shape == (5001,)
    for args in [(0, 1, 10), (0, 1, 1000), (0.0, 0.1, 100.0), (np.int64(0), 1.0, 500.0)]:  # the single-model path's pair
        lst, arr = fixed_time_axis_pair(*args)
        assert lst == loop(*args) and [type(v) for v in lst[:2]] == [type(v) for v in fixed_time_axis(*args)[:2]]
        assert arr is None or (arr.tolist() == [float(v) for v in lst] and arr.dtype == np.array(lst).dtype)
    assert fixed_time_axis_pair(0, 1, 10)[1] is None and fixed_time_axis_pair(0, 1, 1000)[1] is not None
    assert len(fixed_time_axis(0, 0.1, 1)) == 12 and fixed_time_axis(3.0, 1.0, 2.0) == [3.0]
    assert fixed_time_axis(0.0, 0.1, 100.0) == loop(0.0, 0.1, 100.0)


def test_pinned_pool_drops_unused_buffers_by_identity():
    """More than 7 pooled buffers: the oldest unused ones are dropped — by identity (the entries hold tensors, and
    list.remove would compare them element by element and raise on different sizes)."""
    import torch
    from fiberis_b200 import _engine as E
    pool = E._PinnedPool()
    real_empty = torch.empty
    try:  # no GPU here: page-locked allocation is replaced by a plain one
        torch.empty = lambda *a, pin_memory=False, **k: real_empty(*a, **k)
        sizes = [1000 * (k + 1) ** 2 for k in range(12)]
        for n in sizes:
            e = pool.take(n)
            assert e[0].numel() >= n
            pool.release(e)  # (what the death of the array that aliased the buffer amounts to)
        assert len(pool.entries) <= 8
        # a reservation holds until the caller installs its weak reference or releases: the same buffer is never handed
        # out twice, also not to a second thread that asks between take() and the caller's bookkeeping
        a, b = pool.take(5000), pool.take(5000)
        assert a is not b and a[0].data_ptr() != b[0].data_ptr()
        pool.release(a)
        assert pool.take(5000) is a
        import threading
        got, lock = [], threading.Lock()

        def grab():
            for _ in range(200):
                e = pool.take(4096)
                with lock:
                    got.append(id(e))
        ts = [threading.Thread(target=grab) for _ in range(4)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert len(set(got)) == len(got) == 800
    finally:
        torch.empty = real_empty
