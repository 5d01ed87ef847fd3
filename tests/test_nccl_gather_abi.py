"""fbr_nccl_allgather_f64: the path's one collective through the C ABI with a communicator the CALLER created (SURVEY 8(b):
"accept an externally created ncclComm_t").  Two processes, two GPUs, NCCL bound with ctypes (ncclGetUniqueId /
ncclCommInitRank) — no torch.distributed anywhere.  Skipped with fewer than two GPUs."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nccl_path():
    import torch  # noqa: F401  (its bundled NCCL)
    import site
    for base in site.getsitepackages():
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return hits[0]
    return "libnccl.so.2"


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _worker(rank, world, id_file, out_dir):
    import torch

    from fiberis_b200 import _cabi
    torch.cuda.set_device(rank)
    nccl = C.CDLL(_nccl_path(), mode=C.RTLD_GLOBAL)
    uid = _UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
        with open(id_file + ".tmp", "wb") as f:
            f.write(C.string_at(C.addressof(uid), 128))  # (the id contains NUL bytes: not bytes(uid.internal))
        os.rename(id_file + ".tmp", id_file)
    else:
        import time
        while not os.path.exists(id_file):
            time.sleep(0.05)
        raw = open(id_file, "rb").read()
        assert len(raw) == 128
        C.memmove(C.addressof(uid), raw, 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    n = 1000
    send = (torch.arange(n, dtype=torch.float64, device="cuda") + 10000.0 * rank).contiguous()
    recv = torch.zeros(world * n, dtype=torch.float64, device="cuda")
    lib = _cabi.load()
    _cabi.check(lib.fbr_nccl_allgather_f64(_cabi.ptr(send), _cabi.ptr(recv), n, comm, _cabi.current_stream()))
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), recv.cpu().numpy())
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)


def test_allgather_with_a_caller_made_communicator(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=world, join=True)
    want = np.concatenate([np.arange(1000.0) + 10000.0 * r for r in range(world)])
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{r}.npy"), want)
