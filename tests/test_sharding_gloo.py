"""world_size-2 gloo test of the only collective on the path: the all-gather of per-member
misfits over contiguous shards (NCCL on GPUs, gloo here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fiberis_b200.simulator.core.ensemble import gather_misfit, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_models, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_models, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float64) * 1.5 + 0.25  # stands in for this rank's misfits
    full = gather_misfit(local, n_models)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_models", [10, 11, 1])
def test_gather_misfit_two_ranks(tmp_path, n_models):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_models, str(tmp_path)), nprocs=world, join=True)
    want = np.arange(n_models) * 1.5 + 0.25
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{r}.npy"), want)
