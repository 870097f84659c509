"""Host-side logic of the sharded ensemble reduction on CPU: world_size 2 and 3 over gloo. The runs -> columns
exchange and the gather on rank 0 are exercised with uneven run counts; the per-column selection itself is the CUDA
kernel in the product (tests/test_gpu_ensemble.py) and is replaced here by a numpy checker."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from memilio_b200 import ensemble, study


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _numpy_select(p):
    def sel(segs, counts):
        data = np.concatenate([s[:, :c].numpy() for s, c in zip(segs, counts)], axis=1)
        srt = np.sort(data, axis=1)
        total = data.shape[1]
        return np.stack([srt[:, int(total * q)] for q in p]), data.mean(axis=1)
    return sel


def _worker(rank, world, port, num_runs, ncol, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)
        full = rng.normal(size=(ncol, num_runs))  # every rank can rebuild the global matrix
        run0, cnt = study.distribute_runs(num_runs, world, rank)
        stride = cnt + 3  # padded rows like the device record
        local = np.full((ncol, stride), np.nan)
        local[:, :cnt] = full[:, run0:run0 + cnt]
        p = ensemble.DEFAULT_PERCENTILES
        pct, mean = ensemble.reduce_distributed(torch.from_numpy(local), cnt, rank, world, p, True,
                                                select=_numpy_select(p))
        if rank == 0:
            srt = np.sort(full, axis=1)
            for i, q in enumerate(p):
                np.testing.assert_array_equal(pct[i], srt[:, int(num_runs * q)])
            np.testing.assert_allclose(mean, full.mean(axis=1), rtol=1e-12, atol=1e-14)
            np.save(out_path, pct)
        else:
            assert pct is None and mean is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_runs,ncol", [(2, 11, 23), (3, 10, 7), (2, 1, 5)])
def test_sharded_reduction_plumbing(tmp_path, world, num_runs, ncol):
    out = str(tmp_path / "pct.npy")
    mp.spawn(_worker, args=(world, _free_port(), num_runs, ncol, out), nprocs=world, join=True)
    assert os.path.exists(out)


def test_column_partition_covers_all_columns():
    for ncol, world in ((410060, 8), (5, 8), (24, 2)):
        parts = ensemble.column_partition(ncol, world)
        assert parts[0][0] == 0 and sum(c for _, c in parts) == ncol
        for (a, n), (b, _) in zip(parts[:-1], parts[1:]):
            assert a + n == b
