"""Ensemble results on the device: record of the interpolated daily results of all runs and their reduction
(`sum_nodes`, `ensemble_mean`, `ensemble_percentile`, cpp/memilio/data/analyze_result.h:267-343), single GPU and
sharded over ranks (runs -> columns exchange over torch.distributed, NCCL on GPUs)."""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import cdefs, flat
from .study import distribute_runs

ENS_FULL, ENS_AGE_SUM = 0, 1
DEFAULT_PERCENTILES = (0.05, 0.25, 0.5, 0.75, 0.95)  # save_results, cpp/memilio/io/result_io.h:150-190


def _lib():
    from . import capi
    L = capi.lib()
    if not getattr(L, "_ens_bound", False):
        L.mio_cuda_ensemble_create.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int64, C.POINTER(C.c_void_p)]
        L.mio_cuda_ensemble_destroy.argtypes = [C.c_void_p]
        L.mio_cuda_ensemble_destroy.restype = None
        L.mio_cuda_attach_ensemble.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.mio_cuda_ensemble_info.argtypes = [C.c_void_p, cdefs.c_int64_p, cdefs.c_int64_p, cdefs.c_int64_p,
                                             C.POINTER(C.c_void_p)]
        L.mio_cuda_ensemble_set_window.argtypes = [C.c_void_p, C.c_int32]
        L.mio_cuda_comm_get_unique_id.argtypes = [C.c_char_p]
        L.mio_cuda_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
        L.mio_cuda_comm_destroy.argtypes = [C.c_void_p]
        L.mio_cuda_comm_destroy.restype = None
        L.mio_cuda_ensemble_reduce_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                                       cdefs.c_double_p, cdefs.c_double_p, cdefs.c_double_p]
        L.mio_cuda_ensemble_reduce.argtypes = [C.c_void_p, C.c_int32, cdefs.c_double_p, cdefs.c_double_p,
                                               cdefs.c_double_p]
        L.mio_cuda_select_percentiles.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_void_p), cdefs.c_int64_p,
                                                  cdefs.c_int64_p, C.c_int64, C.c_int32, cdefs.c_double_p,
                                                  cdefs.c_double_p, cdefs.c_double_p]
        L._ens_bound = True
    return L


def _check(rc):
    from . import capi
    capi._check(rc)


class EnsembleRecord:
    """Device matrix [column][run]; columns = per-node (node, day, element) then node-summed (day, element)."""

    def __init__(self, num_groups: int, num_nodes: int, num_days: int, capacity_runs: int, mode: int = ENS_AGE_SUM,
                 device: int = 0, model: int = flat.MODEL_OSECIR):
        self.G, self.N, self.D, self.mode, self.device = num_groups, num_nodes, num_days, mode, device
        self.C = flat.num_compartments(model)
        self.n = self.G * self.C
        self.El = self.C if mode == ENS_AGE_SUM else self.n
        self._h = C.c_void_p()
        _check(_lib().mio_cuda_ensemble_create(device, model, num_groups, num_nodes, num_days, mode, capacity_runs,
                                               C.byref(self._h)))

    def attach(self, sim, run_offset: int):
        _check(_lib().mio_cuda_attach_ensemble(sim._h, self._h, run_offset))

    def set_window(self, first_day: int):
        """Record the days [first_day, first_day + num_days) since t0 into the slots [0, num_days)."""
        _check(_lib().mio_cuda_ensemble_set_window(self._h, int(first_day)))

    def detach(self, sim):
        _check(_lib().mio_cuda_attach_ensemble(sim._h, None, 0))

    def info(self):
        a = np.zeros(3, dtype=np.int64)
        p = C.c_void_p()
        _check(_lib().mio_cuda_ensemble_info(self._h, cdefs.lptr(a[0:1]), cdefs.lptr(a[1:2]), cdefs.lptr(a[2:3]),
                                             C.byref(p)))
        return dict(num_columns=int(a[0]), stride=int(a[1]), num_runs=int(a[2]), device_ptr=int(p.value or 0))

    def split(self, cols: np.ndarray):
        """[..., num_columns] -> (per node [..., N, D, El], node sum [..., D, n])"""
        k = self.N * self.D * self.El
        lead = cols.shape[:-1]
        return cols[..., :k].reshape(lead + (self.N, self.D, self.El)), cols[..., k:].reshape(lead + (self.D, self.n))

    def reduce(self, p: Sequence[float] = DEFAULT_PERCENTILES, mean: bool = True):
        """(percentiles[num_p][num_columns], mean[num_columns] or None) of the recorded runs (one GPU)."""
        ncol = self.info()["num_columns"]
        pa = np.ascontiguousarray(p, dtype=np.float64)
        out = np.zeros((len(pa), ncol))
        mn = np.zeros(ncol) if mean else None
        _check(_lib().mio_cuda_ensemble_reduce(self._h, len(pa), cdefs.dptr(pa), cdefs.dptr(out), cdefs.dptr(mn)))
        return out, mn

    def as_torch(self):
        """The record as a torch CUDA tensor [num_columns][stride] (no copy)."""
        import torch
        i = self.info()

        class _Arr:
            __cuda_array_interface__ = {"shape": (i["num_columns"], i["stride"]), "typestr": "<f8",
                                        "data": (i["device_ptr"], False), "version": 2}
        return torch.as_tensor(_Arr(), device=f"cuda:{self.device}")

    def close(self):
        if self._h:
            _lib().mio_cuda_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- sharded reduction in the C library (NCCL) -------------------------------------------------------------------
COMM_ID_BYTES = 128
_COMMS = {}


class Comm:
    """mio_comm: the library's own NCCL communicator (one rank per GPU). The 128-byte unique id of rank 0 reaches the
    other ranks through torch.distributed here (the application's job: MPI_Bcast in the reference's setting)."""

    def __init__(self, device: int, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist
        buf = C.create_string_buffer(COMM_ID_BYTES)
        if rank == 0:
            _check(_lib().mio_cuda_comm_get_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.to(f"cuda:{device}")
        dist.broadcast(t, src=0, group=group)
        uid = bytes(t.cpu().tolist())
        self._h = C.c_void_p()
        _check(_lib().mio_cuda_comm_create(device, rank, world, uid, C.byref(self._h)))
        self.device, self.rank, self.world = device, rank, world

    def close(self):
        if self._h:
            _lib().mio_cuda_comm_destroy(self._h)
            self._h = C.c_void_p()


def get_comm(device: int, rank: int, world: int, group=None) -> Comm:
    key = (device, rank, world)
    if key not in _COMMS:
        _COMMS[key] = Comm(device, rank, world, group)
    return _COMMS[key]


def reduce_sharded(rec: Optional["EnsembleRecord"], count_local: int, num_columns: int, comm: Comm,
                   p: Sequence[float] = DEFAULT_PERCENTILES, mean: bool = True):
    """mio_cuda_ensemble_reduce_sharded: exchange (NCCL), exact selection of this rank's column block over the runs of
    all ranks, gather on rank 0. Returns (percentiles, mean) on rank 0, (None, None) elsewhere. Collective."""
    pa = np.ascontiguousarray(p, dtype=np.float64)
    root = comm.rank == 0
    out = np.zeros((len(pa), num_columns)) if root else None
    mn = np.zeros(num_columns) if (root and mean) else None
    _check(_lib().mio_cuda_ensemble_reduce_sharded(rec._h if rec is not None else None, comm._h, count_local, num_columns,
                                                   len(pa), cdefs.dptr(pa), cdefs.dptr(out), cdefs.dptr(mn)))
    return (out, mn) if root else (None, None)


# ---- sharded reduction -------------------------------------------------------------------------------------------
def column_partition(num_columns: int, world: int) -> List[Tuple[int, int]]:
    """(first column, count) owned by each rank after the exchange (same rule as distribute_runs)."""
    return [distribute_runs(num_columns, world, r) for r in range(world)]


def exchange_columns(local, count_local: int, rank: int, world: int, group=None):
    """Re-shard [num_columns][>=count_local] (runs of this rank) into the column block of this rank with the runs of
    every rank: returns a list of `world` tensors [my_columns][count_s]. Point-to-point exchange (the all-to-all of
    SURVEY 8e) over torch.distributed; NCCL on GPUs, gloo in the CPU tests."""
    import torch
    import torch.distributed as dist
    ncol = local.shape[0]
    parts = column_partition(ncol, world)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([count_local], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    c0, cn = parts[rank]
    recv = [torch.empty((cn, counts[s]), dtype=local.dtype, device=local.device) for s in range(world)]
    send = [local[parts[d][0]:parts[d][0] + parts[d][1], :count_local].contiguous() for d in range(world)]
    recv[rank].copy_(send[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if send[peer].numel() > 0:
            ops.append(dist.P2POp(dist.isend, send[peer], peer, group=group))
        if recv[peer].numel() > 0:
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv, counts


def select_cuda(segments, counts, p: Sequence[float], mean: bool, device: int):
    """ensemble_percentile / ensemble_mean over run segments held as torch CUDA tensors (the product path)."""
    S = len(segments)
    ncol = int(segments[0].shape[0])
    ptrs = (C.c_void_p * S)(*[C.c_void_p(int(t.data_ptr()) if t.numel() else 0) for t in segments])
    strides = np.array([max(int(t.shape[1]), 0) for t in segments], dtype=np.int64)
    cnts = np.array(counts, dtype=np.int64)
    pa = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros((len(pa), ncol))
    mn = np.zeros(ncol) if mean else None
    _check(_lib().mio_cuda_select_percentiles(device, S, ptrs, cdefs.lptr(strides), cdefs.lptr(cnts), ncol, len(pa),
                                              cdefs.dptr(pa), cdefs.dptr(out), cdefs.dptr(mn)))
    return out, mn


def gather_columns(local_cols, rank: int, world: int, num_columns: int, device, group=None):
    """Gather the per-rank column blocks [k][my_columns] on rank 0 -> [k][num_columns] (like the reference gathers the
    results on the root, parameter_studies.h:190-212). Returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    parts = column_partition(num_columns, world)
    k = local_cols.shape[0]
    t = torch.as_tensor(np.ascontiguousarray(local_cols)).to(device)
    if rank == 0:
        out = np.zeros((k, num_columns))
        out[:, parts[0][0]:parts[0][0] + parts[0][1]] = local_cols
        for s in range(1, world):
            if parts[s][1] == 0:
                continue
            buf = torch.empty((k, parts[s][1]), dtype=torch.float64, device=device)
            dist.recv(buf, src=s, group=group)
            out[:, parts[s][0]:parts[s][0] + parts[s][1]] = buf.cpu().numpy()
        return out
    if parts[rank][1] > 0:
        dist.send(t, dst=0, group=group)
    return None


def reduce_distributed(local, count_local: int, rank: int, world: int, p: Sequence[float] = DEFAULT_PERCENTILES,
                       mean: bool = True, select: Optional[Callable] = None, group=None, rec=None):
    """Sharded ensemble reduction: exchange, local exact selection, gather on rank 0.
    `local` is this rank's record [num_columns][stride] (torch tensor). Returns (percentiles, mean) on rank 0.
    On GPUs (NCCL process group) the whole reduction runs in the C library (mio_cuda_ensemble_reduce_sharded, `rec` =
    the record, or None on a rank without runs); the torch.distributed path below serves the CPU tests (gloo)."""
    ncol = int(local.shape[0])
    if world > 1 and select is None and local.is_cuda:
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            dev = local.device.index if local.device.index is not None else 0
            return reduce_sharded(rec if count_local > 0 else None, count_local, ncol, get_comm(dev, rank, world, group),
                                  p, mean)
    if world == 1:
        segs, counts = [local], [count_local]  # row length = stride, the first count_local entries are valid
    else:
        segs, counts = exchange_columns(local, count_local, rank, world, group)
    if select is None:
        dev = local.device.index if local.device.index is not None else 0
        select = lambda s, c: select_cuda(s, c, p, mean, dev)  # noqa: E731
    pct, mn = select(segs, counts)
    rows = pct if mn is None else np.concatenate([pct, mn[None, :]], axis=0)
    if world == 1:
        full = rows
    else:
        full = gather_columns(rows, rank, world, ncol, local.device, group)
    if full is None:
        return None, None
    return (full[:len(p)], full[len(p)] if mean else None)


def windowed_advance_reduce(sim, rec, local, count: int, t0: float, tmax: float, num_days: int, day_block: int,
                            rank: int, world: int, p: Sequence[float] = DEFAULT_PERCENTILES, mean: bool = True,
                            timer: Optional[Callable] = None):
    """Advance the resident runs of `sim` (None on a rank without runs) to tmax one window of `day_block` days at a
    time and reduce every window (sharded when world > 1); `rec` is sized for one window. Returns the rows
    [len(p) (+1 for the mean)][N*num_days*El + num_days*n] of the whole period on rank 0 (None elsewhere): the same
    columns as a record of all days would give. `timer(kind)` is called before ("advance"/"reduce") phases."""
    N, El, n = rec.N, rec.El, rec.n
    blk = max(1, min(int(day_block), num_days))
    k = len(p) + (1 if mean else 0)
    full = np.zeros((k, N * num_days * El + num_days * n)) if rank == 0 else None
    day0 = int(np.ceil(t0 - 1e-14))
    for d0 in range(0, num_days, blk):
        d1 = min(num_days, d0 + blk)  # window = days [d0, d1) since t0
        if timer:
            timer("advance")
        if sim is not None:
            rec.set_window(d0)
            sim.advance(min(tmax, float(day0 + d1 - 1)))
        if timer:
            timer("reduce")
        pct, mn = reduce_distributed(local, count, rank, world, p, mean, rec=rec if sim is not None else None)
        if rank == 0:
            rows = pct if mn is None else np.concatenate([pct, mn[None, :]], axis=0)
            node_part = rows[:, :N * blk * El].reshape(k, N, blk, El)
            sum_part = rows[:, N * blk * El:].reshape(k, blk, n)
            full[:, :N * num_days * El].reshape(k, N, num_days, El)[:, :, d0:d1] = node_part[:, :, :d1 - d0]
            full[:, N * num_days * El:].reshape(k, num_days, n)[:, d0:d1] = sum_part[:, :d1 - d0]
    if sim is not None and tmax > sim.t:
        sim.advance(tmax)
    if timer:
        timer("end")
    return full
