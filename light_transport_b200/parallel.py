"""Multi-GPU plumbing: photons (sample indices) shard across ranks, films add up.

SURVEY.md §8(e): rank ``r`` of ``R`` renders samples ``[r*spp/R, (r+1)*spp/R)`` of EVERY pixel.
Philox counters are the global path id ``(i*W+j)*spp + s`` so the ranks' streams are disjoint by
construction and the union over ranks is bit-identical to what one rank would have drawn.  The
only exchange of a pass is ONE sum-reduction of the ``H*W*3`` float64 film.  On GPUs that collective
lives behind the C-ABI (``lt_film_reduce``: ``ncclReduce`` over NVLink/NVSwitch on the library's own
communicator, fused with the finalize on the root, NCCL errors mapped to ``LT_E_NCCL``);
``torch.distributed`` is the plumbing that carries the 128-byte NCCL id from rank 0 to the others.
CPU tensors (the gloo tests) go through ``torch.distributed`` itself.  The reference has no
counterpart — it is single-process (``numba.prange`` threads only, R/src/path_tracing_old.py:142-149).
"""
from __future__ import annotations

import ctypes as C


def sample_range(rank: int, world_size: int, spp: int) -> tuple[int, int]:
    """Contiguous, gap-free, balanced split of ``range(spp)``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return rank * spp // world_size, (rank + 1) * spp // world_size


class FilmComm:
    """The library's NCCL communicator of one ``torch.distributed`` group on one device (``lt_comm``)."""

    def __init__(self, device: int, group=None):
        import torch
        import torch.distributed as dist
        from . import host
        L = host.lib()
        for name in ("lt_comm_unique_id", "lt_comm_init", "lt_film_reduce", "lt_nccl_version"):
            getattr(L, name).restype = C.c_int
        L.lt_comm_destroy.restype = None
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self._ctx = host._get_ctx(device)
        ident = (C.c_char * 128)()
        if self.rank == 0:
            host._check(L.lt_comm_unique_id(ident))
        # the id travels through torch.distributed (a device tensor for NCCL groups, a host tensor for gloo)
        backend = dist.get_backend(group)
        dev = torch.device(f"cuda:{device}") if backend == "nccl" else torch.device("cpu")
        t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).to(dev)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        self._comm = C.c_void_p()
        host._check(L.lt_comm_init(self._ctx, C.c_int(self.size), C.c_int(self.rank), C.c_char_p(raw), C.byref(self._comm)))

    def reduce(self, film, root: int | None = 0, spp_total: int = 0):
        """``film`` (CUDA float64 tensor) <- sum over ranks on ``root`` (None: every rank); ``spp_total > 0`` also finalizes
        there, in place: ``clip(film / spp_total, 0, 1)``.  Runs on the context's current stream, synchronous on return."""
        from . import host
        host._check(host.lib().lt_film_reduce(self._ctx, self._comm, C.cast(C.c_void_p(film.data_ptr()), host.PD), C.c_int64(film.numel()),
                                              C.c_int(-1 if root is None else root), C.c_int64(int(spp_total))))
        return film

    def close(self):
        from . import host
        if self._comm and host._ctx_cache.get(self.device) is self._ctx:
            host.lib().lt_comm_destroy(self._comm)
        self._comm = C.c_void_p()


_comms: dict = {}


def film_comm(device: int, group=None) -> FilmComm:
    """One communicator per (device, group), created on first use (collective: every rank of the group must call)."""
    key = (device, id(group) if group is not None else None)
    c = _comms.get(key)
    if c is None:
        c = _comms[key] = FilmComm(device, group)
    return c


def close_comms():
    for c in _comms.values():
        c.close()
    _comms.clear()


def reduce_film(film, group=None, root: int | None = None, *, device: int | None = None, spp_total: int = 0):
    """Sum ``film`` (a torch tensor) over the ranks, in place.  ``root=None`` → every rank gets the sum; else only ``root``.
    CUDA tensors go through the library's own NCCL communicator (``lt_film_reduce``; with ``spp_total > 0`` the finalize is
    fused on the receiving rank); CPU tensors (gloo) through ``torch.distributed``."""
    import torch.distributed as dist
    if film.is_cuda:
        dev = film.device.index if device is None else device
        return film_comm(dev, group).reduce(film, root, spp_total)
    if root is None:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(film, dst=dist.get_global_rank(group, root) if group is not None else root, op=dist.ReduceOp.SUM, group=group)
    if spp_total > 0 and (root is None or dist.get_rank(group) == root):
        film.div_(spp_total).clamp_(0, 1)
    return film
