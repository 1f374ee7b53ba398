"""ddstore_b200/comm.py -- communicator adapters.

The reference is constructed with an mpi4py communicator (`PyDDStore(MPI.Comm comm, int method=0)`,
src/pyddstore.pyx:61). MPI is not required here: the store needs exactly two collectives (a bootstrap
all-gather of ~100 bytes per rank, and a barrier for the fences), so anything that can provide those
two is a communicator:

  SelfComm()                 one rank (MPI_COMM_SELF)
  ShmComm(key, rank, size)   native POSIX-shm rendezvous for the ranks of one box (processes or threads)
  TorchDistComm(group=None)  torch.distributed process group (gloo or nccl)
  as_dds_comm(obj)           any of the above, or a duck-typed mpi4py-style comm
                             (Get_rank / Get_size / allgather / Barrier)
"""
import ctypes as C

from . import _capi


class _Comm:
    handle = None

    def Get_rank(self):
        return _capi.lib().dds_comm_rank(self.handle)

    def Get_size(self):
        return _capi.lib().dds_comm_size(self.handle)

    def Barrier(self):
        _capi.raise_for(_capi.lib().dds_comm_barrier(self.handle))

    def allgather_bytes(self, payload: bytes):
        n = len(payload)
        size = self.Get_size()
        send = C.create_string_buffer(payload, n)
        recv = C.create_string_buffer(n * size)
        _capi.raise_for(_capi.lib().dds_comm_allgather(self.handle, send, recv, n))
        return [recv.raw[i * n:(i + 1) * n] for i in range(size)]

    def close(self):
        if self.handle:
            _capi.lib().dds_comm_free(self.handle)
            self.handle = None

    def Split(self, color, key):
        """MPI_Comm_split: the ranks that pass the same `color` form a new communicator, ordered by `key` (ties by
        old rank). COLLECTIVE over this communicator. The reference builds its replica groups this way
        (examples/vae/distdataset.py:28: comm.Split(rank // ddstore_width, rank))."""
        me = (int(color), int(key), self.Get_rank())
        parts = self.allgather_bytes(b"".join(int(x).to_bytes(8, "little", signed=True) for x in me))
        rows = [tuple(int.from_bytes(p[i * 8:(i + 1) * 8], "little", signed=True) for i in range(3)) for p in parts]
        members = sorted((k, r) for (c, k, r) in rows if c == int(color))
        return self._sub([r for _, r in members], int(color))

    def _sub(self, old_ranks, color):
        raise NotImplementedError(f"{type(self).__name__} cannot be split")


class SelfComm(_Comm):
    def __init__(self):
        self.handle = _capi.lib().dds_comm_self()

    def _sub(self, old_ranks, color):
        return SelfComm()


class ShmComm(_Comm):
    def __init__(self, key, rank, size):
        self.key = str(key)
        self.handle = _capi.lib().dds_comm_shm(self.key.encode(), int(rank), int(size))
        if not self.handle:
            raise RuntimeError(_capi.last_error())

    def _sub(self, old_ranks, color):
        # a fresh shm rendezvous per colour, named after the parent's key
        return ShmComm(f"{self.key}.c{color}", old_ranks.index(self.Get_rank()), len(old_ranks))


class CallbackComm(_Comm):
    """rank/size + two Python callables: allgather(bytes) -> list[bytes], barrier()."""

    def __init__(self, rank, size, allgather, barrier):
        self._allgather, self._barrier = allgather, barrier

        def _ag(ctx, send, recv, n):
            try:
                parts = self._allgather(C.string_at(send, n))
                C.memmove(recv, b"".join(parts), n * size)
                return 0
            except Exception:  # noqa: BLE001 -- must not unwind through C
                import traceback
                traceback.print_exc()
                return 1

        def _bar(ctx):
            try:
                self._barrier()
                return 0
            except Exception:  # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        self._cb = (_capi.ALLGATHER_FN(_ag), _capi.BARRIER_FN(_bar))  # keep alive
        self.handle = _capi.lib().dds_comm_callbacks(int(rank), int(size), self._cb[0], self._cb[1], None)
        if not self.handle:
            raise RuntimeError(_capi.last_error())


class TorchDistComm(CallbackComm):
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._group = group
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        on_cuda = dist.get_backend(group) == "nccl"

        def allgather(payload):
            dev = torch.device("cuda", torch.cuda.current_device()) if on_cuda else torch.device("cpu")
            t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
            outs = [torch.empty_like(t) for _ in range(size)]
            dist.all_gather(outs, t, group=group)
            return [bytes(o.cpu().numpy().tobytes()) for o in outs]

        def barrier():
            dist.barrier(group=group)

        super().__init__(rank, size, allgather, barrier)

    def Split(self, color, key):
        import torch.distributed as dist
        me = (int(color), int(key), self.Get_rank())
        parts = self.allgather_bytes(b"".join(int(x).to_bytes(8, "little", signed=True) for x in me))
        rows = [tuple(int.from_bytes(p[i * 8:(i + 1) * 8], "little", signed=True) for i in range(3)) for p in parts]
        mine = None
        # new_group is collective over the WHOLE default group: every rank creates every colour's group, in one order
        for c in sorted({c for (c, _, _) in rows}):
            local = [r for _, r in sorted((k, r) for (cc, k, r) in rows if cc == c)]
            ranks = local if self._group is None else [dist.get_global_rank(self._group, r) for r in local]
            g = dist.new_group(ranks=ranks)
            if c == int(color):
                mine = g
        return TorchDistComm(mine)


def as_dds_comm(obj):
    """Coerce what a caller passes as `comm` into a communicator object with a `.handle`."""
    if obj is None:
        return SelfComm()
    if isinstance(obj, _Comm):
        return obj
    if all(hasattr(obj, a) for a in ("Get_rank", "Get_size", "allgather", "Barrier")):
        # mpi4py.MPI.Comm and look-alikes
        c = CallbackComm(obj.Get_rank(), obj.Get_size(), lambda b: list(obj.allgather(b)), obj.Barrier)
        if hasattr(obj, "Split"):
            c.Split = lambda color, key: as_dds_comm(obj.Split(color, key))
        return c
    raise TypeError(f"cannot use {type(obj).__name__} as a DDStore communicator")
