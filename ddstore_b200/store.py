"""ddstore_b200/store.py -- `PyDDStore`: the reference's Python surface over the C-ABI.

Mirrors /root/reference/src/pyddstore.pyx:58-131 method for method (same names, argument order and
meaning, same exception types/texts): add / get / epoch_begin / epoch_end / free / init / update, plus
`get_batch`, the batched form of get() that is this package's hot path. Arrays may be NumPy arrays
(host) or CUDA tensors / __cuda_array_interface__ objects (device; fetched bytes then never leave HBM).
"""
import ctypes as C

import numpy as np

from . import _capi
from .comm import as_dds_comm

# dtypes the reference's if-chains accept (src/pyddstore.pyx:69-80, 88-99, 118-129)
_NP_OK = {np.dtype(np.int32), np.dtype(np.int64), np.dtype(np.uint8), np.dtype(np.float32),
          np.dtype(np.float64), np.dtype(np.bool_)}
_TORCH_OK = {"torch.int32", "torch.int64", "torch.uint8", "torch.float32", "torch.float64", "torch.bool"}


class _Buf:
    """pointer / shape / itemsize / residency of an ndarray, a CUDA tensor or a CAI object"""

    def __init__(self, arr, writable=False):  # `writable` documents intent at the call sites; nothing is copied
        self.keep = arr
        if isinstance(arr, np.ndarray):
            assert arr.flags.c_contiguous  # src/pyddstore.pyx:66,85,116
            if arr.dtype not in _NP_OK:
                raise NotImplementedError
            self.ptr, self.on_device = arr.ctypes.data, 0
            self.shape, self.itemsize, self.size = arr.shape, arr.dtype.itemsize, arr.size
        elif hasattr(arr, "data_ptr") and hasattr(arr, "element_size"):  # torch.Tensor
            assert arr.is_contiguous()
            if str(arr.dtype) not in _TORCH_OK:
                raise NotImplementedError
            self.ptr, self.on_device = arr.data_ptr(), 1 if arr.is_cuda else 0
            self.shape, self.itemsize, self.size = tuple(arr.shape), arr.element_size(), arr.numel()
        elif hasattr(arr, "__cuda_array_interface__"):
            cai = arr.__cuda_array_interface__
            if cai.get("strides") is not None:
                raise ValueError("device array must be C-contiguous")
            dt = np.dtype(cai["typestr"])
            if dt not in _NP_OK:
                raise NotImplementedError
            self.ptr, self.on_device = cai["data"][0], 1
            self.shape, self.itemsize = tuple(cai["shape"]), dt.itemsize
            self.size = int(np.prod(self.shape, dtype=np.int64))
        else:
            raise TypeError(f"unsupported array type {type(arr).__name__}")
        self.nbytes = self.size * self.itemsize


class _DevMem:
    """a raw device range as a __cuda_array_interface__ object (uint8)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _i64(x):
    """host int64 contiguous ndarray view of an index list/array"""
    return np.ascontiguousarray(x, dtype=np.int64)


class PyDDStore:
    def __init__(self, comm=None, method=0, device=None):
        # src/pyddstore.pyx:61-63. `comm`: see ddstore_b200.comm.as_dds_comm; `device`: CUDA ordinal
        # for this rank's shards (default: the current device).
        self._L = _capi.lib()
        self._comm = as_dds_comm(comm)
        self._h = self._L.dds_create(self._comm.handle, -1 if device is None else int(device), int(method))
        if not self._h:
            raise RuntimeError(_capi.last_error())
        self.rank, self.size = self._L.dds_rank(self._h), self._L.dds_size(self._h)
        self._itemsize = {}  # per-variable itemsize, cached for the hot path
        self._cname = {}     # name -> bytes, cached for the per-sample get() loop
        self._rowbytes = {}
        self.last_bad_index = -1

    # ---------------------------------------------------------------- reference surface
    def add(self, name, arr):
        # src/pyddstore.pyx:65-82
        b = _Buf(arr)
        nrows = b.shape[0]
        disp = b.size // b.shape[0] if b.shape[0] else int(np.prod(b.shape[1:], dtype=np.int64))
        _capi.raise_for(self._L.dds_add(self._h, name.encode(), b.ptr, nrows, disp, b.itemsize, b.on_device))

    def get(self, name, arr, start=0):
        # src/pyddstore.pyx:84-101: count = arr.shape[0]; fills arr in place
        cn = self._cname.get(name)
        if cn is None:
            cn = self._cname[name] = name.encode()
        if type(arr) is np.ndarray:  # the legacy per-sample loop: keep the Python side of the call short
            assert arr.flags.c_contiguous
            if arr.dtype not in _NP_OK:
                raise NotImplementedError
            rc = self._L.dds_get(self._h, cn, int(start), arr.shape[0], arr.dtype.itemsize, arr.ctypes.data, 0)
        else:
            b = _Buf(arr, writable=True)
            rc = self._L.dds_get(self._h, cn, int(start), b.shape[0], b.itemsize, b.ptr, b.on_device)
        if rc:
            _capi.raise_for(rc)

    def epoch_begin(self):
        _capi.raise_for(self._L.dds_epoch_begin(self._h))  # src/pyddstore.pyx:103-104

    def epoch_end(self):
        _capi.raise_for(self._L.dds_epoch_end(self._h))  # src/pyddstore.pyx:106-107

    def free(self):
        if self._h:
            self._itemsize.clear()
            _capi.raise_for(self._L.dds_free(self._h))  # src/pyddstore.pyx:109-110

    def init(self, name, nrows, disp, itemsize=1):
        _capi.raise_for(self._L.dds_init(self._h, name.encode(), int(nrows), int(disp), int(itemsize)))  # :112-113

    def update(self, name, arr, offset, stream=None, wait=True):
        # src/pyddstore.pyx:115-131. wait=False: enqueue the copy on `stream` and return (streaming ingest; the
        # source must stay valid -- pinned -- until the stream reaches it).
        b = _Buf(arr)
        if wait:
            rc = self._L.dds_update(self._h, name.encode(), b.ptr, b.shape[0], int(offset), b.itemsize, b.on_device)
        else:
            rc = self._L.dds_update_async(self._h, name.encode(), b.ptr, b.shape[0], int(offset), b.itemsize,
                                          b.on_device, self._stream_arg(stream))
        _capi.raise_for(rc)

    def ingest(self, name, arr, offset):
        """update() for a chunk of pageable host rows, pipelined inside the library (parallel staging copy + async H2D).
        Returns when `arr` has been consumed; call ingest_wait() (or an epoch fence) before reading the rows back."""
        b = _Buf(arr)
        if b.on_device:
            raise ValueError("ingest takes host arrays (use update for device arrays)")
        _capi.raise_for(self._L.dds_ingest(self._h, name.encode(), b.ptr, b.shape[0], int(offset), b.itemsize))

    def ingest_wait(self):
        _capi.raise_for(self._L.dds_ingest_wait(self._h))

    # ---------------------------------------------------------------- the batched hot path
    def get_batch(self, name, starts, counts=None, out=None, count=None, offsets=None, stream=None, wait=True,
                  overlap=False):
        """Fetch len(starts) requests in ONE kernel launch, packed back to back in request order.

        starts/counts: int64 index arrays (host ndarray/list, or CUDA int64 tensors). counts=None means
        every request fetches `count` rows (default 1): the fixed-stride fast path.
        out: destination (host ndarray or CUDA tensor) of at least the packed size; row layout is the
        caller's business exactly as with get() (src/pyddstore.pyx:84-87 never checks it either).
        offsets: optional int64 array of len(starts)+1 receiving the byte offset of every request
        (same residency as `out`). Returns the number of packed bytes.
        wait=False (device indices + device out only): enqueue on `stream` and return at once; call
        `wait()` later for the status. Several such batches may be queued on one stream.
        overlap=True (with wait=False): this batch is independent of the one queued just before it (different `out`
        and `offsets`, indices not written by it) and may overlap with its tail -- double-buffered prefetch.
        stream: the cudaStream_t handle everything of this call is enqueued on (index copy, kernels, result copy).
        stream=None means the STORE'S OWN stream, which is not ordered with anything the caller has queued elsewhere:
        device tensors passed in (indices, out, offsets) must then be complete / free to overwrite before the call --
        e.g. produced by a synchronous copy. When they were produced or are consumed on torch's current stream, pass
        stream=torch.cuda.current_stream().cuda_stream.
        Raises the reference's ValueError for the first invalid request (requests before it are delivered).
        """
        itemsize = self._itemsize.get(name)
        if itemsize is None:
            itemsize = self._itemsize[name] = self.query(name)["itemsize"]
        if out is None:
            raise ValueError("get_batch needs an `out` buffer (like get(), it never allocates)")
        ob = _Buf(out, writable=True)
        s_dev = hasattr(starts, "data_ptr") and getattr(starts, "is_cuda", False)
        if s_dev:
            nreq = starts.numel()
            sp = starts.data_ptr()
            cp = counts.data_ptr() if counts is not None else None
            keep = (starts, counts)
        else:
            sa = _i64(starts)
            nreq = sa.size
            sp = sa.ctypes.data
            ca = _i64(counts) if counts is not None else None
            cp = ca.ctypes.data if ca is not None else None
            keep = (sa, ca)
        flags = (_capi.IDX_ON_DEVICE if s_dev else 0) | (_capi.DST_ON_DEVICE if ob.on_device else 0)
        if not wait:
            flags |= _capi.NO_SYNC | (_capi.OVERLAP if overlap else 0)
        op = None
        if offsets is not None:
            fb = _Buf(offsets, writable=True)
            if fb.on_device != ob.on_device or fb.itemsize != 8 or fb.size < nreq + 1:
                raise ValueError("offsets must be int64[len(starts)+1] with the same residency as out")
            op = fb.ptr
        total, bad = C.c_int64(0), C.c_int64(-1)
        rc = self._L.dds_get_batch(self._h, name.encode(), sp, cp, 1 if count is None else int(count), nreq,
                                   itemsize, ob.ptr, ob.nbytes, op, flags,
                                   self._stream_arg(stream), C.byref(total), C.byref(bad))
        del keep
        self.last_bad_index = bad.value
        _capi.raise_for(rc)
        return total.value

    # ---------------------------------------------------------------- collective owner-push fetch
    def push_setup(self, max_requests, max_bytes):
        """COLLECTIVE: allocate and peer-map the windows of the push fetch (see dds_push_setup)."""
        _capi.raise_for(self._L.dds_push_setup(self._h, int(max_requests), int(max_bytes)))

    def get_batch_push(self, name, starts, count=1, stream=None):
        """COLLECTIVE fetch of len(starts) requests (CUDA int64 tensor) of `count` rows each, by owner-push. Returns a
        uint8 CUDA tensor VIEW of the packed rows inside this rank's window (valid until the next-but-one push step);
        the step is enqueued on `stream`, wait() reports errors."""
        import torch
        itemsize = self._itemsize.get(name)
        if itemsize is None:
            itemsize = self._itemsize[name] = self.query(name)["itemsize"]
        if not (hasattr(starts, "data_ptr") and starts.is_cuda):
            raise ValueError("get_batch_push takes a CUDA int64 tensor of start rows")
        out = C.c_void_p()
        _capi.raise_for(self._L.dds_get_batch_push(self._h, name.encode(), starts.data_ptr(), int(count), starts.numel(), itemsize,
                                                   C.byref(out), self._stream_arg(stream)))
        rb = self._rowbytes.get(name)
        if rb is None:
            rb = self._rowbytes[name] = self.query(name)["disp"] * itemsize
        nbytes = starts.numel() * int(count) * rb
        return torch.as_tensor(_DevMem(out.value or 0, nbytes), device=starts.device)

    # ---------------------------------------------------------------- per-sample index (variable-length datasets)
    def set_sample_index(self, name, row_start, row_count):
        """Register, for variable `name`, which GLOBAL rows every sample owns: sample i = rows
        [row_start[i], row_start[i] + row_count[i]). Tables: int64 host arrays or CUDA tensors; copied once."""
        dev = hasattr(row_start, "data_ptr") and getattr(row_start, "is_cuda", False)
        if dev:
            n, sp, cp, keep = row_start.numel(), row_start.data_ptr(), row_count.data_ptr(), (row_start, row_count)
        else:
            sa, ca = _i64(row_start), _i64(row_count)
            n, sp, cp, keep = sa.size, sa.ctypes.data, ca.ctypes.data, (sa, ca)
        _capi.raise_for(self._L.dds_set_sample_index(self._h, name.encode(), sp, cp, n, 1 if dev else 0))
        del keep

    def get_samples(self, name, sample_ids, out, offsets=None, stream=None, wait=True, overlap=False):
        """get_batch by SAMPLE ID: the id -> (start, count) lookup runs inside the launch, against the index
        registered with set_sample_index. Same packing / offsets / error behaviour as get_batch."""
        itemsize = self._itemsize.get(name)
        if itemsize is None:
            itemsize = self._itemsize[name] = self.query(name)["itemsize"]
        ob = _Buf(out, writable=True)
        s_dev = hasattr(sample_ids, "data_ptr") and getattr(sample_ids, "is_cuda", False)
        if s_dev:
            nreq, sp, keep = sample_ids.numel(), sample_ids.data_ptr(), sample_ids
        else:
            sa = _i64(sample_ids)
            nreq, sp, keep = sa.size, sa.ctypes.data, sa
        flags = (_capi.IDX_ON_DEVICE if s_dev else 0) | (_capi.DST_ON_DEVICE if ob.on_device else 0)
        if not wait:
            flags |= _capi.NO_SYNC | (_capi.OVERLAP if overlap else 0)
        op = None
        if offsets is not None:
            fb = _Buf(offsets, writable=True)
            if fb.on_device != ob.on_device or fb.itemsize != 8 or fb.size < nreq + 1:
                raise ValueError("offsets must be int64[len(sample_ids)+1] with the same residency as out")
            op = fb.ptr
        total, bad = C.c_int64(0), C.c_int64(-1)
        rc = self._L.dds_get_samples(self._h, name.encode(), sp, nreq, itemsize, ob.ptr, ob.nbytes, op, flags,
                                     self._stream_arg(stream), C.byref(total), C.byref(bad))
        del keep
        self.last_bad_index = bad.value
        _capi.raise_for(rc)
        return total.value

    def get_samples_multi(self, names, sample_ids, outs, offsets=None, stream=None, wait=True, overlap=False):
        """The rows of the same samples in several variables (<= 4, each with a sample index) in ONE launch:
        outs[v] (CUDA tensors) receive variable names[v]'s packed rows, offsets[v] (optional int64 CUDA tensors of
        len(ids)+1) the per-sample byte offsets. Returns the list of packed sizes (None when wait=False)."""
        nv = len(names)
        obs = [_Buf(o, writable=True) for o in outs]
        if not all(o.on_device for o in obs):
            raise ValueError("get_samples_multi delivers into device buffers")
        s_dev = hasattr(sample_ids, "data_ptr") and getattr(sample_ids, "is_cuda", False)
        if s_dev:
            nreq, sp, keep = sample_ids.numel(), sample_ids.data_ptr(), sample_ids
        else:
            sa = _i64(sample_ids)
            nreq, sp, keep = sa.size, sa.ctypes.data, sa
        flags = (_capi.IDX_ON_DEVICE if s_dev else 0) | _capi.DST_ON_DEVICE | (0 if wait else _capi.NO_SYNC)
        if overlap and not wait:
            flags |= _capi.OVERLAP
        c_names = (C.c_char_p * nv)(*[n.encode() for n in names])
        c_dsts = (C.c_void_p * nv)(*[o.ptr for o in obs])
        c_caps = (C.c_int64 * nv)(*[o.nbytes for o in obs])
        c_offs = None
        if offsets is not None:
            fbs = [_Buf(f, writable=True) for f in offsets]
            c_offs = (C.c_void_p * nv)(*[f.ptr for f in fbs])
        totals = (C.c_int64 * nv)()
        bad = C.c_int64(-1)
        rc = self._L.dds_get_samples_multi(self._h, nv, c_names, sp, nreq, c_dsts, c_caps, c_offs, flags,
                                           self._stream_arg(stream), totals, C.byref(bad))
        del keep
        self.last_bad_index = bad.value
        _capi.raise_for(rc)
        return [totals[v] for v in range(nv)] if wait else None

    @staticmethod
    def _stream_arg(stream):
        """None -> the store's own stream; a cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)
        otherwise. Handle 0 is CUDA's legacy default stream, which the C-ABI spells cudaStreamLegacy (0x1)
        because NULL there means "the store's stream"."""
        if stream is None:
            return None
        h = int(stream)
        return C.c_void_p(h if h != 0 else 1)

    def wait(self):
        """complete the batches queued with wait=False; raises like get_batch; returns packed bytes of the last"""
        total, bad = C.c_int64(0), C.c_int64(-1)
        rc = self._L.dds_batch_wait(self._h, C.byref(total), C.byref(bad))
        self.last_bad_index = bad.value
        _capi.raise_for(rc)
        return total.value

    # ---------------------------------------------------------------- extras
    def query(self, name):
        vi = _capi.VarInfo()
        _capi.raise_for(self._L.dds_query(self._h, name.encode(), C.byref(vi)))
        return {"itemsize": vi.itemsize, "disp": vi.disp, "nranks": vi.nranks, "fence_active": bool(vi.fence_active),
                "local_nrows": vi.local_nrows, "total_nrows": vi.total_nrows,
                "lenlist": [vi.lenlist[i] for i in range(vi.nranks)], "local_base": vi.local_base}

    def synth_fill(self, name, seed):
        _capi.raise_for(self._L.dds_synth_fill(self._h, name.encode(), int(seed)))

    def synth_verify(self, name, packed, starts, counts=None, count=1, offsets=None, seed=0, stream=None):
        """Check a packed batch (CUDA tensors) of variable `name` -- filled by synth_fill(name, seed) -- against the
        generator on the device. Returns (mismatching elements, rows checked, requests per owner rank)."""
        res = (C.c_uint64 * 66)()
        _capi.raise_for(self._L.dds_synth_verify(
            self._h, name.encode(), packed.data_ptr(), starts.data_ptr(), counts.data_ptr() if counts is not None else None,
            int(count), offsets.data_ptr() if offsets is not None else None, starts.numel(), int(seed),
            self._stream_arg(stream), res))
        return int(res[0]), int(res[1]), [int(res[2 + r]) for r in range(self.size)]

    def close(self):
        """non-collective teardown of this rank's handle (the collective one is free())"""
        if self._h:
            self._L.dds_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
