# distutils: language=c++
# cython: language_level=3
"""pyddstore -- Cython binding with the reference module's name and Python surface
(/root/reference/src/pyddstore.pyx:58-131: PyDDStore(comm, method=0) with add / get / epoch_begin /
epoch_end / free / init / update), bound to the C++ class in include/ddstore_b200.hpp, which is a thin
wrapper over the C-ABI (include/ddstore_b200.h) of libddstore_b200.so.

Differences from the reference binding, all at the edges:
  * `comm` is anything ddstore_b200.comm.as_dds_comm accepts (an mpi4py-style communicator, a
    torch.distributed adapter, ShmComm, None) -- MPI is not required;
  * arrays may also be CUDA tensors / __cuda_array_interface__ objects (then nothing touches the host);
  * `get_batch` fetches a whole batch in one kernel launch.
"""
from libcpp.string cimport string
from libcpp cimport bool as cbool

import numpy as np

from ddstore_b200.comm import as_dds_comm
from ddstore_b200.store import _Buf, _i64

cdef extern from *:
    """
    #include <stdexcept>
    #include <Python.h>
    static void dds_translate_exception() {
        try { throw; }
        catch (const std::invalid_argument &e) { PyErr_SetString(PyExc_ValueError, e.what()); }
        catch (const std::out_of_range &e) { PyErr_SetString(PyExc_KeyError, e.what()); }
        catch (const std::exception &e) { PyErr_SetString(PyExc_RuntimeError, e.what()); }
    }
    """
    void dds_translate_exception()

cdef extern from "ddstore_b200.h":
    ctypedef struct dds_comm_t:
        pass

cdef extern from "ddstore_b200.hpp" nogil:
    cdef cppclass DDStore:
        DDStore(int method, dds_comm_t* comm, int device) except +dds_translate_exception
        void add[T](string name, T* buffer, long nrows, int disp) except +dds_translate_exception
        void add_device[T](string name, const T* buffer, long nrows, int disp) except +dds_translate_exception
        void get[T](string name, long start, long count, T* buffer) except +dds_translate_exception
        void get_device[T](string name, long start, long count, T* buffer) except +dds_translate_exception
        long get_batch[T](string name, const long* starts, const long* counts, long fixed_count, long nreq, T* dst,
                          long cap, long* offsets, cbool on_device, void* stream) except +dds_translate_exception
        void epoch_begin() except +dds_translate_exception
        void epoch_end() except +dds_translate_exception
        void free() except +dds_translate_exception
        void init(string name, long nrows, int disp, int itemsize) except +dds_translate_exception
        void update[T](string name, T* buffer, long nrows, long offset) except +dds_translate_exception
        int rank()
        int size()


cdef class PyDDStore:
    cdef DDStore* c_ddstore
    cdef object _comm

    def __cinit__(self, comm=None, int method=0, device=None):
        self._comm = as_dds_comm(comm)
        cdef size_t h = <size_t> self._comm.handle
        cdef int dev = -1 if device is None else int(device)
        # every call below may block on the other ranks (collectives) or on the GPU: never hold the GIL across it,
        # so thread-ranks of one interpreter cannot deadlock each other
        with nogil:
            self.c_ddstore = new DDStore(method, <dds_comm_t*> h, dev)

    def __dealloc__(self):
        if self.c_ddstore != NULL:
            del self.c_ddstore
            self.c_ddstore = NULL

    @property
    def rank(self):
        return self.c_ddstore.rank()

    @property
    def size(self):
        return self.c_ddstore.size()

    def add(self, str name, arr):
        b = _Buf(arr)
        cdef long nrows = b.shape[0]
        cdef int disp = (b.size // b.shape[0]) if b.shape[0] else int(np.prod(b.shape[1:], dtype=np.int64))
        cdef size_t p = b.ptr
        cdef string nm = name.encode()
        cdef int w = b.itemsize
        if b.on_device:
            with nogil:
                if w == 1: self.c_ddstore.add_device[char](nm, <const char*> p, nrows, disp)
                elif w == 4: self.c_ddstore.add_device[int](nm, <const int*> p, nrows, disp)
                else: self.c_ddstore.add_device[long](nm, <const long*> p, nrows, disp)
        else:
            with nogil:
                if w == 1: self.c_ddstore.add[char](nm, <char*> p, nrows, disp)
                elif w == 4: self.c_ddstore.add[int](nm, <int*> p, nrows, disp)
                else: self.c_ddstore.add[long](nm, <long*> p, nrows, disp)

    def get(self, str name, arr, long start=0):
        b = _Buf(arr, writable=True)
        cdef long count = b.shape[0]
        cdef size_t p = b.ptr
        cdef string nm = name.encode()
        cdef int w = b.itemsize
        if b.on_device:
            with nogil:
                if w == 1: self.c_ddstore.get_device[char](nm, start, count, <char*> p)
                elif w == 4: self.c_ddstore.get_device[int](nm, start, count, <int*> p)
                else: self.c_ddstore.get_device[long](nm, start, count, <long*> p)
        else:
            with nogil:
                if w == 1: self.c_ddstore.get[char](nm, start, count, <char*> p)
                elif w == 4: self.c_ddstore.get[int](nm, start, count, <int*> p)
                else: self.c_ddstore.get[long](nm, start, count, <long*> p)

    def get_batch(self, str name, starts, counts=None, out=None, count=None, offsets=None, stream=None):
        """one kernel launch for len(starts) requests, packed in request order into `out`; see
        ddstore_b200.store.PyDDStore.get_batch. `out` decides the element width checked against the variable."""
        if out is None:
            raise ValueError("get_batch needs an `out` buffer")
        ob = _Buf(out, writable=True)
        s_dev = hasattr(starts, "data_ptr") and getattr(starts, "is_cuda", False)
        if bool(s_dev) != bool(ob.on_device):
            raise ValueError("the Cython get_batch wants indices and out on the same side (both host or both device)")
        cdef size_t sp, cp = 0, op = 0, dp = ob.ptr
        cdef long nreq
        if s_dev:
            nreq = starts.numel(); sp = starts.data_ptr()
            if counts is not None: cp = counts.data_ptr()
            keep = (starts, counts)
        else:
            sa = _i64(starts); nreq = sa.size; sp = sa.ctypes.data
            ca = _i64(counts) if counts is not None else None
            if ca is not None: cp = ca.ctypes.data
            keep = (sa, ca)
        if offsets is not None:
            fb = _Buf(offsets, writable=True)
            op = fb.ptr
        cdef long fixed = 1 if count is None else int(count)
        cdef long cap = ob.nbytes
        cdef size_t st = 0
        if stream is not None:
            st = int(stream) if int(stream) != 0 else 1
        cdef string nm = name.encode()
        cdef int w = ob.itemsize
        cdef cbool dev = bool(ob.on_device)
        cdef long total
        with nogil:
          if w == 1:
            total = self.c_ddstore.get_batch[char](nm, <const long*> sp, <const long*> cp, fixed, nreq, <char*> dp, cap, <long*> op, dev, <void*> st)
          elif w == 4:
            total = self.c_ddstore.get_batch[int](nm, <const long*> sp, <const long*> cp, fixed, nreq, <int*> dp, cap, <long*> op, dev, <void*> st)
          else:
            total = self.c_ddstore.get_batch[long](nm, <const long*> sp, <const long*> cp, fixed, nreq, <long*> dp, cap, <long*> op, dev, <void*> st)
        del keep
        return total

    def epoch_begin(self):
        with nogil:
            self.c_ddstore.epoch_begin()

    def epoch_end(self):
        with nogil:
            self.c_ddstore.epoch_end()

    def free(self):
        with nogil:
            self.c_ddstore.free()

    def init(self, str name, long nrows, int disp, int itemsize=1):
        cdef string nm = name.encode()
        with nogil:
            self.c_ddstore.init(nm, nrows, disp, itemsize)

    def update(self, str name, arr, long offset):
        b = _Buf(arr)
        if b.on_device:
            raise NotImplementedError("update() from a device array: use ddstore_b200.PyDDStore")
        cdef long nrows = b.shape[0]
        cdef size_t p = b.ptr
        cdef string nm = name.encode()
        cdef int w = b.itemsize
        with nogil:
            if w == 1: self.c_ddstore.update[char](nm, <char*> p, nrows, offset)
            elif w == 4: self.c_ddstore.update[int](nm, <int*> p, nrows, offset)
            else: self.c_ddstore.update[long](nm, <long*> p, nrows, offset)
