"""mpsort (B200): drop-in for the `mpsort` Python package of MP-sort v0.1.19, backed
by libmpsort-b200.so (hand-written sm_100a CUDA + NCCL). Same names and argument
meaning as the reference's mpsort/__init__.py; `comm` is an `mpsort.Comm`.

Importing this package loads the CUDA library and the compiled binding; both are
required (there is no CPU fallback).
"""
import numpy

from .version import __version__  # noqa: F401
from . import _capi  # noqa: F401  (raises ImportError if libmpsort-b200.so is missing)
from .comm import Comm, run_local, world  # noqa: F401
from .binding import sort as _sort, radix_desc  # noqa: F401


class DeviceArray(object):
    """A 1-d struct array in GPU memory exposing `__cuda_array_interface__` -- the
    smallest possible provider, so that device-resident sorting needs neither CuPy
    nor PyTorch. `dtype` is the record layout (any numpy dtype)."""

    def __init__(self, n, dtype, device=0):
        self.dtype = numpy.dtype(dtype)
        self.n = int(n)
        self.device = device
        self.nbytes = self.n * self.dtype.itemsize
        self.ptr = _capi.lib.mpsort_util_dev_malloc(device, self.nbytes)

    def __len__(self):
        return self.n

    @property
    def __cuda_array_interface__(self):
        d = {"shape": (self.n,), "typestr": self.dtype.str if self.dtype.fields is None else "|V%d" % self.dtype.itemsize,
             "data": (self.ptr, False), "version": 3, "strides": None}
        if self.dtype.fields is not None:
            d["descr"] = self.dtype.descr
        return d

    @classmethod
    def from_host(cls, array, device=0):
        array = numpy.ascontiguousarray(array)
        self = cls(len(array), array.dtype, device)
        _capi.lib.mpsort_util_memcpy(device, self.ptr, array.ctypes.data, self.nbytes)
        return self

    def to_host(self):
        out = numpy.empty(self.n, dtype=self.dtype)
        _capi.lib.mpsort_util_memcpy(self.device, out.ctypes.data, self.ptr, self.nbytes)
        return out

    def free(self):
        if self.ptr:
            _capi.lib.mpsort_util_dev_free(self.device, self.ptr)
            self.ptr = None


def sort(source, orderby=None, out=None, comm=None, tuning=[]):
    """
        Sort source array with orderby as the key. Store result to out.
        (reference: mpsort/__init__.py:23-85)

        Parameters
        ----------
        source : array, 1d, distributed (numpy array, flatiter, or device array)

        orderby : array, 1d, distributed or string.
            Only integer types are supported.
            must be on the same partition as that of source.
            If orderby is string, it refers to the field in source.

        out : array, 1d distributed
            the total length must be the same as source.
            the itemsize must be the same as source
            if None, the sort is in-place.

        tuning: list of strings
            'DISABLE_SPARSE_ALLTOALLV', 'REQUIRE_SPARSE_ALLTOALLV',
            'DISABLE_GATHER_SORT', 'REQUIRE_GATHER_SORT'
            ('ENABLE_SPARSE_ALLTOALLV' of the reference's docstring is accepted: it is
            the default policy.)

        Returns
        -------
            out
    """
    key = orderby
    if isinstance(key, (str, bytes)):
        return _sort(source, key, out, comm=comm, tuning=tuning)

    if hasattr(source, "__cuda_array_interface__") and not isinstance(source, numpy.ndarray):
        if key is not None:
            raise ValueError("device arrays are sorted by a field name (or by themselves with orderby=None)")
        return _sort(source, None, out, comm=comm, tuning=tuning)

    # pack (data, key) into one struct array so that the C sort sees a single record
    if key is None:
        D, I = "DD"
        packed = numpy.empty(len(source), dtype=[("D", guess_dtype(source))])
        packed["D"][...] = source
    else:
        D, I = "DI"
        packed = numpy.empty(len(source), dtype=[("D", guess_dtype(source)), ("I", guess_dtype(key))])
        packed["D"][...] = source
        packed["I"][...] = key

    if out is None:
        out = source
        _sort(packed, orderby=I, comm=comm, tuning=tuning)
        out[...] = packed[D][...]
    else:
        packed_out = numpy.empty(len(out), dtype=packed.dtype)
        _sort(packed, orderby=I, out=packed_out, comm=comm, tuning=tuning)
        out[...] = packed_out[D][...]
    return out


def globalrange(array, comm):
    """start and end of the local chunk in the global array (__init__.py:87-94)"""
    sizes = comm.allgather(len(array))
    start = sum(sizes[:comm.rank])
    return (start, start + sizes[comm.rank])


def globalindices(array, comm):
    """indices of the local chunk in the global array (__init__.py:96-109)"""
    start, end = globalrange(array, comm)
    globalsize = comm.bcast(end, root=comm.size - 1)
    dtype = "i8" if globalsize > 1024 * 1024 * 1024 else "i4"
    return numpy.arange(start, end, dtype=dtype)


def guess_dtype(array):
    if isinstance(array, numpy.flatiter):
        return array.base.dtype, ()
    return array.dtype, array.shape[1:]


def permute(source, argindex, comm, out=None):
    """source[argindex], distributed like argindex (__init__.py:116-149): two sorts."""
    source_size = comm.allreduce(len(source))
    argindex_size = comm.allreduce(len(argindex))
    if source_size != argindex_size:
        raise ValueError("Global size of source and argindex is different")
    if out is None:
        out = numpy.empty(len(argindex), guess_dtype(source))
    originind = globalindices(argindex, comm)
    originind2 = numpy.empty(len(source), dtype=originind.dtype)
    sort(originind, orderby=argindex, out=originind2, comm=comm)
    sort(source, orderby=originind2, out=out, comm=comm)
    return out


def histogram(array, bins, comm, right=False):
    """global histogram of a distributed array over collective bin edges
    (__init__.py:151-172)"""
    if len(array) == 0:
        originrank = []
    else:
        originrank = numpy.digitize(array, bins, right)
    recv = numpy.bincount(originrank, minlength=len(bins) + 1)
    return comm.allreduce(recv)


def take(source, argindex, comm, out=None):
    """source[argindex] distributed like argindex; argindex need not be a permutation
    (__init__.py:174-204): three sorts."""
    start, end = globalrange(source, comm)
    bins = comm.allgather(end)
    h = histogram(argindex, bins, comm)
    nactive = h[comm.rank]
    if out is None:
        out = numpy.empty(len(argindex), guess_dtype(source))
    originind = globalindices(argindex, comm)
    myargindex = numpy.empty(nactive, dtype=guess_dtype(argindex))
    myoriginind = numpy.empty(nactive, dtype=originind.dtype)
    sort(originind, orderby=argindex, out=myoriginind, comm=comm)
    sort(argindex, orderby=argindex, out=myargindex, comm=comm)
    myresult = source[myargindex - start]
    sort(myresult, orderby=myoriginind, out=out, comm=comm)
    return out
