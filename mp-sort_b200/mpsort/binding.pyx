# cython: language_level=3, embedsignature=True
"""Cython binding of libmpsort-b200.so: same surface as the reference's
mpsort/binding.pyx (`sort(data, orderby, out, comm, tuning)`, binding.pyx:123-209).

What differs from the reference, and why:
  * the radix() callback + RadixData (binding.pyx:44-121) become a
    `struct mpsort_radix_desc` {offset, width, nwords, is_signed}: the key is read on
    the GPU, a host function pointer cannot run there;
  * `comm` is an `mpsort.Comm` (NCCL / in-process group) instead of an mpi4py
    communicator: no MPI exists on the target;
  * `data` / `out` may also be device arrays: any object with
    `__cuda_array_interface__` is sorted in place on the GPU without a host copy;
  * the GIL is released around the C call so that rank-threads of an in-process
    group can run their collective concurrently.
There is no CPU fallback: the C library aborts without a CUDA device.
"""
from libc.stdint cimport uint32_t, int32_t, uintptr_t

import numpy

cdef extern from "mpsort.h":
    int MPSORT_DISABLE_SPARSE_ALLTOALLV
    int MPSORT_DISABLE_GATHER_SORT
    int MPSORT_REQUIRE_GATHER_SORT
    int MPSORT_REQUIRE_SPARSE_ALLTOALLV

    struct mpsort_radix_desc:
        size_t offset
        uint32_t width
        uint32_t nwords
        int32_t is_signed
        int32_t reserved

    struct mpsort_comm:
        pass
    ctypedef mpsort_comm * mpsort_comm_t

    void mpsort_mpi_set_options(int options)
    void mpsort_mpi_unset_options(int options)
    void mpsort_mpi_newarray_desc_impl(void * base, size_t nmemb,
            void * out, size_t outnmemb, size_t elsize,
            const mpsort_radix_desc * desc, mpsort_comm_t comm,
            const int line, const char * file) nogil


def _describe(obj, name):
    """(pointer, nmemb, dtype, keepalive) of a numpy array or a device array."""
    if isinstance(obj, numpy.ndarray):
        if not obj.flags['C_CONTIGUOUS']:
            raise ValueError("%s must be C_CONTIGUOUS" % name)
        return obj.ctypes.data, len(obj), obj.dtype, obj
    cai = getattr(obj, '__cuda_array_interface__', None)
    if cai is None:
        raise TypeError("%s must be a numpy.ndarray or expose __cuda_array_interface__" % name)
    if cai.get('strides') is not None:
        raise ValueError("%s must be C_CONTIGUOUS" % name)
    descr = cai.get('descr')
    if descr is not None and not (len(descr) == 1 and descr[0][0] == ''):
        dtype = numpy.dtype([tuple(d) for d in descr])
    else:
        dtype = numpy.dtype(cai['typestr'])
    shape = cai['shape']
    if len(shape) != 1:
        raise ValueError("%s must be one dimensional" % name)
    ptr = cai['data'][0]
    return (ptr if ptr is not None else 0), int(shape[0]), dtype, obj


def radix_desc(dtype, radixkey):
    """(offset, width, nwords, is_signed) of data[radixkey]: the rules of
    radix_data_init (binding.pyx:50-79)."""
    dtype = numpy.dtype(dtype)
    if radixkey is not None:
        if dtype.fields is None or radixkey not in dtype.fields:
            raise ValueError("no field of name %s" % (radixkey,))
        radixdtype, offset = dtype.fields[radixkey][:2]
    else:
        radixdtype, offset = dtype, 0
    if len(radixdtype.shape) == 0:
        nmemb = 1
    elif len(radixdtype.shape) == 1:
        nmemb = radixdtype.shape[0]
    else:
        raise ValueError("data[%s] is not 1d nor 2d" % (radixkey,))
    base = radixdtype.base
    if base == numpy.dtype('u8'):
        width, signed = 8, 0
    elif base == numpy.dtype('i8'):
        width, signed = 8, 1
    elif base == numpy.dtype('u4'):
        width, signed = 4, 0
    elif base == numpy.dtype('i4'):
        width, signed = 4, 1
    else:
        raise TypeError("data[%s] is not u8 or i8" % (radixkey,))
    return offset, width, nmemb, signed


def sort(data, orderby=None, out=None, comm=None, tuning=[]):
    """
        Parallel sort of distributed data set `data' over communicator `comm',
        ordered by key given in 'orderby'.

        Parameters
        ----------
        data : numpy.ndarray or device array (__cuda_array_interface__)
            the input data; must be C contiguous.

        orderby : string or None
            data[orderby] must be of integer types (u8, i8, u4, i4).
            data[orderby] can be 2d, in which case the latter elements in a row has
            more significance. if orderby is None, use data itself.

        out : numpy.ndarray, device array or None
            the output array; if None, inplace

        comm : mpsort.Comm or None
            the communicator, None for the world communicator of the launcher

        tuning: list of strings
            'DISABLE_SPARSE_ALLTOALLV'
            'DISABLE_GATHER_SORT'
            'REQUIRE_GATHER_SORT'
            'REQUIRE_SPARSE_ALLTOALLV'
            ('ENABLE_SPARSE_ALLTOALLV', which the reference documents but never
             implemented, is accepted and means the default policy)
    """
    cdef mpsort_radix_desc desc
    cdef uintptr_t inptr, outptr, commptr
    cdef size_t nin, nout, elsize
    cdef mpsort_comm_t ccomm

    from .comm import Comm, world

    if isinstance(data, numpy.ndarray):
        # assert you can access the orderby columns (binding.pyx:157)
        key = data[orderby] if orderby is not None else data

    p_in, n_in, dtype, keep_in = _describe(data, "data")

    if out is None:
        out = data
    p_out, n_out, odtype, keep_out = _describe(out, "out")

    if comm is None:
        comm = world()
    elif not isinstance(comm, Comm):
        raise ValueError("only mpsort.Comm objects are supported")

    Ntot = comm.allreduce(n_in)
    Ntotout = comm.allreduce(n_out)

    if Ntot != Ntotout:
        raise ValueError("total size of array changed %d != %d" % (Ntot, Ntotout))

    if dtype.itemsize != odtype.itemsize:
        raise ValueError("item size mismatch")

    offset, width, nwords, signed = radix_desc(dtype, orderby)
    desc.offset = offset
    desc.width = width
    desc.nwords = nwords
    desc.is_signed = signed
    desc.reserved = 0

    # process-global, like the reference (binding.pyx:193-204); rank-threads of one
    # process all write the same bits
    # only the four tuning bits are reset per call: bits outside the reference's set (the
    # MPSORT_VERIFY_CHECKSUM extension, also set from the environment, which is parsed once)
    # stay as the caller or the environment left them
    mpsort_mpi_unset_options(MPSORT_DISABLE_SPARSE_ALLTOALLV | MPSORT_DISABLE_GATHER_SORT
                             | MPSORT_REQUIRE_GATHER_SORT | MPSORT_REQUIRE_SPARSE_ALLTOALLV)
    if 'DISABLE_SPARSE_ALLTOALLV' in tuning:
        mpsort_mpi_set_options(MPSORT_DISABLE_SPARSE_ALLTOALLV)
    if 'DISABLE_GATHER_SORT' in tuning:
        mpsort_mpi_set_options(MPSORT_DISABLE_GATHER_SORT)
    if 'REQUIRE_GATHER_SORT' in tuning:
        mpsort_mpi_set_options(MPSORT_REQUIRE_GATHER_SORT)
    if 'REQUIRE_SPARSE_ALLTOALLV' in tuning:
        mpsort_mpi_set_options(MPSORT_REQUIRE_SPARSE_ALLTOALLV)
    # every rank-thread must have set its bits before any of them reads them
    comm.barrier()

    inptr = p_in
    outptr = p_out
    nin = n_in
    nout = n_out
    elsize = dtype.itemsize
    commptr = comm.handle.value
    ccomm = <mpsort_comm_t> commptr
    with nogil:
        mpsort_mpi_newarray_desc_impl(<void *> inptr, nin, <void *> outptr, nout, elsize,
                                      &desc, ccomm, 0, "mpsort/binding.pyx")
    return out
