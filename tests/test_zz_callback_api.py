"""The reference's OWN signatures -- mpsort_mpi_impl / mpsort_mpi_newarray_impl /
radix_sort with a host radix() callback, rsize and arg (reference mpsort.h:1-4,25-47)
-- on top of the descriptor path.

CPU part (no GPU): the host pieces of that path (descriptor for an rsize-byte radix,
pack {radix | record}, unpack) composed with the ORACLE's descriptor sort must give
exactly what the unmodified reference's radix_sort gives with the same callback
(oracle/_ref/libradixsort-ref.so) -- i.e. the augmented-record construction preserves
the reference's ordering for every rsize class of radixsort.c:146-195.
GPU part: the entry points themselves, against the same references.

Collected last on purpose (file name): these are the newest entry points."""
import ctypes
import os

import numpy as np
import pytest

import mpsort_oracle as O

import mpsort
from mpsort import _capi as C

lib = C.lib
REF_SO = os.path.join(O.REF_DIR, "libradixsort-ref.so")


def ref_radix_sort(rec, cb, rsize):
    """the unmodified reference radix_sort (radixsort.c:35-44) with a Python callback"""
    dll = ctypes.CDLL(REF_SO, mode=getattr(os, "RTLD_DEEPBIND", 0) | os.RTLD_NOW)
    f = dll.radix_sort
    f.restype = None
    f.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, C.RadixFunc, ctypes.c_size_t, ctypes.c_void_p]
    out = np.ascontiguousarray(rec).copy()
    f(out.ctypes.data, len(out), out.shape[1], cb, rsize, None)
    return out


def two_field_key(rsize, elsize, rng_fields):
    """test-issue7.c:12-17 style: the radix is assembled from two separate fields of the
    record (low part from `lo`, high part from `hi`), truncated / zero-extended to rsize"""
    (lo_off, lo_len), (hi_off, hi_len) = rng_fields

    def radix(ptr, out, arg):
        src = (ctypes.c_ubyte * elsize).from_address(ptr)
        dst = (ctypes.c_ubyte * rsize).from_address(out)
        key = bytes(src[lo_off:lo_off + lo_len]) + bytes(src[hi_off:hi_off + hi_len])
        key = (key + bytes(rsize))[:rsize]
        for i in range(rsize):
            dst[i] = key[i]
    return C.RadixFunc(radix)


def python_keys(rec, rsize, rng_fields):
    (lo_off, lo_len), (hi_off, hi_len) = rng_fields
    k = np.concatenate([rec[:, lo_off:lo_off + lo_len], rec[:, hi_off:hi_off + hi_len],
                        np.zeros((len(rec), rsize), np.uint8)], axis=1)[:, :rsize]
    return np.ascontiguousarray(k)


def stable_order(keys):
    """argsort of little-endian rsize-byte integers, stable (radixsort.c:65-98)"""
    cols = [keys[:, i] for i in range(keys.shape[1])]        # lexsort: last key is primary
    return np.lexsort(cols)


# (rsize, elsize, ((lo_off, lo_len), (hi_off, hi_len))): every comparator class of
# radixsort.c:146-195 -- 2, 4, 8 native; 16 and 24 u64 words; 3, 5, 12 byte-wise
CASES = [
    (8, 8, ((0, 8), (0, 0))),          # bench-mpi.c: the record is the key
    (4, 4, ((0, 4), (0, 0))),          # main-mpi.c: raw int
    (12, 24, ((0, 4), (16, 8))),       # test-issue7.c: 4 bytes @0 below 8 bytes @16
    (16, 40, ((8, 8), (24, 8))),
    (24, 40, ((0, 8), (16, 16))),
    (2, 6, ((3, 2), (0, 0))),
    (3, 7, ((1, 1), (4, 2))),
    (5, 16, ((11, 5), (0, 0))),
    (8, 16, ((8, 4), (0, 4))),
]


def make_records(n, elsize, seed, distinct=None):
    rng = np.random.default_rng(seed)
    rec = rng.integers(0, 256, size=(n, elsize), dtype=np.uint8)
    if distinct:                       # heavy duplicates: few distinct values per byte
        rec = (rec % distinct).astype(np.uint8)
        rec[:, -1] = np.arange(n, dtype=np.uint64) % 251      # payload byte that tells equal keys apart
    return rec


@pytest.mark.parametrize("rsize,elsize,fields", CASES)
def test_callback_pack_desc_unpack_reproduce_the_reference_order(rsize, elsize, fields):
    cb = two_field_key(rsize, elsize, fields)
    d = C.RadixDesc()
    rpad = ctypes.c_size_t(0)
    assert lib.mpsort_callback_desc(rsize, ctypes.byref(d), ctypes.byref(rpad)) == 0
    assert rpad.value % 8 == 0 and rpad.value >= rsize and d.width * d.nwords == rsize and d.offset == 0
    for n, distinct in ((0, None), (1, None), (257, None), (300, 3)):
        rec = make_records(n, elsize, 77 + n + rsize, distinct)
        aug = np.full((n, rpad.value + elsize), 0xAB, np.uint8)
        lib.mpsort_callback_pack(rec.ctypes.data, n, elsize, cb, rsize, None, aug.ctypes.data)
        assert np.array_equal(aug[:, :rsize], python_keys(rec, rsize, fields))
        assert not aug[:, rsize:rpad.value].any()
        assert np.array_equal(aug[:, rpad.value:], rec)
        sorted_aug = O.c_radix_sort(aug, O.Desc(d.offset, d.width, d.nwords, d.is_signed, 0)) if n else aug
        got = np.zeros_like(rec)
        lib.mpsort_callback_unpack(sorted_aug.ctypes.data, n, elsize, rsize, got.ctypes.data)
        exp = rec[stable_order(python_keys(rec, rsize, fields))] if n else rec
        assert np.array_equal(got, exp)
        if O.have_ref() and os.path.exists(REF_SO):
            assert np.array_equal(ref_radix_sort(rec, cb, rsize), exp), "the reference itself disagrees"


def test_callback_desc_rejects_bad_rsize():
    d = C.RadixDesc()
    rpad = ctypes.c_size_t(0)
    assert lib.mpsort_callback_desc(0, ctypes.byref(d), ctypes.byref(rpad)) != 0
    assert lib.mpsort_callback_desc(129, ctypes.byref(d), ctypes.byref(rpad)) != 0
    assert lib.mpsort_callback_desc(128, ctypes.byref(d), ctypes.byref(rpad)) == 0 and d.nwords == 16 and d.width == 8


# ------------------------------------------------------------------ on the GPU
@pytest.mark.gpu
@pytest.mark.parametrize("rsize,elsize,fields", CASES)
def test_radix_sort_with_callback_matches_reference(rsize, elsize, fields):
    cb = two_field_key(rsize, elsize, fields)
    for n, distinct in ((0, None), (1, None), (1000, None), (5000, 3)):
        rec = make_records(n, elsize, 99 + n + rsize, distinct)
        got = rec.copy()
        lib.radix_sort(got.ctypes.data, n, elsize, cb, rsize, None)
        exp = rec[stable_order(python_keys(rec, rsize, fields))] if n else rec
        assert np.array_equal(got, exp), "n=%d" % n


@pytest.mark.gpu
@pytest.mark.parametrize("rsize,elsize,fields", [CASES[0], CASES[2], CASES[7]])
@pytest.mark.parametrize("inplace", [False, True])
def test_mpsort_mpi_with_callback_matches_contract(rsize, elsize, fields, inplace):
    """4 rank threads, uneven sizes, duplicates: global stable sort by the callback's radix,
    cut by the output sizes (the contract sentence of SURVEY.md 8a), through
    mpsort_mpi_impl / mpsort_mpi_newarray_impl"""
    p = 4
    sizes = [700, 0, 1301, 555]
    outsizes = sizes if inplace else [639, 639, 639, 639]
    recs = [make_records(sizes[r], elsize, 500 + r, 5) for r in range(p)]
    cbs = [two_field_key(rsize, elsize, fields) for _ in range(p)]
    ins = [r.copy() for r in recs]
    outs = ins if inplace else [np.zeros((outsizes[r], elsize), np.uint8) for r in range(p)]
    lib.mpsort_mpi_unset_options(-1)
    lib.mpsort_mpi_set_options(C.MPSORT_DISABLE_GATHER_SORT)

    def work(comm):
        r = comm.rank
        if inplace:
            lib.mpsort_mpi_impl(ins[r].ctypes.data, len(ins[r]), elsize, cbs[r], rsize, None, comm.handle, 0, b"test_cb")
        else:
            lib.mpsort_mpi_newarray_impl(ins[r].ctypes.data, len(ins[r]), outs[r].ctypes.data, len(outs[r]), elsize,
                                         cbs[r], rsize, None, comm.handle, 0, b"test_cb")

    mpsort.run_local(p, work)
    lib.mpsort_mpi_unset_options(-1)
    allrec = np.concatenate(recs)
    exp = allrec[stable_order(python_keys(allrec, rsize, fields))]
    cuts = np.cumsum([0] + list(outsizes))
    for r in range(p):
        assert np.array_equal(outs[r], exp[cuts[r]:cuts[r + 1]]), "rank %d" % r
