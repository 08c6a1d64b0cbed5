"""The reference's own pytest file (mpsort/tests/test_mpsort.py), re-hosted on
`mpsort.Comm`: same test names, same split/heal helpers, same assertions (numpy is
the oracle there too). `mpirun -n 4 pytest --with-mpi` becomes 4 rank threads of an
in-process group sharing cuda:0 (`mpsort.run_local`); test_issue7 uses 12."""
import base64  # noqa: F401
from itertools import product

import numpy
import pytest
from numpy.testing import assert_array_equal

import mpsort
from conftest import load_golden

pytestmark = pytest.mark.gpu


def split(array, comm, localsize=None):
    array = comm.bcast(array)
    if localsize is None:
        sp = numpy.array_split(array, comm.size)
        return comm.scatter(sp)
    else:
        g = comm.allgather(localsize)
        return comm.scatter(numpy.array_split(array, numpy.cumsum(g)[:-1]))


def heal(array, comm):
    a = comm.allgather(array)
    a = numpy.concatenate(a, axis=0)
    return a


def adjustsize(size, comm):
    ressize = size + 1 - 2 * ((comm.rank) % 2)
    if comm.size % 2 == 1:
        if comm.rank == 0:
            ressize = size
    return ressize


def mpi(nranks=4):
    """decorator: run the test body on `nranks` rank threads"""
    def deco(fn):
        def wrapper(*args, **kwargs):
            mpsort.run_local(nranks, lambda comm: fn(comm, *args, **kwargs))
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def _scalar_sort(comm, make):
    rng = numpy.random.RandomState(comm.size)
    s = make(rng)
    local = split(s, comm)
    s = heal(local, comm)
    mpsort.sort(local, orderby=None, out=None, comm=comm)
    r = heal(local, comm)
    s.sort()
    assert_array_equal(s, r)


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_sort_i4(nranks):
    mpsort.run_local(nranks, lambda comm: _scalar_sort(comm, lambda rng: numpy.int32(rng.random_sample(1000) * 1000 - 400)))


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_sort_i8(nranks):
    mpsort.run_local(nranks, lambda comm: _scalar_sort(comm, lambda rng: numpy.int64(rng.random_sample(1000) * 1000 - 400)))


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_sort_u8(nranks):
    mpsort.run_local(nranks, lambda comm: _scalar_sort(
        comm, lambda rng: numpy.int64(rng.uniform(size=1000, low=-1000000, high=1000000) * 1000 - 400).astype(numpy.uint64)))


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_sort_u4(nranks):
    mpsort.run_local(nranks, lambda comm: _scalar_sort(
        comm, lambda rng: numpy.int64(rng.random_sample(1000) * 1000 - 400).astype(numpy.uint32)))


TUNINGS = [
    [],
    ['DISABLE_SPARSE_ALLTOALLV'],
    ['REQUIRE_SPARSE_ALLTOALLV'],
    ['REQUIRE_GATHER_SORT'],
    ['DISABLE_GATHER_SORT'],
]


@pytest.mark.parametrize("tuning", TUNINGS)
def test_sort_tunings(tuning):
    @mpi(4)
    def body(comm):
        s = numpy.int32(numpy.random.RandomState(1).random_sample(size=1000) * 1000)
        local = split(s, comm)
        s = heal(local, comm)
        mpsort.sort(local, orderby=None, out=None, comm=comm, tuning=tuning)
        r = heal(local, comm)
        s.sort()
        assert_array_equal(s, r)
    body()


@mpi(4)
def test_sort_inplace(comm):
    s = numpy.int32(numpy.random.RandomState(2).random_sample(size=1000) * 1000)
    local = split(s, comm)
    s = heal(local, comm)
    mpsort.sort(local, local, out=None, comm=comm)
    r = heal(local, comm)
    s.sort()
    assert_array_equal(s, r)


@mpi(4)
def test_sort_mismatched_zeros(comm):
    s = numpy.int32(numpy.random.RandomState(3).random_sample(size=1000) * 1000)
    local = split(s, comm, [0, 400, 0, 600][comm.rank])
    s = heal(local, comm)
    res = split(s, comm, [200, 200, 0, 600][comm.rank])
    res[:] = numpy.int32(numpy.random.random(size=res.size) * 1000)
    mpsort.sort(local, local, out=res, comm=comm, tuning=['REQUIRE_GATHER_SORT'])
    s.sort()
    r = heal(res, comm)
    assert_array_equal(s, r)


@pytest.mark.parametrize("nranks", [3, 4])
def test_sort_outplace(nranks):
    @mpi(nranks)
    def body(comm):
        s = numpy.int32(numpy.random.RandomState(4).random_sample(size=1000) * 1000)
        local = split(s, comm)
        s = heal(local, comm)
        res = numpy.zeros(adjustsize(local.size, comm), dtype=local.dtype)
        mpsort.sort(local, local, out=res, comm=comm)
        s.sort()
        r = heal(res, comm)
        assert_array_equal(s, r)
    body()


@mpi(4)
def test_sort_flatiter(comm):
    s = numpy.int32(numpy.random.RandomState(5).random_sample(size=1000) * 1000)
    local = split(s, comm)
    s = heal(local, comm)
    res = numpy.zeros(adjustsize(local.size, comm), dtype=local.dtype)
    mpsort.sort(local.flat, local.flat, out=res.flat, comm=comm)
    s.sort()
    r = heal(res, comm)
    assert_array_equal(s, r)


@mpi(4)
def test_sort_struct(comm):
    s = numpy.empty(10, dtype=[('value', 'i8'), ('key', 'i8')])
    rng = numpy.random.RandomState(1234)
    s['value'] = numpy.int32(rng.random_sample(size=10) * 1000 - 400)
    s['key'] = s['value']
    backup = s.copy()
    local = split(s, comm)
    s = heal(local, comm)
    res = numpy.zeros_like(local)
    mpsort.sort(local, 'key', out=res, comm=comm)
    r = heal(res, comm)
    backup.sort(order='key')
    assert_array_equal(backup['value'], r['value'])


@mpi(4)
def test_sort_struct_vector(comm):
    s = numpy.empty(10, dtype=[('value', 'i8'), ('key', 'i8'), ('vkey', ('i8', 2))])
    s['value'] = numpy.int32(numpy.random.RandomState(6).random_sample(size=len(s)) * 1000)
    s['key'][:][...] = s['value']
    s['vkey'][:, 0][...] = s['value']
    s['vkey'][:, 1][...] = s['value']
    local = split(s, comm)
    s = heal(local, comm)
    res = numpy.zeros_like(local)
    mpsort.sort(local, 'vkey', out=res, comm=comm)
    r = heal(res, comm)
    s.sort(order='key')
    assert_array_equal(s['value'], r['value'])


@mpi(4)
def test_sort_vector(comm):
    s = numpy.empty(10, dtype=[('value', 'i8')])
    s['value'] = numpy.int32(numpy.random.RandomState(7).random_sample(size=len(s)) * 1000)
    local = split(s, comm)
    s = heal(local, comm)
    k = numpy.empty(len(local), ('i8', 2))
    k[:, 0][...] = local['value']
    k[:, 1][...] = local['value']
    res = numpy.zeros_like(local)
    mpsort.sort(local, k, out=res, comm=comm)
    s.sort(order='value')
    r = heal(res, comm)
    assert_array_equal(s['value'], r['value'])


@mpi(4)
def test_permute(comm):
    s = numpy.arange(10)
    local = split(s, comm)
    i = numpy.arange(9, -1, -1)
    ind = split(i, comm, adjustsize(local.size, comm))
    res = mpsort.permute(local, ind, comm)
    r = heal(res, comm)
    s = s[i]
    assert res.size == ind.size
    assert_array_equal(r, s)


@mpi(4)
def test_permute_out(comm):
    s = numpy.arange(10)
    local = split(s, comm)
    i = numpy.arange(9, -1, -1)
    ind = split(i, comm, adjustsize(local.size, comm))
    res = numpy.empty(ind.size, local.dtype)
    mpsort.permute(local, ind, comm, out=res)
    r = heal(res, comm)
    s = s[i]
    assert_array_equal(r, s)


@mpi(4)
def test_take(comm):
    s = numpy.arange(10)
    local = split(s, comm)
    i = numpy.arange(9, -1, -1)
    ind = split(i, comm, adjustsize(local.size, comm))
    res = mpsort.take(local, ind, comm)
    r = heal(res, comm)
    s = s[i]
    assert res.size == ind.size
    assert_array_equal(r, s)


@mpi(4)
def test_take_out(comm):
    s = numpy.arange(10)
    local = split(s, comm)
    i = numpy.arange(9, -1, -1)
    ind = split(i, comm, adjustsize(local.size, comm))
    res = numpy.empty(ind.size, local.dtype)
    mpsort.take(local, ind, comm, out=res)
    r = heal(res, comm)
    s = s[i]
    assert_array_equal(r, s)


def test_version():
    assert hasattr(mpsort, "__version__")


@mpi(4)
def test_histogram_empty(comm):
    mpsort.histogram([], [1], comm)
    # no error shall be raised


@pytest.mark.parametrize("tuning", TUNINGS)
def test_empty_sort(tuning):
    @mpi(4)
    def body(comm):
        s = numpy.empty(0, dtype=[('vkey', ('u8', 3)), ('vector', ('u4', 3))])
        mpsort.sort(s, 'vkey', out=s, comm=comm, tuning=tuning)
    body()


@pytest.mark.parametrize("tuning", TUNINGS)
def test_few_items(tuning):
    """all 3^4 size combinations of {0,1,2} per rank (405 cases in the reference)"""
    @mpi(4)
    def body(comm):
        for sizes in product(*([[0, 1, 2]] * 4)):
            A = [range(sizes[i]) for i in range(len(sizes))]
            s = numpy.empty(len(A[comm.rank]), dtype=[('vkey', ('u8', 3)), ('vector', ('u4', 3))])
            s['vkey'] = numpy.array(A[comm.rank], dtype='u8')[:, None]
            s['vector'] = 1
            S = numpy.concatenate(comm.allgather(s))
            S.sort()
            r = numpy.empty(len(A[comm.rank]), dtype=s.dtype)
            mpsort.sort(s, 'vkey', out=r, comm=comm, tuning=tuning)
            R = numpy.concatenate(comm.allgather(r))
            assert_array_equal(R['vkey'], S['vkey'])
    body()


@pytest.mark.parametrize("tuning", TUNINGS)
def test_issue7(tuning):
    """12 ranks, 40-byte records, 16-byte radix; whole-record equality with lexsort.
    The fixture is the reference's Issue7B64 (tests/golden/issue7.npz holds the same
    records, generated by tests/golden/make_golden.py)."""
    g = load_golden("issue7")
    dt = numpy.dtype([('radix', ('u8', 2)), ('ext', ('u8', 3))])

    @mpi(12)
    def body(comm):
        s = g["recs"][comm.rank].copy().view(dt).reshape(-1)
        S = numpy.concatenate(comm.allgather(s))
        ind = numpy.lexsort(S['radix'].T)
        S = S[ind]
        r = numpy.empty(len(s), dtype=s.dtype)
        mpsort.sort(s, orderby='radix', out=r, comm=comm, tuning=tuning)
        R = numpy.concatenate(comm.allgather(r))
        assert_array_equal(R.flatten(), S.flatten())
    body()


# ---- beyond the reference's file: argument errors and device arrays ----------
@mpi(2)
def test_argument_errors(comm):
    a = numpy.zeros(10, dtype=[('k', 'u8'), ('f', 'f8')])
    with pytest.raises(TypeError):
        mpsort.sort(a, 'f', comm=comm)
    with pytest.raises(ValueError):
        mpsort.sort(a[::2], 'k', comm=comm)
    with pytest.raises(ValueError):
        mpsort.sort(a, 'k', out=numpy.zeros(9 + comm.rank * 3, dtype=a.dtype), comm=comm)
    with pytest.raises(ValueError):
        mpsort.sort(a, 'k', out=numpy.zeros(10, dtype='u8'), comm=comm)


@mpi(4)
def test_device_arrays(comm):
    dt = numpy.dtype([('key', 'i8'), ('tag', 'u8')])
    rng = numpy.random.RandomState(comm.rank)
    a = numpy.zeros(5000 + comm.rank, dtype=dt)
    a['key'] = rng.randint(-50, 50, size=len(a))
    a['tag'] = (comm.rank << 40) + numpy.arange(len(a))
    S = numpy.concatenate(comm.allgather(a))
    S = S[numpy.argsort(S['key'], kind='stable')]
    dev = mpsort.DeviceArray.from_host(a, comm.device)
    mpsort.sort(dev, 'key', comm=comm)
    R = numpy.concatenate(comm.allgather(dev.to_host()))
    dev.free()
    assert_array_equal(R, S)


def test_verify_checksum_survives_the_python_call():
    """MPSORT_VERIFY_CHECKSUM (mpsort.h; the reference checksums every call, mpsort-mpi.c:193,324) is not
    one of the four tuning bits the binding resets on every call: set once, it stays in force through
    mpsort.sort -- two checksum launches per sort, the bit still set afterwards."""
    from mpsort import _capi as C
    comm = mpsort.Comm.self(0)
    s = numpy.int64(numpy.random.default_rng(1).integers(-99999, 99999, size=20000))
    C.lib.mpsort_mpi_set_options(C.MPSORT_VERIFY_CHECKSUM)
    try:
        C.lib.mpsort_util_kernel_timing(comm.handle, 1)
        r = s.copy()
        mpsort.sort(r, orderby=None, comm=comm, tuning=['DISABLE_GATHER_SORT'])
        mpsort.sort(r, orderby=None, comm=comm)
        kt = C.kernel_times(comm.handle)
        assert C.lib.mpsort_mpi_has_options(C.MPSORT_VERIFY_CHECKSUM)
        assert kt["checksum"][1] == 4
        assert not C.lib.mpsort_mpi_has_options(C.MPSORT_DISABLE_GATHER_SORT)     # the tuning bits are per call
        assert_array_equal(r, numpy.sort(s))
    finally:
        C.lib.mpsort_util_kernel_timing(comm.handle, 0)
        C.lib.mpsort_mpi_unset_options(C.MPSORT_VERIFY_CHECKSUM)
        comm.destroy()
