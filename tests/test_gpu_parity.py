"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI of
libmpsort-b200.so (ctypes, plain pointers), compared with the oracle on the same
inputs -- bit-exact, integer/byte work. Multi-rank cases run as an in-process group
of rank threads sharing cuda:0, the way the reference's `mpirun -n 4/12` cases are
run on one machine; NCCL process-per-GPU cases need >= 2 GPUs.

Full-size cases (BASELINE.json: 2^28 16-byte records) are checked through
size-independent properties -- sortedness, tie order by tag, an order-independent multiset
hash of whole records -- and, for configs B, 4 and 5 on one GPU, byte for byte against the unmodified
reference run on the host cores (oracle/_ref)."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import mpsort_oracle as O
from conftest import GOLDEN, GOLDEN_CASES, ROOT, load_golden

import mpsort
from mpsort import _capi as C

pytestmark = pytest.mark.gpu
lib = C.lib

TUNINGS = [0, C.MPSORT_DISABLE_SPARSE_ALLTOALLV, C.MPSORT_REQUIRE_SPARSE_ALLTOALLV,
           C.MPSORT_REQUIRE_GATHER_SORT, C.MPSORT_DISABLE_GATHER_SORT]


def cdesc(d):
    return C.RadixDesc(d.offset, d.width, d.nwords, d.is_signed, 0)


def sort_group(recs, outsizes, desc, tuning=0, inplace=False, devices=None):
    """one collective sort over len(recs) rank threads through the C ABI (host buffers)"""
    p = len(recs)
    elsize = recs[0].shape[1]
    ins = [np.ascontiguousarray(r).copy() for r in recs]
    outs = ins if inplace else [np.zeros((outsizes[r], elsize), np.uint8) for r in range(p)]
    d = cdesc(desc)
    lib.mpsort_mpi_unset_options(-1)
    if tuning:
        lib.mpsort_mpi_set_options(tuning)

    def work(comm):
        r = comm.rank
        lib.mpsort_mpi_newarray_desc_impl(ins[r].ctypes.data, len(ins[r]), outs[r].ctypes.data, len(outs[r]),
                                          elsize, ctypes.byref(d), comm.handle, 0, b"test_gpu_parity")
        return C.last_stats(comm.handle, p)

    stats = mpsort.run_local(p, work, devices)
    lib.mpsort_mpi_unset_options(-1)
    return outs, stats


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


# ------------------------------------------------------------------ local sort
LOCAL_CASES = [
    (O.Desc(0, 8, 1, 0, 0), 16), (O.Desc(8, 8, 1, 1, 0), 16), (O.Desc(0, 8, 1, 1, 0), 48),
    (O.Desc(0, 8, 1, 0, 0), 8), (O.Desc(4, 4, 1, 0, 0), 12), (O.Desc(0, 4, 1, 1, 0), 8),
    (O.Desc(0, 8, 2, 0, 0), 40), (O.Desc(8, 8, 3, 1, 0), 36), (O.Desc(0, 4, 3, 1, 0), 20),
    (O.Desc(2, 2, 1, 0, 1), 7), (O.Desc(1, 1, 5, 0, 1), 9), (O.Desc(0, 8, 1, 0, 0), 24),
    (O.Desc(0, 8, 1, 0, 0), 32), (O.Desc(16, 8, 1, 1, 0), 64), (O.Desc(3, 8, 1, 0, 0), 19),
]


@pytest.mark.parametrize("desc,elsize", LOCAL_CASES)
def test_radix_sort_desc_matches_oracle(desc, elsize):
    """radix_sort replacement vs the oracle's stable merge sort (radixsort.c:35-44)"""
    rng = np.random.default_rng(elsize * 131 + desc.offset)
    tile = 6144
    for n in [0, 1, 2, 31, 32, 33, 1000, tile - 1, tile, tile + 1, 3 * tile + 17, 200003]:
        a = rng.integers(0, 256, size=(n, elsize), dtype=np.uint8)
        got = a.copy()
        lib.radix_sort_desc(got.ctypes.data, n, elsize, ctypes.byref(cdesc(desc)), 0)
        assert np.array_equal(got, O.c_radix_sort(a, desc)), "n=%d" % n


@pytest.mark.parametrize("distinct", [1, 2, 5, 300])
def test_radix_sort_desc_stable_on_duplicates(distinct):
    rng = np.random.default_rng(distinct)
    dt = np.dtype([("key", "i8"), ("tag", "u8")])
    a = np.zeros(400000, dtype=dt)
    a["key"] = rng.integers(-distinct, distinct, size=len(a))
    a["tag"] = np.arange(len(a))
    got = a.copy()
    lib.radix_sort_desc(got.ctypes.data, len(a), 16, ctypes.byref(C.RadixDesc(0, 8, 1, 1, 0)), 0)
    assert np.array_equal(got, a[np.argsort(a["key"], kind="stable")])


def test_sorted_reverse_and_constant_inputs():
    d = C.RadixDesc(0, 8, 1, 0, 0)
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    n = 100000
    for keys in (np.arange(n), np.arange(n)[::-1], np.full(n, 7), np.arange(n) << 40, np.arange(n) % 2):
        a = np.zeros(n, dtype=dt)
        a["key"] = keys
        a["tag"] = np.arange(n)
        got = a.copy()
        lib.radix_sort_desc(got.ctypes.data, n, 16, ctypes.byref(d), 0)
        assert np.array_equal(got, a[np.argsort(a["key"], kind="stable")])


# ------------------------------------------------------------------ record mode + hybrid
def _sort_single(a, desc):
    got = a.copy()
    comm = mpsort.Comm.self(0)
    lib.mpsort_mpi_desc_impl(got.ctypes.data, len(got), a.dtype.itemsize, ctypes.byref(desc), comm.handle, 0, b"hybrid")
    st = C.last_stats(comm.handle, 1)
    comm.destroy()
    return got, st


@pytest.mark.parametrize("keyfield,signed", [("a", 0), ("b", 1)])
def test_hybrid_sort_uniform_keys(keyfield, signed):
    """>= 2^22 records with (nearly) distinct high 32 bits: four passes + run fix-up must
    give the bytes of the full stable sort (key in the low or the high half, signed or not)"""
    rng = np.random.default_rng(11 + signed)
    n = (1 << 22) + 12345
    dt = np.dtype([("a", "i8" if keyfield == "a" and signed else "u8"), ("b", "i8" if keyfield == "b" and signed else "u8")])
    a = np.zeros(n, dtype=dt)
    other = "b" if keyfield == "a" else "a"
    a[keyfield] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64).astype(dt[keyfield]) * 2 + rng.integers(0, 2, size=n).astype(dt[keyfield])
    # make sure equal high parts and fully equal keys exist: ties must stay in input order
    a[keyfield][1000:3000] = a[keyfield][0:2000]
    hi = a[keyfield][5000:9000].view("u8") & np.uint64(0xFFFFFFFF00000000)
    a[keyfield][9000:13000] = (hi | rng.integers(0, 1 << 32, size=4000, dtype=np.uint64)).view(dt[keyfield])
    a[other] = np.arange(n)
    desc = C.RadixDesc(0 if keyfield == "a" else 8, 8, 1, signed, 0)
    got, st = _sort_single(a, desc)
    assert st["record_mode"] == 1 and st["hybrid"] == 1 and st["first_sort_passes"] == 4
    assert np.array_equal(got, a[np.argsort(a[keyfield], kind="stable")])


@pytest.mark.parametrize("signed", [0, 1])
def test_record_mode_bare_8_byte_keys(signed):
    """elsize 8 = the reference's own bench-mpi workload (bare int64 items,
    bench-mpi.c:13-15): keys are carried through the passes, hybrid for >= 2^22"""
    rng = np.random.default_rng(21 + signed)
    dt = np.dtype("i8" if signed else "u8")
    for n, expect_hybrid in (((1 << 22) + 999, 1), (50001, 0)):
        a = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
        a = (a * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)).view(dt).copy()
        a[100:1100] = a[0:1000]                                   # duplicates
        got = a.copy()
        comm = mpsort.Comm.self(0)
        lib.radix_sort_desc(got.ctypes.data, n, 8, ctypes.byref(C.RadixDesc(0, 8, 1, signed, 0)), 0)
        lib.mpsort_mpi_desc_impl(got.ctypes.data, n, 8, ctypes.byref(C.RadixDesc(0, 8, 1, signed, 0)), comm.handle, 0, b"bare8")
        st = C.last_stats(comm.handle, 1)
        comm.destroy()
        assert st["record_mode"] == 1 and st["hybrid"] == expect_hybrid
        assert np.array_equal(got, np.sort(a, kind="stable"))
    # multi-rank: 3 ranks of bare keys
    recs = [O.as_bytes(rng.integers(0, 1 << 62, size=m, dtype=np.uint64).view(dt)) for m in (70000, 0, 90001)]
    outs = [60000, 50000, 50001]
    desc = O.Desc(0, 8, 1, signed, 0)
    out, stats = sort_group(recs, outs, desc, C.MPSORT_DISABLE_GATHER_SORT)
    assert same(out, O.numpy_sort(recs, desc, outs))
    assert all(st["record_mode"] == 1 for st in stats)


def test_hybrid_sort_long_runs_and_fallback():
    rng = np.random.default_rng(5)
    n = 1 << 22
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    for nruns, runlen, expect_fallback in ((3, 3000, False), (100, 400, True)):
        a = np.zeros(n, dtype=dt)
        a["key"] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
        pos = rng.permutation(n)[: nruns * runlen].reshape(nruns, runlen)
        for r in range(nruns):      # runs that share the high 32 bits but differ (with repeats) below
            a["key"][pos[r]] = (np.uint64(r + 1) << np.uint64(40)) | rng.integers(0, 1000, size=runlen).astype(np.uint64)
        a["tag"] = np.arange(n)
        got, st = _sort_single(a, C.RadixDesc(0, 8, 1, 0, 0))
        assert st["hybrid"] == 1 and st["hybrid_long_runs"] == nruns
        assert (st["first_sort_passes"] > 4) == expect_fallback
        assert np.array_equal(got, a[np.argsort(a["key"], kind="stable")])


def test_hybrid_declined_for_clustered_keys():
    """5 % of the records share one key: the predictor must refuse the hybrid"""
    rng = np.random.default_rng(6)
    n = 1 << 22
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    a = np.zeros(n, dtype=dt)
    a["key"] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    a["key"][rng.random(n) < 0.05] = 77
    a["tag"] = np.arange(n)
    got, st = _sort_single(a, C.RadixDesc(0, 8, 1, 0, 0))
    assert st["record_mode"] == 1 and st["hybrid"] == 0 and st["first_sort_passes"] == 8
    assert np.array_equal(got, a[np.argsort(a["key"], kind="stable")])


def test_range_compression_of_narrow_key_ranges():
    """index mode, single-word keys in a narrow range far from zero (small signed ids
    vary in every byte after the sign flip): sorted relative to their minimum with
    fewer passes; same bytes as without, also across ranks (splitters see real keys)"""
    rng = np.random.default_rng(77)
    dt = np.dtype([("key", "i8"), ("tag", "u8"), ("pad", "u8")])         # 24 bytes: index mode
    n = (1 << 20) + 77
    a = np.zeros(n, dtype=dt)
    a["key"] = rng.integers(-(1 << 20), 1 << 24, size=n)
    a["tag"] = np.arange(n)
    got, st = _sort_single(a, C.RadixDesc(0, 8, 1, 1, 0))
    assert st["record_mode"] == 0 and st["rebased"] == 1 and st["first_sort_passes"] == 4
    exp = a[np.argsort(a["key"], kind="stable")]
    assert np.array_equal(got, exp)
    os.environ["MPSORT_NO_REBASE"] = "1"
    try:
        got2, st2 = _sort_single(a, C.RadixDesc(0, 8, 1, 1, 0))
    finally:
        del os.environ["MPSORT_NO_REBASE"]
    assert st2["rebased"] == 0 and st2["first_sort_passes"] == 8 and np.array_equal(got2, exp)
    # 3 ranks, different local minima
    parts = [a[:400000], a[400000:400000 + (1 << 20) // 2], a[400000 + (1 << 20) // 2:]]
    parts[1]["key"] += 5000000
    recs = [O.as_bytes(x) for x in parts]
    outs = [len(x) for x in parts]
    out, stats = sort_group(recs, outs, O.Desc(0, 8, 1, 1, 0), C.MPSORT_DISABLE_GATHER_SORT)
    assert same(out, O.numpy_sort(recs, O.Desc(0, 8, 1, 1, 0), outs))


def test_record_mode_equals_index_mode():
    rng = np.random.default_rng(8)
    dt = np.dtype([("tag", "u8"), ("key", "i8")])
    for n in (1, 3071, 3072, 3073, 100001):
        a = np.zeros(n, dtype=dt)
        a["key"] = rng.integers(-1000, 1000, size=n)
        a["tag"] = np.arange(n)
        got, st = _sort_single(a, C.RadixDesc(8, 8, 1, 1, 0))
        assert st["record_mode"] == 1
        os.environ["MPSORT_NO_REC16"] = "1"
        try:
            got2, st2 = _sort_single(a, C.RadixDesc(8, 8, 1, 1, 0))
        finally:
            del os.environ["MPSORT_NO_REC16"]
        assert st2["record_mode"] == 0
        exp = a[np.argsort(a["key"], kind="stable")]
        assert np.array_equal(got, exp) and np.array_equal(got2, exp)


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tuning", TUNINGS)
def test_golden_vectors(name, tuning):
    """the reference's own fixtures (issue7: 12 ranks / 40-byte records / 16-byte key,
    mismatched zeros, struct) and reference-generated ones (ties, synthetic kinds)"""
    g = load_golden(name)
    out, stats = sort_group(g["recs"], g["outsizes"], g["desc"], tuning)
    assert same(out, g["exp"])


def test_golden_issue7_inplace():
    g = load_golden("issue7")
    out, _ = sort_group(g["recs"], g["outsizes"], g["desc"], C.MPSORT_DISABLE_GATHER_SORT, inplace=True)
    assert same(out, g["exp"])


def test_few_items_golden():
    """test_mpsort.py:330-351: all 81 size combinations x 5 tunings on 4 ranks"""
    z = np.load(os.path.join(GOLDEN, "few_items.npz"))
    desc = C.RadixDesc(*[int(v) for v in z["desc"]][:4], 0)
    dt = np.dtype([("vkey", ("u8", 3)), ("vector", ("u4", 3))])
    cases, off = [], 0
    for sizes in z["sizes"]:
        n = int(sum(sizes))
        cases.append(([int(s) for s in sizes], z["expected"][off:off + n]))
        off += n
    failures = []
    bar = threading.Barrier(4)

    def work(comm):
        r = comm.rank
        for tuning in TUNINGS:
            for sizes, exp in cases:
                s = np.empty(sizes[r], dtype=dt)
                s["vkey"] = np.array(range(sizes[r]), dtype="u8")[:, None]
                s["vector"] = 1
                out = np.empty(sizes[r], dtype=dt)
                if r == 0:
                    lib.mpsort_mpi_unset_options(-1)
                    if tuning:
                        lib.mpsort_mpi_set_options(tuning)
                bar.wait()
                lib.mpsort_mpi_newarray_desc_impl(s.ctypes.data, len(s), out.ctypes.data, len(out), dt.itemsize,
                                                  ctypes.byref(desc), comm.handle, 0, b"few_items")
                lo = sum(sizes[:r])
                if not np.array_equal(O.as_bytes(out), exp[lo:lo + sizes[r]]):
                    failures.append((sizes, tuning, r))
                bar.wait()

    mpsort.run_local(4, work)
    lib.mpsort_mpi_unset_options(-1)
    assert not failures, failures[:5]


# ------------------------------------------------------------------ randomized parity
@pytest.mark.parametrize("p", [2, 3, 4, 8, 12])
@pytest.mark.parametrize("distinct", [None, 3])
def test_distributed_sort_matches_oracle_and_reference_layout(p, distinct):
    """output bit-exact vs the oracle; SendCount rows equal the reference algorithm's
    (mpsort-mpi.c:483-485) -- ties split by (source rank, source index)"""
    rng = np.random.default_rng(p * 17 + (distinct or 0))
    desc = O.Desc(0, 8, 1, 0, 0)
    for trial in range(3):
        sizes = [int(rng.integers(0, 30000)) for _ in range(p)]
        if trial == 1:
            sizes[int(rng.integers(0, p))] = 0
        total = sum(sizes)
        cuts = sorted(int(c) for c in rng.integers(0, total + 1, size=p - 1))
        outsizes = [b - a for a, b in zip([0] + cuts, cuts + [total])]
        recs = []
        for r in range(p):
            a = np.zeros(sizes[r], dtype=[("key", "u8"), ("tag", "u8")])
            a["key"] = rng.integers(0, distinct, size=sizes[r]) if distinct else rng.integers(0, 1 << 63, size=sizes[r])
            a["tag"] = (r << 40) + np.arange(sizes[r])
            recs.append(O.as_bytes(a))
        exp, info = O.c_sort(recs, desc, outsizes, O.DISABLE_GATHER_SORT)
        out, stats = sort_group(recs, outsizes, desc, C.MPSORT_DISABLE_GATHER_SORT)
        assert same(out, exp)
        if info["nleaders"] == p:
            for r in range(p):
                assert stats[r]["sendcounts"] == list(info["sendcounts"][r])


@pytest.mark.parametrize("desc,elsize", [(O.Desc(0, 8, 2, 0, 0), 40), (O.Desc(8, 8, 3, 1, 0), 36),
                                          (O.Desc(0, 4, 1, 1, 0), 8), (O.Desc(0, 8, 1, 1, 0), 48),
                                          (O.Desc(4, 4, 3, 0, 0), 16)])
def test_distributed_sort_other_key_shapes(desc, elsize):
    rng = np.random.default_rng(elsize)
    for p in (2, 4):
        sizes = [int(rng.integers(0, 5000)) for _ in range(p)]
        recs = [rng.integers(0, 256, size=(n, elsize), dtype=np.uint8) for n in sizes]
        for r in recs:          # make the most significant word collide often
            lo = desc.offset + (desc.nwords - 1) * desc.width
            r[:, lo + 1:lo + desc.width] = 0
            r[:, lo] %= 3
        exp = O.numpy_sort(recs, desc, sizes)
        for tuning in (0, C.MPSORT_DISABLE_GATHER_SORT, C.MPSORT_REQUIRE_GATHER_SORT):
            out, _ = sort_group(recs, sizes, desc, tuning)
            assert same(out, exp)


@pytest.mark.parametrize("p,elsize,desc,distinct", [
    (2, 16, O.Desc(0, 8, 1, 0, 0), None), (3, 16, O.Desc(0, 8, 1, 0, 0), 7), (8, 16, O.Desc(8, 8, 1, 1, 0), None),
    (4, 48, O.Desc(0, 8, 1, 1, 0), 100), (5, 12, O.Desc(4, 4, 1, 0, 0), None), (12, 24, O.Desc(8, 8, 1, 0, 0), 2),
    (4, 7, O.Desc(1, 2, 1, 0, 1), None)])
def test_second_sort_merge_path(p, elsize, desc, distinct):
    """SecondSort as a stable p-way merge (replaces the second radix_sort,
    mpsort-mpi.c:597): bit-exact vs the oracle incl. ties across runs"""
    rng = np.random.default_rng(p * 1000 + elsize)
    sizes = [int(rng.integers(40000, 90000)) for _ in range(p)]
    sizes[int(rng.integers(0, p))] = 0
    total = sum(sizes)
    outsizes = [total // p] * p
    outsizes[-1] += total - sum(outsizes)
    recs = []
    for r in range(p):
        a = rng.integers(0, 256, size=(sizes[r], elsize), dtype=np.uint8)
        if distinct:
            lo, hi = desc.offset, desc.offset + desc.width * desc.nwords
            a[:, lo:hi] = 0
            a[:, lo] = rng.integers(0, distinct, size=sizes[r])
        recs.append(a)
    exp = O.numpy_sort(recs, desc, outsizes)
    out, stats = sort_group(recs, outsizes, desc, C.MPSORT_DISABLE_GATHER_SORT)
    assert same(out, exp)
    assert all(st["second_sort_merge_tiles"] > 0 for st in stats), "the merge path did not run"
    # every rank that keeps records merged them from its send buffer (no self copy); with the copy: the same bytes
    assert [st["own_slices_in_place"] for st in stats] == [int(st["sendcounts"][r] > 0) for r, st in enumerate(stats)]
    os.environ["MPSORT_NO_SELF_IN_PLACE"] = "1"
    try:
        out1, stats1 = sort_group(recs, outsizes, desc, C.MPSORT_DISABLE_GATHER_SORT)
    finally:
        del os.environ["MPSORT_NO_SELF_IN_PLACE"]
    assert same(out1, exp) and all(st["own_slices_in_place"] == 0 and st["second_sort_merge_tiles"] > 0 for st in stats1)
    # and the radix SecondSort gives the same bytes
    os.environ["MPSORT_NO_MERGE"] = "1"
    try:
        out2, stats2 = sort_group(recs, outsizes, desc, C.MPSORT_DISABLE_GATHER_SORT)
    finally:
        del os.environ["MPSORT_NO_MERGE"]
    assert same(out2, exp) and all(st["second_sort_merge_tiles"] == 0 for st in stats2)


def test_all_empty_and_single_rank():
    desc = O.Desc(0, 8, 1, 0, 0)
    for tuning in TUNINGS:
        out, _ = sort_group([np.zeros((0, 16), np.uint8)] * 4, [0, 0, 0, 0], desc, tuning)
        assert all(len(o) == 0 for o in out)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(1000, 16), dtype=np.uint8)
    out, _ = sort_group([a], [1000], desc)
    assert same(out, O.numpy_sort([a], desc))


# ------------------------------------------------------------------ device pointers
def test_device_pointers_inplace_and_newarray():
    comm = mpsort.Comm.self(0)
    d = C.RadixDesc(0, 8, 1, 0, 0)
    n, E = 300001, 16
    host = O.generate(n, E, 0, 42, 0, 1)
    exp = O.numpy_sort([host], O.Desc(0, 8, 1, 0, 0))[0]
    din = lib.mpsort_util_dev_malloc(0, n * E)
    dout = lib.mpsort_util_dev_malloc(0, n * E)
    lib.mpsort_util_memcpy(0, din, host.ctypes.data, n * E)
    lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, ctypes.byref(d), comm.handle, 0, b"t")
    got = np.zeros_like(host)
    lib.mpsort_util_memcpy(0, got.ctypes.data, dout, n * E)
    assert np.array_equal(got, exp)
    lib.mpsort_util_memcpy(0, din, host.ctypes.data, n * E)
    lib.mpsort_mpi_desc_impl(din, n, E, ctypes.byref(d), comm.handle, 0, b"t")
    lib.mpsort_util_memcpy(0, got.ctypes.data, din, n * E)
    assert np.array_equal(got, exp)
    lib.mpsort_util_dev_free(0, din)
    lib.mpsort_util_dev_free(0, dout)
    comm.destroy()


def test_device_generator_and_checksum_match_oracle():
    comm = mpsort.Comm.self(0)
    for kind, E in ((0, 16), (1, 16), (2, 48), (3, 24)):
        n = 50001
        buf = lib.mpsort_util_dev_malloc(0, n * E)
        lib.mpsort_util_generate(comm.handle, buf, n, E, kind, 0x5EED0001)
        got = np.zeros((n, E), np.uint8)
        lib.mpsort_util_memcpy(0, got.ctypes.data, buf, n * E)
        assert np.array_equal(got, O.generate(n, E, kind, 0x5EED0001, 0, 1))
        assert lib.mpsort_util_checksum(comm.handle, buf, n * E) == O.checksum(got)
        assert lib.mpsort_util_checksum(comm.handle, ctypes.c_void_p(buf + 3), n * E - 5) == O.checksum(got.reshape(-1)[3:-2])
        lib.mpsort_util_dev_free(0, buf)
    comm.destroy()


def test_verify_checksum_option():
    a = O.generate(20000, 16, 0, 1, 0, 1)
    lib.mpsort_mpi_unset_options(-1)
    out, _ = sort_group([a[:9000], a[9000:]], [10000, 10000], O.Desc(0, 8, 1, 0, 0), C.MPSORT_VERIFY_CHECKSUM)
    assert same(out, O.numpy_sort([a[:9000], a[9000:]], O.Desc(0, 8, 1, 0, 0), [10000, 10000]))


def test_randomised_cases_against_the_oracle():
    """tests/support/hostflow_fuzz.py (the generator of the CPU host-flow tests) against the REAL library: 150 random
    cases on rank threads of one GPU -- ranks, sizes with zeros, output layouts, record and key shapes, key
    distributions, options, 1-4 exchange parts on tiny inputs, every shipped host-side switch -- byte for byte
    against the oracle's contract"""
    env = dict(os.environ, FUZZ_SWITCHES="shipped")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), "20261019", "150"],
                        env=env, timeout=1200, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"FUZZ OK" in rc.stdout, rc.stdout.decode()[-4000:]


def test_host_buffers_in_chunks():
    """SURVEY 8 f4 on the GPU: host input in chunks behind which the histogram pass runs, host output leaving range by
    range beside the fix-up / the merges (tests/support/chunk_worker.py, tiny chunks), byte for byte against the oracle"""
    env = dict(os.environ, MPSORT_CHUNK_MIN_BYTES="4096", MPSORT_CHUNK_BYTES="1000000", MPSORT_EXCHANGE_PHASES="2",
               MPSORT_PHASES_MIN_RECORDS="1000")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "support", "chunk_worker.py")], env=env, timeout=600,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"CHUNK OK" in rc.stdout, rc.stdout.decode()[-4000:]


# ------------------------------------------------------------------ full size, by properties
def test_multiset_hash_matches_oracle_and_sees_what_a_byte_sum_cannot():
    """the whole-record multiset hash of the property checks: equal to the numpy restatement,
    invariant under any reordering of records, and changed by a swap of two payloads between
    keys or of two bytes inside a record -- both of which keep the reference's signed-byte
    sum (mpsort-mpi.c:148-159)"""
    comm = mpsort.Comm.self(0)
    rng = np.random.default_rng(3)
    for E in (16, 48, 8, 24, 7, 19):
        a = rng.integers(0, 256, size=(50001, E), dtype=np.uint8)
        h = C.multiset_hash(comm.handle, a.ctypes.data, len(a), E)
        assert h == O.multiset_hash(a)
        b = a[rng.permutation(len(a))].copy()
        assert C.multiset_hash(comm.handle, b.ctypes.data, len(b), E) == h
        c = a.copy()
        if E >= 16:
            c[[10, 20], 8:] = c[[20, 10], 8:]                   # payloads swapped between two keys
        else:
            c[10, [0, E - 1]] = c[10, [E - 1, 0]]                # two bytes swapped inside a record
        if not np.array_equal(c, a):
            assert O.checksum(c) == O.checksum(a)
            assert C.multiset_hash(comm.handle, c.ctypes.data, len(c), E) != h
    # device pointers, unaligned base
    buf = lib.mpsort_util_dev_malloc(0, 16 * 1000 + 8)
    a = rng.integers(0, 256, size=(1000, 16), dtype=np.uint8)
    lib.mpsort_util_memcpy(0, ctypes.c_void_p(buf + 8), a.ctypes.data, a.nbytes)
    assert C.multiset_hash(comm.handle, ctypes.c_void_p(buf + 8), 1000, 16) == O.multiset_hash(a)
    lib.mpsort_util_dev_free(0, buf)
    comm.destroy()


def _property_check(comm, n, E, kind, desc, seed):
    buf = lib.mpsort_util_dev_malloc(0, n * E)
    lib.mpsort_util_generate(comm.handle, buf, n, E, kind, seed)
    s1 = C.multiset_hash(comm.handle, buf, n, E)
    lib.mpsort_mpi_desc_impl(buf, n, E, ctypes.byref(desc), comm.handle, 0, b"fullsize")
    s2 = C.multiset_hash(comm.handle, buf, n, E)
    fl = (ctypes.c_uint64 * 2)()
    bad = lib.mpsort_util_check_sorted(comm.handle, buf, n, E, ctypes.byref(desc), 1, 8, fl)
    # idempotence: sorting the sorted array changes nothing (stable)
    lib.mpsort_mpi_desc_impl(buf, n, E, ctypes.byref(desc), comm.handle, 0, b"fullsize")
    s3 = C.multiset_hash(comm.handle, buf, n, E)
    bad2 = lib.mpsort_util_check_sorted(comm.handle, buf, n, E, ctypes.byref(desc), 1, 8, fl)
    lib.mpsort_util_dev_free(0, buf)
    assert s1 == s2 == s3, "the multiset of records changed"
    assert bad == 0 and bad2 == 0, "order or tie order violated"


def _combine(hashes):
    """multiset hashes of the ranks' parts -> hash of their union: sums add, xors xor"""
    s = x = 0
    for a, b in hashes:
        s = (s + a) & ((1 << 64) - 1)
        x ^= b
    return s, x


def test_full_size_config_b_by_properties():
    """BASELINE.json configs[1]: 2^28 uniform u64-keyed 16-byte records on one B200"""
    comm = mpsort.Comm.self(0)
    _property_check(comm, 1 << 28, 16, 0, C.RadixDesc(0, 8, 1, 0, 0), 0x5EED0001)
    comm.destroy()


def _full_size_against_the_reference(tmp_path, log2n, E, kind):
    """the unmodified reference (oracle/_ref/bench16 -> mpsort_mpi_newarray, one MPI-shim rank per host core) and
    one B200 sort the same 2^log2n records. The reference's R ranks generate R chunks; the GPU holds their
    rank-order concatenation, whose stable sort is the reference's output contract. Returns the GPU sort's statistics."""
    cores = os.cpu_count() or 1
    R = 1
    while R * 2 <= min(cores, 32):
        R *= 2
    n, per = 1 << log2n, (1 << log2n) // R
    comm = mpsort.Comm.self(0)
    buf = lib.mpsort_util_dev_malloc(0, n * E)
    for r in range(R):
        lib.mpsort_util_generate_as(comm.handle, ctypes.c_void_p(buf + r * per * E), per, E, kind, 0x5EED0001, r, R)
    lib.mpsort_mpi_desc_impl(buf, n, E, ctypes.byref(C.RadixDesc(0, 8, 1, 1 if kind == 2 else 0, 0)), comm.handle, 0, b"fullsize_ref")
    st = C.last_stats(comm.handle, 1)
    res = O.run_bench16(R, per, elsize=E, kind=kind, reps=1, timeout=1500, outdir=str(tmp_path))
    got = np.empty((per, E), np.uint8)
    for r in range(R):
        lib.mpsort_util_memcpy(0, got.ctypes.data, ctypes.c_void_p(buf + r * per * E), per * E)
        exp = np.fromfile(str(tmp_path / ("out.%d" % r)), dtype=np.uint8).reshape(per, E)
        assert np.array_equal(got, exp), "chunk %d of %d differs from the reference's output" % (r, R)
        os.unlink(str(tmp_path / ("out.%d" % r)))
    lib.mpsort_util_dev_free(0, buf)
    comm.destroy()
    print("reference: %d ranks, %.1f s per sort" % (R, res["best_seconds"]))
    return st


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not in this tree")
def test_full_size_config_b_bytes_equal_the_reference(tmp_path):
    """BASELINE.json configs[1] at FULL size, byte for byte: 2^28 uniform 16-byte records (record mode + hybrid)"""
    st = _full_size_against_the_reference(tmp_path, 28, 16, 0)
    assert st["hybrid"] == 1 and st["record_mode"] == 1


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not in this tree")
@pytest.mark.parametrize("log2n,E,kind", [(28, 16, 1), (27, 48, 2)], ids=["mostly_sorted16", "particles48"])
def test_full_size_configs_4_and_5_bytes_equal_the_reference(tmp_path, log2n, E, kind):
    """BASELINE.json configs[3] and [4] on one GPU at full per-GPU size, byte for byte against the unmodified
    reference: 2^28 mostly sorted 16-byte records (the predictor's choice of depth, runs of equal high parts fixed up
    in place) and 2^27 48-byte particles with skewed signed ids and 5 % duplicates (index mode, range compression,
    ties in input order, the payload gather)"""
    st = _full_size_against_the_reference(tmp_path, log2n, E, kind)
    assert st["record_mode"] == (1 if E == 16 else 0)


def test_large_particles48_and_mostly_sorted_by_properties():
    comm = mpsort.Comm.self(0)
    _property_check(comm, 1 << 26, 48, 2, C.RadixDesc(0, 8, 1, 1, 0), 7)     # heavy duplicates: tie order by tag
    _property_check(comm, 1 << 27, 16, 1, C.RadixDesc(0, 8, 1, 0, 0), 9)
    comm.destroy()


@pytest.mark.parametrize("phases", [1, 4])
def test_large_multirank_hybrid_records_by_properties(phases, monkeypatch):
    """4 rank threads x 2^22 uniform 16-byte records: record mode + hybrid FirstSort,
    exchange (optionally pipelined in 4 parts over virtual ranks), merge; checked by
    global order, tie order and checksum of checksums"""
    monkeypatch.setenv("MPSORT_EXCHANGE_PHASES", str(phases))
    p, n, E = 4, 1 << 22, 16
    desc = C.RadixDesc(0, 8, 1, 0, 0)
    res = [None] * p

    def work(comm):
        r = comm.rank
        buf = lib.mpsort_util_dev_malloc(0, n * E)
        out = lib.mpsort_util_dev_malloc(0, n * E)
        lib.mpsort_util_generate(comm.handle, buf, n, E, 0, 0x5EED0001)
        s1 = C.multiset_hash(comm.handle, buf, n, E)
        lib.mpsort_mpi_newarray_desc_impl(buf, n, out, n, E, ctypes.byref(desc), comm.handle, 0, b"multirank16")
        st = C.last_stats(comm.handle, p)
        s2 = C.multiset_hash(comm.handle, out, n, E)
        fl = (ctypes.c_uint64 * 2)()
        bad = lib.mpsort_util_check_sorted(comm.handle, out, n, E, ctypes.byref(desc), 1, 8, fl)
        lib.mpsort_util_dev_free(0, buf)
        lib.mpsort_util_dev_free(0, out)
        res[r] = (s1, s2, bad, fl[0], fl[1], st)

    mpsort.run_local(p, work)
    assert _combine([x[0] for x in res]) == _combine([x[1] for x in res]), "the multiset of records changed"
    assert all(x[2] == 0 for x in res)
    assert all(res[r - 1][4] <= res[r][3] for r in range(1, p))
    assert all(x[5]["record_mode"] == 1 and x[5]["hybrid"] == 1 and x[5]["second_sort_merge_tiles"] > 0 for x in res)
    assert all(x[5]["exchange_phases"] == phases for x in res)


def test_large_multirank_by_properties():
    """4 rank threads x 2^22 records with a 5 % equal-key run spanning ranks: global
    order, tie order across rank boundaries, checksum of checksums"""
    p, n, E = 4, 1 << 22, 48
    desc = C.RadixDesc(0, 8, 1, 1, 0)
    res = [None] * p

    def work(comm):
        r = comm.rank
        buf = lib.mpsort_util_dev_malloc(0, n * E)
        lib.mpsort_util_generate(comm.handle, buf, n, E, 2, 0x5EED0001)
        s1 = C.multiset_hash(comm.handle, buf, n, E)
        lib.mpsort_mpi_desc_impl(buf, n, E, ctypes.byref(desc), comm.handle, 0, b"multirank")
        s2 = C.multiset_hash(comm.handle, buf, n, E)
        fl = (ctypes.c_uint64 * 2)()
        bad = lib.mpsort_util_check_sorted(comm.handle, buf, n, E, ctypes.byref(desc), 1, 8, fl)
        first = np.zeros((1, E), np.uint8)
        last = np.zeros((1, E), np.uint8)
        lib.mpsort_util_memcpy(0, first.ctypes.data, buf, E)
        lib.mpsort_util_memcpy(0, last.ctypes.data, ctypes.c_void_p(buf + (n - 1) * E), E)
        lib.mpsort_util_dev_free(0, buf)
        res[r] = (s1, s2, bad, first, last)
        return None

    mpsort.run_local(p, work)
    assert _combine([x[0] for x in res]) == _combine([x[1] for x in res]), "the multiset of records changed"
    assert all(x[2] == 0 for x in res)
    for r in range(1, p):
        a = res[r - 1][4].view("<i8").reshape(-1)
        b = res[r][3].view("<i8").reshape(-1)
        assert (a[0], a[1]) < (b[0], b[1]), "rank boundary %d out of (key, tag) order" % r


# ------------------------------------------------------------------ NCCL, one process per GPU
@pytest.mark.skipif(lib.mpsort_util_device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
def test_nccl_process_per_gpu():
    n = min(lib.mpsort_util_device_count(), 8)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    rc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                         "--master-addr", "127.0.0.1", "--master-port", "29517",
                         os.path.join(ROOT, "tests", "nccl_worker.py")], env=env, timeout=600,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0, rc.stdout.decode()[-4000:]
    assert b"NCCL PARITY OK" in rc.stdout


@pytest.mark.skipif(lib.mpsort_util_device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref (the compiled reference) is not in this tree")
@pytest.mark.parametrize("kind", [0, 1], ids=["uniform16", "mostly_sorted16"])
def test_nccl_full_size_bytes_equal_the_reference(kind):
    """two processes, two GPUs, 2^28 16-byte records each through the production transport (IPC-mapped DMA exchange in
    parts, peer splitter kernel, own slice merged in place): every byte equal to the unmodified reference's output for the
    same 2^29 records (tests/support/nccl_fullsize_worker.py)"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519", KIND=str(kind))
    rc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                         "--master-addr", "127.0.0.1", "--master-port", "29519",
                         os.path.join(ROOT, "tests", "support", "nccl_fullsize_worker.py")], env=env, timeout=1500,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    print(rc.stdout.decode()[-1500:])
    assert rc.returncode == 0, rc.stdout.decode()[-4000:]
    assert b"NCCL FULL SIZE OK" in rc.stdout

