"""CPU-only checks of the product's host side: the C-ABI exports what include/*.h
declares, option state, the layout solver against the oracle's (= the reference's)
SendCount matrix, key range / start level, and the Python-side argument checks of
the binding (same exceptions as the reference's binding.pyx). No compute call needs
a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

import mpsort_oracle as O
from conftest import ROOT

import mpsort
from mpsort import _capi as C

lib = C.lib


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if not n.startswith("mpiu_malloc_func") and n not in ("defined", "void", "int", "sizeof")))


@pytest.mark.parametrize("header", ["mpsort.h", "mpsort_util.h"])
def test_cabi_exports_every_declared_symbol(header):
    names = declared_functions(header)
    assert len(names) >= 10
    dll = ctypes.CDLL(C.LIB_PATH)
    missing = [n for n in names if n not in ("mpiu_malloc_func", "mpiu_free_func") and not hasattr(dll, n)]
    assert not missing, "declared in include/%s but not exported: %s" % (header, missing)


def test_reference_symbol_names_and_option_bits():
    """same names / bit values as reference mpsort.h:17-24"""
    assert C.MPSORT_DISABLE_SPARSE_ALLTOALLV == 2 and C.MPSORT_DISABLE_GATHER_SORT == 8
    assert C.MPSORT_REQUIRE_GATHER_SORT == 16 and C.MPSORT_REQUIRE_SPARSE_ALLTOALLV == 64
    lib.mpsort_mpi_unset_options(-1)
    assert lib.mpsort_mpi_has_options(-1) == 0
    lib.mpsort_mpi_set_options(C.MPSORT_DISABLE_GATHER_SORT | C.MPSORT_REQUIRE_SPARSE_ALLTOALLV)
    assert lib.mpsort_mpi_has_options(C.MPSORT_DISABLE_GATHER_SORT)
    assert lib.mpsort_mpi_has_options(C.MPSORT_REQUIRE_SPARSE_ALLTOALLV)
    assert not lib.mpsort_mpi_has_options(C.MPSORT_REQUIRE_GATHER_SORT)
    lib.mpsort_mpi_unset_options(C.MPSORT_DISABLE_GATHER_SORT)
    assert not lib.mpsort_mpi_has_options(C.MPSORT_DISABLE_GATHER_SORT)
    lib.mpsort_mpi_unset_options(-1)


def test_no_device_means_no_fallback():
    """on a CPU-only box the library reports zero devices; nothing sorts on the CPU"""
    if lib.mpsort_util_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        mpsort.Comm.local_group(2)


def local_counts(sorted_keys, splitters):
    clt = [int(np.searchsorted(sorted_keys, s, side="left")) for s in splitters]
    cle = [int(np.searchsorted(sorted_keys, s, side="right")) for s in splitters]
    return clt, cle


@pytest.mark.parametrize("p,distinct", [(2, None), (4, None), (4, 3), (8, 5), (12, 2), (5, 1)])
def test_layout_solver_matches_reference_sendcounts(p, distinct):
    """mpsort_solve_layout (product host code) vs the SendCount matrix of the oracle's
    restatement of _solve_for_layout_mpi (mpsort-mpi.c:663-727)"""
    rng = np.random.default_rng(1000 + p)
    desc = O.Desc(0, 8, 1, 0, 0)
    for trial in range(5):
        sizes = [int(rng.integers(0, 400)) for _ in range(p)]
        if trial == 1:
            sizes[0] = 0
        total = sum(sizes)
        cuts = sorted(int(c) for c in rng.integers(0, total + 1, size=p - 1))
        outsizes = [b - a for a, b in zip([0] + cuts, cuts + [total])]
        recs = []
        for r in range(p):
            a = np.zeros(sizes[r], dtype=[("key", "u8"), ("tag", "u8")])
            a["key"] = rng.integers(0, distinct, size=sizes[r]) if distinct else rng.integers(0, 1 << 62, size=sizes[r])
            a["tag"] = (r << 40) + np.arange(sizes[r])
            recs.append(O.as_bytes(a))
        _, info = O.c_sort(recs, desc, outsizes, O.DISABLE_GATHER_SORT)
        assert info["nleaders"] == p or 0 in [s + o for s, o in zip(sizes, outsizes)]
        if info["nleaders"] != p:
            continue
        # what the device computes: splitter b = key at global rank C[b]-1; local counts
        keys = [np.sort(r[:, :8].copy().view("<u8").reshape(-1)) for r in recs]
        allk = np.sort(np.concatenate(keys))
        Cc = (ctypes.c_int64 * (p + 1))()
        lib.mpsort_cumulative_counts(p, (ctypes.c_int64 * p)(*outsizes), Cc)
        assert list(Cc) == [0] + list(np.cumsum(outsizes))
        splitters = [allk[Cc[b] - 1] if Cc[b] > 0 else (allk[0] if total else 0) for b in range(1, p)]
        clt = (ctypes.c_int64 * (p * (p - 1)))()
        cle = (ctypes.c_int64 * (p * (p - 1)))()
        for j in range(p):
            a, b = local_counts(keys[j], splitters)
            for i in range(p - 1):
                clt[j * (p - 1) + i] = a[i]
                cle[j * (p - 1) + i] = b[i]
        cut = (ctypes.c_int64 * (p * (p + 1)))()
        rc = lib.mpsort_solve_layout(p, Cc, clt, cle, (ctypes.c_int64 * p)(*sizes), cut)
        assert rc == 0
        sc = np.array([[cut[j * (p + 1) + k + 1] - cut[j * (p + 1) + k] for k in range(p)] for j in range(p)])
        assert np.array_equal(sc, info["sendcounts"])


@pytest.mark.parametrize("p,Q,distinct", [(2, 2, None), (4, 2, 3), (8, 2, None), (3, 4, 2), (8, 4, 1), (5, 3, None)])
def test_layout_for_pipelined_exchange_refines_the_reference_layout(p, Q, distinct):
    """The exchange runs in Q parts per rank (default 2): the layout is solved for p sources x
    p*Q virtual destinations (mpsort_solve_layout2). Taking every Q-th cut must give exactly
    the reference layout (mpsort-mpi.c:663-727), and the slices of one virtual destination,
    concatenated in source order and stably merged, must be that part of the global stable sort."""
    c_i64 = ctypes.c_int64
    lib.mpsort_solve_layout2.restype = ctypes.c_int
    lib.mpsort_solve_layout2.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(c_i64)] * 5
    rng = np.random.default_rng(77 * p + Q)
    for trial in range(4):
        sizes = [int(rng.integers(0, 300)) for _ in range(p)]
        if trial == 1:
            sizes[p - 1] = 0
        total = sum(sizes)
        cuts = sorted(int(c) for c in rng.integers(0, total + 1, size=p - 1))
        outsizes = [b - a for a, b in zip([0] + cuts, cuts + [total])]
        keys, tags = [], []
        for r in range(p):
            k = rng.integers(0, distinct, size=sizes[r]) if distinct else rng.integers(0, 1 << 62, size=sizes[r])
            order = np.argsort(k, kind="stable")
            keys.append(np.asarray(k, dtype="u8")[order])
            tags.append(((r << 40) + np.arange(sizes[r]))[order])          # (source rank, source index)
        allk = np.sort(np.concatenate(keys)) if total else np.zeros(0, dtype="u8")
        pv = p * Q
        Cv = [0]
        for j in range(p):
            for b in range(Q):
                Cv.append(Cv[-1] + (outsizes[j] * (b + 1) // Q - outsizes[j] * b // Q))
        assert Cv[-1] == total and [Cv[j * Q] for j in range(p + 1)] == [0] + list(np.cumsum(outsizes))

        def solve(C, nd):
            splitters = [allk[C[b] - 1] if C[b] > 0 else (allk[0] if total else 0) for b in range(1, nd)]
            clt = (c_i64 * (p * (nd - 1)))()
            cle = (c_i64 * (p * (nd - 1)))()
            for j in range(p):
                a, b = local_counts(keys[j], splitters)
                for i in range(nd - 1):
                    clt[j * (nd - 1) + i] = a[i]
                    cle[j * (nd - 1) + i] = b[i]
            cut = (c_i64 * (p * (nd + 1)))()
            rc = lib.mpsort_solve_layout2(p, nd, (c_i64 * (nd + 1))(*C), clt, cle, (c_i64 * p)(*sizes), cut)
            assert rc == 0
            return [[cut[j * (nd + 1) + k] for k in range(nd + 1)] for j in range(p)]

        fine = solve(Cv, pv)
        coarse = solve([Cv[j * Q] for j in range(p + 1)], p)
        for j in range(p):
            assert fine[j][0] == 0 and fine[j][pv] == sizes[j]
            assert all(fine[j][v] <= fine[j][v + 1] for v in range(pv))
            assert [fine[j][k * Q] for k in range(p + 1)] == coarse[j]
        # the global stable sort, ties by (source rank, source index)
        gk = np.concatenate(keys) if total else np.zeros(0, dtype="u8")
        gt = np.concatenate(tags) if total else np.zeros(0, dtype="i8")
        order = np.lexsort((gt, gk))
        for v in range(pv):
            assert sum(fine[j][v + 1] - fine[j][v] for j in range(p)) == Cv[v + 1] - Cv[v]
            part_k = np.concatenate([keys[j][fine[j][v]:fine[j][v + 1]] for j in range(p)])
            part_t = np.concatenate([tags[j][fine[j][v]:fine[j][v + 1]] for j in range(p)])
            merged = np.argsort(part_k, kind="stable")            # runs are in source order: stable merge
            assert np.array_equal(part_t[merged], gt[order][Cv[v]:Cv[v + 1]])


def test_layout_solver_reports_reference_bug_conditions():
    """the reference aborts with 'serious bug' (mpsort-mpi.c:707-716); we return codes"""
    p = 2
    Cc = (ctypes.c_int64 * 3)(0, 1, 4)
    nm = (ctypes.c_int64 * 2)(2, 2)
    cut = (ctypes.c_int64 * 6)()
    clt = (ctypes.c_int64 * 2)(2, 2)      # more below the splitter than wanted
    cle = (ctypes.c_int64 * 2)(2, 2)
    assert lib.mpsort_solve_layout(p, Cc, clt, cle, nm, cut) == -1
    clt = (ctypes.c_int64 * 2)(0, 0)      # not enough equal keys to fill the deficit
    cle = (ctypes.c_int64 * 2)(0, 0)
    assert lib.mpsort_solve_layout(p, Cc, clt, cle, nm, cut) == -3


def test_key_range_start_level():
    """_find_Pmax_Pmin_C (mpsort-mpi.c:606-661): empty ranks are skipped, an empty
    world gives 0; the byte-wise descent starts below the common leading bytes"""
    p, nw = 3, 1
    nm = (ctypes.c_int64 * p)(5, 0, 7)
    kmin = (ctypes.c_uint64 * p)(0x1122334455660000, 0xdeadbeef, 0x1122334455000000)
    kmax = (ctypes.c_uint64 * p)(0x11223344556600ff, 0xdeadbeef, 0x1122334455ffffff)
    Pmin = (ctypes.c_uint64 * nw)()
    Pmax = (ctypes.c_uint64 * nw)()
    prefix = (ctypes.c_uint64 * nw)()
    lvl = lib.mpsort_key_range(p, nw, nm, kmin, kmax, Pmin, Pmax, prefix)
    assert Pmin[0] == 0x1122334455000000 and Pmax[0] == 0x1122334455ffffff
    assert lvl == 5 and prefix[0] == 0x1122334455000000
    nm0 = (ctypes.c_int64 * p)(0, 0, 0)
    lvl = lib.mpsort_key_range(p, nw, nm0, kmin, kmax, Pmin, Pmax, prefix)
    assert Pmin[0] == 0 and Pmax[0] == 0 and lvl == 8


def test_binding_radix_desc_rules():
    """radix_data_init (binding.pyx:50-79): offset, width, nwords, signedness"""
    dt = np.dtype([("value", "i8"), ("key", "i8"), ("vkey", ("u4", 3)), ("f", "f8"), ("m", ("u8", (2, 2)))])
    assert mpsort.radix_desc(dt, "key") == (8, 8, 1, 1)
    assert mpsort.radix_desc(dt, "vkey") == (16, 4, 3, 0)
    assert mpsort.radix_desc(np.dtype("u8"), None) == (0, 8, 1, 0)
    assert mpsort.radix_desc(np.dtype("i4"), None) == (0, 4, 1, 1)
    with pytest.raises(TypeError):
        mpsort.radix_desc(dt, "f")
    with pytest.raises(ValueError):
        mpsort.radix_desc(dt, "m")
    with pytest.raises(ValueError):
        mpsort.radix_desc(dt, "nope")
    with pytest.raises(TypeError):
        mpsort.radix_desc(np.dtype("u2"), None)


def test_python_surface_names():
    """same public names as the reference package (mpsort/__init__.py)"""
    for name in ("sort", "permute", "take", "histogram", "globalrange", "globalindices", "guess_dtype", "__version__"):
        assert hasattr(mpsort, name)
    import inspect
    assert list(inspect.signature(mpsort.sort).parameters) == ["source", "orderby", "out", "comm", "tuning"]
