"""Quick on-GPU sanity + timing script (not a pytest file): run under gpurun while
developing.  python tests/gpu_quick.py [--big]"""
import ctypes
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mp-sort_b200"))
from mpsort import _capi as C  # noqa: E402

lib = C.lib


def sort_host(comm, arr, out, desc):
    lib.mpsort_mpi_newarray_desc_impl(arr.ctypes.data, len(arr), out.ctypes.data, len(out),
                                      arr.dtype.itemsize, ctypes.byref(desc), comm, 0, b"gpu_quick")


def check_single(comm, n, dtype, field, desc, rng, name):
    a = np.zeros(n, dtype=dtype)
    raw = a.view(np.uint8).reshape(n, a.dtype.itemsize)
    raw[...] = rng.integers(0, 256, size=raw.shape, dtype=np.uint8)
    if "tag" in a.dtype.names:
        a["tag"] = np.arange(n)
    out = np.zeros_like(a)
    sort_host(comm, a, out, desc)
    key = a[field]
    if key.ndim == 2:
        order = np.lexsort(tuple(key[:, i] for i in range(key.shape[1])))
    else:
        order = np.argsort(key, kind="stable")
    exp = a[order]
    ok = np.array_equal(out.view(np.uint8), exp.view(np.uint8))
    print("%-40s n=%-9d %s  %s" % (name, n, "OK" if ok else "MISMATCH", C.last_stats(comm, 1)["first_sort_passes"]))
    if not ok:
        E = a.dtype.itemsize
        bad = np.nonzero((out.view(np.uint8).reshape(n, E) != exp.view(np.uint8).reshape(n, E)).any(axis=1))[0]
        print("   first bad rows:", bad[:10], "of", len(bad))
    return ok


def run_group(p, arrays, outs, desc, tuning=0):
    devs = (ctypes.c_int * p)(*([0] * p))
    comms = (ctypes.c_void_p * p)()
    assert lib.mpsort_comm_init_local_group(p, devs, comms) == 0
    lib.mpsort_mpi_unset_options(-1)
    if tuning:
        lib.mpsort_mpi_set_options(tuning)
    errs = []

    def work(r):
        try:
            sort_host(comms[r], arrays[r], outs[r], desc)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(p)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    stats = [C.last_stats(comms[r], p) for r in range(p)]
    for r in range(p):
        lib.mpsort_comm_destroy(comms[r])
    assert not errs, errs
    return stats


def check_group(p, sizes, outsizes, rng, name, dup=False, tuning=0):
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    arrays = []
    for r in range(p):
        a = np.zeros(sizes[r], dtype=dt)
        if dup:
            a["key"] = rng.integers(0, 7, size=sizes[r])
        else:
            a["key"] = rng.integers(0, 2 ** 63, size=sizes[r], dtype=np.uint64) * 2 + rng.integers(0, 2, size=sizes[r], dtype=np.uint64)
        a["tag"] = (r << 40) + np.arange(sizes[r])
        arrays.append(a)
    outs = [np.zeros(outsizes[r], dtype=dt) for r in range(p)]
    desc = C.RadixDesc(0, 8, 1, 0, 0)
    stats = run_group(p, arrays, outs, desc, tuning)
    allin = np.concatenate(arrays)
    exp = allin[np.argsort(allin["key"], kind="stable")]
    got = np.concatenate(outs)
    ok = np.array_equal(got.view(np.uint8), exp.view(np.uint8))
    print("%-40s p=%d sizes=%s %s gather=%d rounds=%d" % (name, p, sizes if len(sizes) < 6 else "...", "OK" if ok else "MISMATCH",
                                                          stats[0]["used_gather"], stats[0]["splitter_rounds"]))
    return ok


def main():
    rng = np.random.default_rng(1234)
    ndev = lib.mpsort_util_device_count()
    print("devices:", ndev)
    comm = lib.mpsort_comm_self(0)
    allok = True
    dt16 = np.dtype([("key", "u8"), ("tag", "u8")])
    d16 = C.RadixDesc(0, 8, 1, 0, 0)
    for n in [0, 1, 2, 31, 33, 1000, 6143, 6144, 6145, 100000, 1 << 20, (1 << 22) + 7]:
        allok &= check_single(comm, n, dt16, "key", d16, rng, "u64 key 16B records")
    dti = np.dtype([("tag", "u8"), ("key", "i8")])
    allok &= check_single(comm, 200000, dti, "key", C.RadixDesc(8, 8, 1, 1, 0), rng, "i64 key at offset 8")
    dt48 = np.dtype([("key", "i8"), ("tag", "u8"), ("pos", "f8", 2), ("vel", "f4", 3), ("pad", "u4")])
    assert dt48.itemsize == 48
    allok &= check_single(comm, 300001, dt48, "key", C.RadixDesc(0, 8, 1, 1, 0), rng, "48B records i64 key")
    dt4 = np.dtype([("key", "u4"), ("x", "u4"), ("tag", "u8")])
    allok &= check_single(comm, 123457, dt4, "key", C.RadixDesc(0, 4, 1, 0, 0), rng, "u4 key")
    dti4 = np.dtype([("x", "u4"), ("key", "i4"), ("tag", "u8")])
    allok &= check_single(comm, 123457, dti4, "key", C.RadixDesc(4, 4, 1, 1, 0), rng, "i4 key at offset 4")
    dt40 = np.dtype([("radix", "u8", 2), ("tag", "u8"), ("ext", "u8", 2)])
    allok &= check_single(comm, 77777, dt40, "radix", C.RadixDesc(0, 8, 2, 0, 0), rng, "2-word key 40B records")
    dt3 = np.dtype([("vkey", "u8", 3), ("tag", "u8"), ("vector", "u4", 2)])
    allok &= check_single(comm, 5000, dt3, "vkey", C.RadixDesc(0, 8, 3, 0, 0), rng, "3-word key")
    dt12 = np.dtype([("key", "u4"), ("a", "u4"), ("b", "u4")])
    allok &= check_single(comm, 10001, dt12, "key", C.RadixDesc(0, 4, 1, 0, 0), rng, "12B records (4B aligned)")
    dt7 = np.dtype([("key", "u2"), ("a", "u1", 5)])
    allok &= check_single(comm, 10001, dt7, "key", C.RadixDesc(0, 2, 1, 0, 0), rng, "7B records u2 key")

    # duplicates: few distinct keys, tags make ties visible
    a = np.zeros(500000, dtype=dt16)
    a["key"] = rng.integers(0, 5, size=len(a))
    a["tag"] = np.arange(len(a))
    out = np.zeros_like(a)
    sort_host(comm, a, out, d16)
    exp = a[np.argsort(a["key"], kind="stable")]
    ok = np.array_equal(out.view(np.uint8), exp.view(np.uint8))
    print("%-40s %s" % ("5 distinct keys, stable ties", "OK" if ok else "MISMATCH"))
    allok &= ok

    # multi-rank, in-process group on one device
    for p in [2, 3, 4]:
        allok &= check_group(p, [100000 + 17 * r for r in range(p)], [100000 + 17 * r for r in range(p)], rng, "group uniform")
        allok &= check_group(p, [100000] * p, [100000] * p, rng, "group duplicates", dup=True)
    allok &= check_group(4, [0, 400000, 0, 600000], [200000, 200000, 0, 600000], rng, "group mismatched zeros")
    allok &= check_group(4, [0, 400, 0, 600], [200, 200, 0, 600], rng, "small -> gather path")
    allok &= check_group(4, [0, 400, 0, 600], [200, 200, 0, 600], rng, "small, DISABLE_GATHER", tuning=C.MPSORT_DISABLE_GATHER_SORT)
    allok &= check_group(4, [3, 0, 2, 1], [1, 2, 0, 3], rng, "tiny dup DISABLE_GATHER", dup=True, tuning=C.MPSORT_DISABLE_GATHER_SORT)
    allok &= check_group(12, [0, 0, 2, 7, 1, 0, 4, 0, 0, 6, 14, 5], [0, 0, 2, 7, 1, 0, 4, 0, 0, 6, 14, 5], rng, "12 ranks tiny",
                         tuning=C.MPSORT_DISABLE_GATHER_SORT)
    allok &= check_group(8, [50000] * 8, [50000] * 8, rng, "8 ranks dense", tuning=C.MPSORT_DISABLE_SPARSE_ALLTOALLV)

    print("ALL OK" if allok else "FAILURES")

    if "--big" in sys.argv:
        n = 1 << 28
        E = 16
        din = lib.mpsort_util_dev_malloc(0, n * E)
        dout = lib.mpsort_util_dev_malloc(0, n * E)
        for it in range(4):
            lib.mpsort_util_generate(comm, din, n, E, 0, 0x5EED0001 + it)
            t0 = time.time()
            lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, ctypes.byref(d16), comm, 0, b"big")
            dt = time.time() - t0
            fl = (ctypes.c_uint64 * 2)()
            viol = lib.mpsort_util_check_sorted(comm, dout, n, E, ctypes.byref(d16), 1, 8, fl)
            print("big it=%d %.2f ms  %.2f Grec/s  violations=%d  phases=%s" % (
                it, dt * 1e3, n / dt / 1e9, viol, [(k, round(v * 1e3, 3)) for k, v in C.last_run()]))
    lib.mpsort_comm_destroy(comm)
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
