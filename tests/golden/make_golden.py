"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref,
built from /root/reference under the single-host MPI shim) on the golden inputs of
the reference's own test-suite and on a few seeded inputs that pin what that suite
does not (tie order of equal keys with distinguishable payloads).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixtures are committed; tests never read /root/reference.
"""
import base64
import os
import pickle
import re
import sys
from itertools import product

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import mpsort_oracle as O  # noqa: E402

REF_TESTS = "/root/reference/mpsort/tests/test_mpsort.py"
TUNINGS = [0, O.DISABLE_SPARSE_ALLTOALLV, O.REQUIRE_SPARSE_ALLTOALLV, O.REQUIRE_GATHER_SORT, O.DISABLE_GATHER_SORT]


def save(name, recs, desc, outsizes, expected, **extra):
    sizes = np.array([len(r) for r in recs], dtype=np.int64)
    elsize = recs[0].shape[1]
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        records=np.concatenate(recs, axis=0), sizes=sizes,
                        outsizes=np.array(outsizes, dtype=np.int64),
                        desc=np.array(desc.astuple(), dtype=np.int64), elsize=np.int64(elsize),
                        expected=np.concatenate(expected, axis=0), **extra)


def run_all_tunings(recs, desc, outsizes, tunings=TUNINGS):
    """the reference under every tuning must agree with itself; returns its output"""
    first = None
    for t in tunings:
        out = O.ref_sort(recs, desc, outsizes, t)
        if first is None:
            first = out
        else:
            assert all(np.array_equal(a, b) for a, b in zip(first, out)), "reference disagrees with itself, tuning %d" % t
    return first


def issue7():
    """test_mpsort.py:354-556: 12 ranks, 40-byte records, 16-byte two-word key"""
    src = open(REF_TESTS).read()
    b64 = re.search(r'Issue7B64 = b"""(.*?)"""', src, re.S).group(1).encode()
    A = [np.array(a, dtype="u4").reshape(-1, 10) for a in pickle.loads(base64.decodebytes(b64))]
    dt = np.dtype([("radix", ("u8", 2)), ("ext", ("u8", 3))])
    recs = []
    for a in A:
        s = np.zeros(len(a), dtype=dt)
        s["radix"][:, 0] = a[:, 5] + (a[:, 4].astype("u8") << 32)
        s["radix"][:, 1] = a[:, 0]
        recs.append(O.as_bytes(s))
    desc = O.Desc(0, 8, 2, 0, 0)
    sizes = [len(r) for r in recs]
    out = run_all_tunings(recs, desc, sizes)
    # the reference test's own assertion: equals numpy.lexsort of the radix columns
    S = np.concatenate(recs, axis=0).view(dt).reshape(-1)
    exp = S[np.lexsort(S["radix"].T)]
    assert np.array_equal(np.concatenate(out, axis=0), O.as_bytes(exp))
    save("issue7", recs, desc, sizes, out)
    print("issue7: sizes", sizes)


def few_items():
    """test_mpsort.py:330-351: all 3^4 size combinations of {0,1,2} on 4 ranks, 24-byte key"""
    dt = np.dtype([("vkey", ("u8", 3)), ("vector", ("u4", 3))])
    desc = O.Desc(0, 8, 3, 0, 0)
    allsizes, allexp = [], []
    for sizes in product(*([[0, 1, 2]] * 4)):
        recs = []
        for r in range(4):
            s = np.empty(sizes[r], dtype=dt)
            s["vkey"] = np.array(range(sizes[r]), dtype="u8")[:, None]
            s["vector"] = 1
            recs.append(O.as_bytes(s))
        out = run_all_tunings(recs, desc, list(sizes))
        S = np.concatenate(recs, axis=0).view(dt).reshape(-1).copy()
        S.sort()
        assert np.array_equal(np.concatenate(out, axis=0).view(dt).reshape(-1)["vkey"], S["vkey"])
        allsizes.append(sizes)
        allexp.append(np.concatenate(out, axis=0))
    np.savez_compressed(os.path.join(HERE, "few_items.npz"), sizes=np.array(allsizes, dtype=np.int64),
                        expected=np.concatenate(allexp, axis=0), desc=np.array(desc.astuple(), dtype=np.int64),
                        elsize=np.int64(dt.itemsize))
    print("few_items: 81 size combinations x 5 tunings")


def mismatched_zeros():
    """test_mpsort.py:127-140: in sizes [0,400,0,600] -> out sizes [200,200,0,600], i4 keys"""
    rng = np.random.RandomState(1234)
    s = np.int32(rng.random_sample(size=1000) * 1000)
    dt = np.dtype([("D", "i4")])
    cuts = np.cumsum([0, 0, 400, 0, 600])
    recs = [O.as_bytes(s[cuts[i]:cuts[i + 1]].astype(dt)) for i in range(4)]
    desc = O.Desc(0, 4, 1, 1, 0)
    out = run_all_tunings(recs, desc, [200, 200, 0, 600])
    assert np.array_equal(np.concatenate(out, axis=0).view("i4").reshape(-1), np.sort(s))
    save("mismatched_zeros", recs, desc, [200, 200, 0, 600], out)
    print("mismatched_zeros ok")


def sort_struct():
    """test_mpsort.py:176-200: 10 records {value:i8,key:i8} seeded 1234, by 'key', 4 ranks"""
    np.random.seed(1234)
    s = np.empty(10, dtype=[("value", "i8"), ("key", "i8")])
    s["value"] = np.int32(np.random.random(size=10) * 1000 - 400)
    s["key"] = s["value"]
    parts = np.array_split(s, 4)
    recs = [O.as_bytes(p) for p in parts]
    desc = O.Desc(8, 8, 1, 1, 0)
    sizes = [len(p) for p in parts]
    out = run_all_tunings(recs, desc, sizes)
    b = s.copy()
    b.sort(order="key")
    assert np.array_equal(np.concatenate(out, axis=0).view(s.dtype).reshape(-1)["value"], b["value"])
    save("sort_struct", recs, desc, sizes, out)
    print("sort_struct ok")


def ties():
    """NOT pinned by any reference test: equal keys with distinguishable payloads.
    5 ranks, uneven in/out sizes, 7 distinct keys: the reference's answer is recorded."""
    rng = np.random.RandomState(7)
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    sizes = [700, 0, 1300, 5, 995]
    outsizes = [600, 600, 600, 600, 600]
    recs = []
    for r, n in enumerate(sizes):
        a = np.zeros(n, dtype=dt)
        a["key"] = rng.randint(0, 7, size=n)
        a["tag"] = (r << 40) + np.arange(n)
        recs.append(O.as_bytes(a))
    desc = O.Desc(0, 8, 1, 0, 0)
    out = run_all_tunings(recs, desc, outsizes)
    save("ties", recs, desc, outsizes, out)
    print("ties ok")


def synthetic(name, kind, elsize, p, n, is_signed):
    """the bench generators at a size the reference sorts in a blink"""
    recs = [O.generate(n, elsize, kind, 0x5EED0001, r, p) for r in range(p)]
    desc = O.Desc(0, 8, 1, is_signed, 0)
    out = run_all_tunings(recs, desc, [n] * p, tunings=[0, O.DISABLE_SPARSE_ALLTOALLV])
    save(name, recs, desc, [n] * p, out, kind=np.int64(kind), seed=np.int64(0x5EED0001))
    print(name, "ok")


if __name__ == "__main__":
    O.build()
    assert O.have_ref(), "oracle/_ref missing"
    issue7()
    few_items()
    mismatched_zeros()
    sort_struct()
    ties()
    synthetic("synth_uniform16", 0, 16, 4, 1500, 0)
    synthetic("synth_mostly_sorted16", 1, 16, 8, 700, 0)
    synthetic("synth_particles48", 2, 48, 8, 600, 1)
