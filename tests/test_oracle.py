"""The oracle against the reference's golden vectors (CPU only).

tests/golden/*.npz were produced by the UNMODIFIED reference (make_golden.py). Here
the two restatements (numpy contract, C algorithm) must reproduce them bit for bit;
when oracle/_ref is present (the build container, or prebuilt on the GPU box) the
reference itself is replayed as well."""
import itertools

import numpy as np
import pytest

import mpsort_oracle as O
from conftest import GOLDEN_CASES, load_golden

TUNINGS = [0, O.DISABLE_SPARSE_ALLTOALLV, O.REQUIRE_SPARSE_ALLTOALLV, O.REQUIRE_GATHER_SORT, O.DISABLE_GATHER_SORT]


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_numpy_contract_matches_reference_golden(name):
    g = load_golden(name)
    assert same(O.numpy_sort(g["recs"], g["desc"], g["outsizes"]), g["exp"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tuning", TUNINGS)
def test_c_restatement_matches_reference_golden(name, tuning):
    g = load_golden(name)
    out, info = O.c_sort(g["recs"], g["desc"], g["outsizes"], tuning)
    assert same(out, g["exp"])
    # SendCount rows add up to what every rank holds / receives
    if info["nleaders"] == len(g["recs"]):
        assert list(info["sendcounts"].sum(axis=1)) == [len(r) for r in g["recs"]]
        assert list(info["sendcounts"].sum(axis=0)) == g["outsizes"]


def test_few_items_golden():
    """test_mpsort.py:330-351: 81 size combinations, 24-byte key, every tuning"""
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "few_items.npz"))
    desc = O.Desc(*[int(v) for v in z["desc"]])
    dt = np.dtype([("vkey", ("u8", 3)), ("vector", ("u4", 3))])
    off = 0
    for sizes in z["sizes"]:
        recs = []
        for r in range(4):
            s = np.empty(int(sizes[r]), dtype=dt)
            s["vkey"] = np.array(range(int(sizes[r])), dtype="u8")[:, None]
            s["vector"] = 1
            recs.append(O.as_bytes(s))
        n = int(sum(sizes))
        exp = z["expected"][off:off + n]
        off += n
        for tuning in TUNINGS:
            out, _ = O.c_sort(recs, desc, [int(s) for s in sizes], tuning)
            assert np.array_equal(np.concatenate(out, axis=0), exp)
        assert np.array_equal(np.concatenate(O.numpy_sort(recs, desc, [int(s) for s in sizes]), axis=0), exp)


def random_case(rng, p, elsize, desc, nmax, distinct=None):
    sizes = [int(rng.integers(0, nmax + 1)) for _ in range(p)]
    if rng.random() < 0.3:
        sizes[int(rng.integers(0, p))] = 0
    total = sum(sizes)
    cuts = sorted(int(c) for c in rng.integers(0, total + 1, size=p - 1))
    outsizes = [b - a for a, b in zip([0] + cuts, cuts + [total])]
    recs = []
    for r in range(p):
        a = rng.integers(0, 256, size=(sizes[r], elsize), dtype=np.uint8)
        if distinct:
            lo, hi = desc.offset, desc.offset + desc.width * desc.nwords
            a[:, lo:hi] = 0
            a[:, lo] = rng.integers(0, distinct, size=sizes[r])
        recs.append(a)
    return recs, outsizes


CASES = [
    (O.Desc(0, 8, 1, 0, 0), 16), (O.Desc(8, 8, 1, 1, 0), 16), (O.Desc(0, 8, 1, 1, 0), 48),
    (O.Desc(4, 4, 1, 0, 0), 12), (O.Desc(0, 4, 1, 1, 0), 8), (O.Desc(0, 8, 2, 0, 0), 40),
    (O.Desc(8, 8, 3, 1, 0), 36), (O.Desc(0, 4, 3, 1, 0), 20), (O.Desc(2, 2, 1, 0, 1), 7),
    (O.Desc(1, 1, 5, 0, 1), 9),
]


@pytest.mark.parametrize("desc,elsize", CASES)
def test_c_restatement_equals_numpy_contract(desc, elsize):
    rng = np.random.default_rng(99 + elsize)
    for p, distinct in itertools.product([1, 2, 3, 5, 8], [None, 4]):
        recs, outsizes = random_case(rng, p, elsize, desc, 300, distinct)
        exp = O.numpy_sort(recs, desc, outsizes)
        for tuning in (0, O.DISABLE_GATHER_SORT, O.REQUIRE_GATHER_SORT):
            out, _ = O.c_sort(recs, desc, outsizes, tuning)
            assert same(out, exp)


def test_local_radix_sort_is_stable():
    rng = np.random.default_rng(5)
    dt = np.dtype([("key", "u8"), ("tag", "u8")])
    a = np.zeros(5000, dtype=dt)
    a["key"] = rng.integers(0, 5, size=len(a))
    a["tag"] = np.arange(len(a))
    out = O.c_radix_sort(O.as_bytes(a), O.Desc(0, 8, 1, 0, 0)).view(dt).reshape(-1)
    exp = a[np.argsort(a["key"], kind="stable")]
    assert np.array_equal(out, exp)


def test_checksum_is_signed_byte_sum():
    """mpsort-mpi.c:148-159: bytes are summed as SIGNED chars into a wrapping u64"""
    a = np.array([0x7f, 0x80, 0xff, 0x01], dtype=np.uint8)
    assert O.checksum(a) == (127 - 128 - 1 + 1) % (1 << 64)
    rng = np.random.default_rng(3)
    b = rng.integers(0, 256, size=100003, dtype=np.uint8)
    assert O.checksum(b) == int(b.view(np.int8).astype(np.int64).sum()) % (1 << 64)


def test_generator_kinds():
    g = O.generate(1000, 48, 2, 0x5EED0001, 3, 8)
    ids = g[:, :8].copy().view("<i8").reshape(-1)
    tags = g[:, 8:16].copy().view("<u8").reshape(-1)
    assert np.array_equal(tags, (3 << 40) + np.arange(1000))
    assert ids.min() >= -(1 << 20) and ids.max() < (1 << 24)
    assert 20 <= (ids == 0).sum() <= 100          # the forced 5 % run of id 0
    m = np.concatenate([O.generate(500, 16, 1, 1, r, 4) for r in range(4)])
    keys = m[:, :8].copy().view("<u8").reshape(-1)
    assert (np.diff(keys.astype(np.int64)) < 0).sum() < 40    # mostly sorted


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
class TestAgainstTheReferenceItself:
    """O1 of SURVEY.md 8(c): the unmodified reference under the single-host MPI shim"""

    @pytest.mark.parametrize("name", ["issue7", "ties", "mismatched_zeros"])
    @pytest.mark.parametrize("tuning", TUNINGS)
    def test_reference_reproduces_golden(self, name, tuning):
        g = load_golden(name)
        assert same(O.ref_sort(g["recs"], g["desc"], g["outsizes"], tuning), g["exp"])

    @pytest.mark.parametrize("desc,elsize", CASES[:8])
    def test_restatements_equal_reference_on_random_inputs(self, desc, elsize):
        rng = np.random.default_rng(7 + elsize)
        for p, distinct in [(2, None), (4, 3), (7, None), (12, 5)]:
            recs, outsizes = random_case(rng, p, elsize, desc, 200, distinct)
            ref = O.ref_sort(recs, desc, outsizes, 0)
            assert same(O.numpy_sort(recs, desc, outsizes), ref)
            out, _ = O.c_sort(recs, desc, outsizes, 0)
            assert same(out, ref)

    def test_reference_inplace(self):
        rng = np.random.default_rng(11)
        desc = O.Desc(0, 8, 1, 0, 0)
        recs = [rng.integers(0, 256, size=(n, 16), dtype=np.uint8) for n in (100, 0, 250, 31)]
        ref = O.ref_sort(recs, desc, None, 0, inplace=True)
        assert same(O.numpy_sort(recs, desc), ref)

    def test_reference_bench_drivers_run(self):
        """unmodified bench-mpi.c / main-mpi.c self-check under the shim"""
        import os
        import subprocess
        shim = os.path.join(O.REF_DIR, "mpirun-shim")
        for prog, arg in (("bench-mpi", "20000"), ("main-mpi", "20000")):
            for np_ in (1, 3, 4):
                rc = subprocess.run([shim, "-np", str(np_), os.path.join(O.REF_DIR, prog), arg],
                                    stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=120)
                assert rc.returncode == 0, rc.stderr.decode()
        r = O.run_bench16(4, 20000)
        assert r["records_per_second"] > 0 and "FirstSort" in r["phases"]
