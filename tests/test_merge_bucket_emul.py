"""CPU check of a CANDIDATE kernel's arithmetic (off by default, MPSORT_MERGE_BUCKET=1):
the bucket path of merge_tile_bucket_kernel, emulated on the host from the SAME header the
kernel is compiled from (mp-sort_b200/csrc/mpsort_merge_bucket.cuh). Whenever the spread test
passes, the produced order must be the stable merge of the tile's runs; when keys are not spread
(duplicates, a key outside the boundary keys) the tile must be handed to the merge-path rounds."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "native", "merge_bucket_emul.cpp")
OUT = os.path.join(ROOT, "tests", "native", "_build", "libmbk_emul.so")
TILE = 4096


@pytest.fixture(scope="module")
def emul():
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr = os.path.join(ROOT, "mp-sort_b200", "csrc", "mpsort_merge_bucket.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run([cxx, "-O1", "-std=c++17", "-x", "c++", "-shared", "-fPIC",
                        "-I", os.path.join(ROOT, "mp-sort_b200", "csrc"), "-o", OUT, SRC], check=True)
    dll = ctypes.CDLL(OUT)
    dll.mbk_emul_tile.restype = ctypes.c_int
    dll.mbk_emul_tile.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                  ctypes.c_void_p, ctypes.c_void_p]
    dll.mbk_emul_shift.restype = ctypes.c_uint32
    dll.mbk_emul_shift.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    dll.mbk_emul_bucket.restype = ctypes.c_uint32
    dll.mbk_emul_bucket.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    return dll


def make_tile(rng, p, cnt, draw):
    """p sorted runs laid out run after run (like the receive buffer): keys in run-major order and
    their source positions (increasing with (run, index)), as the kernel's load phase produces them"""
    lens = rng.multinomial(cnt, np.ones(p) / p)
    keys, src, base = [], [], 0
    for r in range(p):
        k = np.sort(draw(lens[r]).astype(np.uint64))
        keys.append(k)
        gap = int(rng.integers(0, 1000))              # the slice sits somewhere inside its run
        src.append(np.arange(base + gap, base + gap + lens[r], dtype=np.uint32))
        base += gap + int(lens[r]) + int(rng.integers(0, 1000))
    return np.concatenate(keys), np.concatenate(src)


def run_tile(dll, keys, src, klo, khi, rng=None):
    cnt = len(keys)
    order = None
    if rng is not None:
        order = rng.permutation(cnt).astype(np.uint32)
    out = np.full(cnt, 0xFFFFFFFF, np.uint32)
    rc = dll.mbk_emul_tile(keys.ctypes.data, src.ctypes.data, cnt, int(klo), int(khi),
                           order.ctypes.data if order is not None else None, out.ctypes.data)
    return rc, out


def expected(keys, src):
    return src[np.lexsort((src, keys))]               # by key, ties by source position


def test_bucket_map_is_monotone_and_in_range(emul):
    rng = np.random.default_rng(5)
    for _ in range(200):
        klo = int(rng.integers(0, 2**63, dtype=np.uint64)) * int(rng.integers(0, 2))
        span = int(rng.integers(0, 2**62, dtype=np.uint64)) >> int(rng.integers(0, 62))
        khi = min(klo + span, 2**64 - 1)
        sh = emul.mbk_emul_shift(klo, khi)
        assert (khi - klo) >> sh < 8192
        if khi - klo >= 4096:
            assert (khi - klo) >> sh >= 4096          # at least half of the buckets are in use
        ks = np.sort(rng.integers(klo, khi, size=50, dtype=np.uint64, endpoint=True))
        bs = [emul.mbk_emul_bucket(int(k), klo, sh) for k in ks]
        assert bs == sorted(bs) and bs[-1] < 8192
    assert emul.mbk_emul_shift(0, 2**64 - 1) == 51 and emul.mbk_emul_bucket(2**64 - 1, 0, 51) == 8191


@pytest.mark.parametrize("p", [2, 3, 8, 32])
def test_spread_tiles_take_the_bucket_path_and_are_exact(emul, p):
    rng = np.random.default_rng(100 + p)
    taken = 0
    for trial in range(40):
        cnt = int(rng.integers(1, TILE + 1)) if trial else TILE
        lo = int(rng.integers(0, 2**63, dtype=np.uint64))
        span = int(2 ** rng.uniform(14, 62))
        keys, src = make_tile(rng, p, cnt, lambda n: rng.integers(lo, lo + span, size=n, dtype=np.uint64))
        # boundary keys as the kernel sees them: at or beyond the extreme keys
        klo, khi = int(keys.min()) - int(rng.integers(0, 3)), int(keys.max()) + int(rng.integers(0, 3))
        rc, out = run_tile(emul, keys, src, max(klo, 0), khi, rng)
        assert rc in (0, 1)
        if rc == 0:
            taken += 1
            assert np.array_equal(out, expected(keys, src))
    assert taken >= 36, "uniform keys in a tight window must nearly always pass the spread test"


def test_ties_are_ordered_by_source_position(emul):
    """pairs and short groups of equal keys spread over the runs: still the bucket path, and equal
    keys come out in (run, index) order = the stable merge of stdlib/msort.c:78"""
    rng = np.random.default_rng(7)
    for trial in range(30):
        distinct = rng.integers(0, 2**40, size=1500, dtype=np.uint64)
        keys, src = make_tile(rng, 8, 3000, lambda n: rng.choice(distinct, size=n))
        rc, out = run_tile(emul, keys, src, int(keys.min()), int(keys.max()), rng)
        assert rc == 0
        assert np.array_equal(out, expected(keys, src))


def test_duplicates_and_out_of_range_keys_fall_back(emul):
    rng = np.random.default_rng(9)
    # few distinct values: some bucket is longer than CMAX
    keys, src = make_tile(rng, 8, TILE, lambda n: rng.integers(0, 32, size=n, dtype=np.uint64) << np.uint64(30))
    assert run_tile(emul, keys, src, int(keys.min()), int(keys.max()), rng)[0] == 1
    # all keys equal
    keys, src = make_tile(rng, 4, 1000, lambda n: np.full(n, 12345, np.uint64))
    assert run_tile(emul, keys, src, 12345, 12345, rng)[0] == 1
    # open first / last tile of a part: the range is the whole key space, everything in one bucket
    keys, src = make_tile(rng, 8, TILE, lambda n: rng.integers(10**6, 10**6 + 10**5, size=n, dtype=np.uint64))
    assert run_tile(emul, keys, src, 0, 2**64 - 1, rng)[0] == 1
    # a key outside the boundary keys must never reach the bucket arithmetic
    keys, src = make_tile(rng, 8, 2000, lambda n: rng.integers(1000, 10**9, size=n, dtype=np.uint64))
    assert run_tile(emul, keys, src, int(keys.min()) + 1, int(keys.max()), rng)[0] == 1
    assert run_tile(emul, keys, src, int(keys.min()), int(keys.max()) - 1, rng)[0] == 1
    # skew inside the window (half of the keys in 1/1000 of it): exact whichever path is chosen
    for trial in range(20):
        def draw(n):
            a = rng.integers(0, 2**40, size=n, dtype=np.uint64)
            b = rng.integers(0, 2**30, size=n, dtype=np.uint64)
            return np.where(rng.random(n) < 0.5, a, b)
        keys, src = make_tile(rng, 8, TILE, draw)
        rc, out = run_tile(emul, keys, src, int(keys.min()), int(keys.max()), rng)
        assert rc in (0, 1)
        if rc == 0:
            assert np.array_equal(out, expected(keys, src))


def test_small_and_empty_tiles(emul):
    rng = np.random.default_rng(11)
    for cnt in (0, 1, 2, 7):
        keys, src = make_tile(rng, 4, cnt, lambda n: rng.integers(0, 2**50, size=n, dtype=np.uint64))
        klo = int(keys.min()) if cnt else 0
        khi = int(keys.max()) if cnt else 0
        rc, out = run_tile(emul, keys, src, klo, khi)
        if rc == 0:
            assert np.array_equal(out, expected(keys, src))
