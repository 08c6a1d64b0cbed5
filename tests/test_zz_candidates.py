"""Opt-in GPU checks of CANDIDATE kernels that are off by default and were written without a GPU
at hand (end of round 1). They run only with MPSORT_TEST_CANDIDATES=1, so that the default suite
states what the shipped paths do; tools/candidates_ab.sh runs them and times each candidate.

  MPSORT_MERGE_BUCKET=1   one-round bucket merge of the received runs (merge_tile_bucket_kernel)
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MPSORT_TEST_CANDIDATES") != "1",
                                 reason="candidate kernels: set MPSORT_TEST_CANDIDATES=1")]

WORKER = r"""
import sys, os
sys.path.insert(0, os.path.join(%(root)r, "mp-sort_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, ctypes
import mpsort
from mpsort import _capi as C
import mpsort_oracle as O
lib = C.lib
ok = True
for p, n, E, kind, signed in ((4, 60000, 16, 0, 0), (8, 40000, 16, 1, 0), (3, 50000, 48, 2, 1), (2, 70000, 24, 3, 0), (8, 30000, 16, 0, 0)):
    sizes = [n + 17 * r for r in range(p)]
    outsizes = sizes[::-1]
    recs = [O.generate(sizes[r], E, kind, 0x5EED0001, r, p) for r in range(p)]
    desc = O.Desc(0, 8, 1, signed, 0)
    exp = O.numpy_sort(recs, desc, outsizes)
    outs = [np.zeros((outsizes[r], E), np.uint8) for r in range(p)]
    d = C.RadixDesc(0, 8, 1, signed, 0)
    lib.mpsort_mpi_unset_options(-1)
    lib.mpsort_mpi_set_options(C.MPSORT_DISABLE_GATHER_SORT)
    def work(comm):
        r = comm.rank
        lib.mpsort_mpi_newarray_desc_impl(recs[r].ctypes.data, len(recs[r]), outs[r].ctypes.data, len(outs[r]), E,
                                          ctypes.byref(d), comm.handle, 0, b"cand")
        return C.last_stats(comm.handle, p)
    stats = mpsort.run_local(p, work)
    good = all(np.array_equal(outs[r], exp[r]) for r in range(p))
    print("p", p, "E", E, "kind", kind, "->", good, "merge tiles", [s["second_sort_merge_tiles"] for s in stats],
          "to the rounds", [s["merge_bucket_fallback_tiles"] for s in stats])
    ok &= good and all(s["second_sort_merge_tiles"] > 0 for s in stats)
print("CANDIDATE OK" if ok else "CANDIDATE FAILED")
sys.exit(0 if ok else 1)
"""


def test_bucket_merge_matches_oracle():
    env = dict(os.environ, MPSORT_MERGE_BUCKET="1")
    rc = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, timeout=900,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"CANDIDATE OK" in rc.stdout, rc.stdout.decode()[-4000:]
