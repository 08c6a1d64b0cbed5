"""Worker of tests/test_hostflow_mock.py: the NCCL transport of the host code (mpsort_comm_init_rank:
one rank per THREAD of this process, which real NCCL allows too) against the mock device, whose NCCL and
CUDA IPC work between threads. Runs the cases of tests/nccl_worker.py plus 2^22 records per rank (the
size at which the exchange is cut into parts) and compares every rank's bytes with the oracle.
usage: nccl_threads_worker.py P [big]      (env: MPSORT_LIB = the mock build + MPSORT_ALLOW_MOCK_DEVICE=1; MOCK_NO_IPC, MPSORT_* switches;
BIG_LOG2N: records per rank of the big cases, default 22 = the default threshold for cutting the exchange into parts)"""
import ctypes
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpsort  # noqa: E402,F401
from mpsort import _capi as C  # noqa: E402
import mpsort_oracle as O  # noqa: E402

lib = C.lib


def run_ranks(p, fn):
    uid = ctypes.create_string_buffer(C.MPSORT_UNIQUE_ID_BYTES)
    assert lib.mpsort_comm_get_unique_id(uid) == 0
    res, err = [None] * p, [None] * p

    def body(r):
        try:
            h = ctypes.c_void_p(lib.mpsort_comm_init_rank(r, p, uid, 0))
            res[r] = fn(h, r)
            lib.mpsort_comm_destroy(h)
        except BaseException as e:  # noqa: B902
            err[r] = e
    ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(p)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(600)
    assert not any(t.is_alive() for t in ts), "rank threads hang (collectives out of order?)"
    for e in err:
        if e is not None:
            raise e
    return res


def main():
    p = int(sys.argv[1])
    big = len(sys.argv) > 2 and sys.argv[2] == "big"
    cases = [(0, 16, 0, 20000, 0), (2, 48, 1, 30000, 0), (1, 16, 0, 25000, 0),
             (2, 48, 1, 30000, C.MPSORT_DISABLE_SPARSE_ALLTOALLV), (3, 24, 0, 9000, 0),
             (0, 16, 0, 50, C.MPSORT_REQUIRE_GATHER_SORT), (0, 16, 0, 50, 0), (0, 16, 0, 0, 0)]
    if big:
        nbig = 1 << int(os.environ.get("BIG_LOG2N", "22"))
        cases = [(0, 16, 0, nbig, 0), (2, 48, 1, nbig, 0)]
    ok = True
    for kind, E, signed, n, opts in cases:
        sizes = [n + 13 * k if n else 0 for k in range(p)]
        if kind == 1 and n:
            sizes = [n] * p                          # the mostly-sorted generator wants equal sizes
        outsizes = sizes[::-1]
        recs = [O.generate(sizes[k], E, kind, 0x5EED0001, k, p) for k in range(p)]
        desc = O.Desc(0, 8, 1, signed, 0)
        exp = O.numpy_sort(recs, desc, outsizes)
        outs = [np.zeros((outsizes[k], E), np.uint8) for k in range(p)]
        d = C.RadixDesc(0, 8, 1, signed, 0)
        lib.mpsort_mpi_unset_options(-1)
        lib.mpsort_mpi_set_options(opts | (C.MPSORT_DISABLE_GATHER_SORT if n > 1000 else 0))

        def work(h, r):
            # (small cases) twice on the same communicator: arena reuse, re-mapping decisions, sequence numbers
            for _ in range(1 if big else 2):
                outs[r][:] = 0
                lib.mpsort_mpi_newarray_desc_impl(recs[r].ctypes.data, len(recs[r]), outs[r].ctypes.data, len(outs[r]), E,
                                                  ctypes.byref(d), h, 0, b"nccl_threads")
            return C.last_stats(h, p)
        stats = run_ranks(p, work)
        good = all(np.array_equal(outs[k], exp[k]) for k in range(p))
        print("kind", kind, "E", E, "n", n, "opts", opts, "->", good, "p2p", [s["p2p_exchange"] for s in stats],
              "phases", [s["exchange_phases"] for s in stats], "gather", stats[0]["used_gather"], "dense", stats[0]["dense_exchange"])
        ok &= good
        if os.environ.get("EXPECT_P2P") is not None and n > 1000:
            ok &= all(s["p2p_exchange"] == int(os.environ["EXPECT_P2P"]) for s in stats)
        if os.environ.get("EXPECT_PHASES") is not None and big:
            ok &= all(s["exchange_phases"] == int(os.environ["EXPECT_PHASES"]) for s in stats)
    print("NCCL THREADS OK" if ok else "NCCL THREADS FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
