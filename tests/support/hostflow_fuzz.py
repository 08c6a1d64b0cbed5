"""Randomised cases for the C host code on the mock device (worker of tests/test_hostflow_mock.py; also a tool:
`MPSORT_ALLOW_MOCK_DEVICE=1 MPSORT_LIB=tests/native/_build/libmpsort-hostmock.so python tests/support/hostflow_fuzz.py SEED NCASES [nccl]`).
Each case: random number of ranks, input and output sizes (zeros included), record size, key shape (width, words,
signedness, offset), key distribution (uniform, few distinct values, all equal, sorted, reverse, narrow signed range),
options and exchange parts; every rank's bytes are compared with the oracle's statement of the contract
(oracle/mpsort_oracle.py: numpy_sort). Exit 1 and the failing case's parameters on a mismatch."""
import ctypes
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402
import mpsort_oracle as O  # noqa: E402

lib = C.lib
NDEV = max(1, lib.mpsort_util_device_count())
REAL_GPU = not hasattr(lib, "mpsk_launch_count") or not hasattr(lib, "mocksync_cudaFree")    # real NCCL wants one device per rank
SEEN = {}
SWITCHES = ["MPSORT_NO_FUSED_PACK", "MPSORT_NO_MERGE", "MPSORT_NO_REC16", "MPSORT_NO_REBASE", "MPSORT_NO_P2P",
            "MPSORT_P2P_PULL", "MOCK_NO_IPC", "MPSORT_PEER_SPLITTER", "MPSORT_NO_PEER_SPLITTER", "MPSORT_NO_HYBRID5", "MPSORT_NO_CHAINED_PARTS", "MPSORT_NO_SELF_IN_PLACE"]
if os.environ.get("FUZZ_SWITCHES") == "shipped":       # on a real GPU: leave the unmeasured candidates out
    SWITCHES = ["MPSORT_NO_MERGE", "MPSORT_NO_REC16", "MPSORT_NO_REBASE", "MPSORT_NO_HYBRID", "MPSORT_NO_FUSED_PACK", "MPSORT_NO_PEER_SPLITTER"]


def make_case(rng):
    p = int(rng.choice([1, 2, 2, 3, 4, 5, 8, 9]))
    width = int(rng.choice([1, 2, 4, 8, 8, 8]))
    nwords = int(rng.choice([1, 1, 1, 2, 3])) if width == 8 else int(rng.choice([1, 1, 2, 3, 5]))
    signed = int(rng.integers(0, 2))
    klen = width * nwords
    offset = int(rng.choice([0, 0, 8, 3])) if klen <= 8 else int(rng.choice([0, 8]))
    E = offset + klen + int(rng.choice([0, 0, 8, 8, 5, 24]))
    if rng.integers(0, 3) == 0:
        E, offset, width, nwords, klen = 16, int(rng.choice([0, 8])), 8, 1, 8      # the record-mode shapes
    scale = int(rng.choice([0, 3, 40, 700, 6000, 60000]))          # 60000: enough per part for the merge of the runs
    big = rng.integers(0, 40) == 0                                  # rarely: enough for range compression (2^20 per rank)
    if big:
        p, scale, width, nwords, klen = int(rng.choice([1, 2])), 1300000, 8, 1, 8
        offset = int(rng.choice([0, 8])); E = int(rng.choice([16, 24]))
    sizes = [int(rng.integers(scale // 2, scale + 1)) if rng.integers(0, 4) else 0 for _ in range(p)]
    if big:
        sizes = [scale] * p
    total = sum(sizes)
    cuts = np.sort(rng.integers(0, total + 1, p - 1)) if p > 1 else np.array([], dtype=np.int64)
    if rng.integers(0, 2):
        outsizes = list(sizes)
    else:
        outsizes = [int(x) for x in np.diff(np.concatenate([[0], cuts, [total]]))]
    dist = 5 if big else int(rng.integers(0, 6))
    recs = []
    gi = 0
    for r in range(p):
        n = sizes[r]
        a = rng.integers(0, 256, size=(n, E), dtype=np.uint8)
        key = a[:, offset:offset + klen]
        if dist == 1:
            key[:] = rng.integers(0, 3, size=(n, klen), dtype=np.uint8) * (rng.integers(0, 2, size=(1, klen), dtype=np.uint8))
        elif dist == 2:
            key[:] = 7
        elif dist in (3, 4) and width == 8 and nwords == 1:
            v = np.arange(gi, gi + n, dtype=np.uint64) * np.uint64(1000003)
            if dist == 4:
                v = np.uint64(1 << 62) - v
            key[:] = v.view(np.uint8).reshape(n, 8)
        elif dist == 5 and width == 8 and nwords == 1:
            v = rng.integers(-500, 500, size=n).astype(np.int64)          # narrow signed range: range compression
            key[:] = v.view(np.uint8).reshape(n, 8)
        gi += n
        recs.append(np.ascontiguousarray(a))
    opts = int(rng.choice([0, 0, C.MPSORT_DISABLE_GATHER_SORT, C.MPSORT_DISABLE_GATHER_SORT, C.MPSORT_REQUIRE_GATHER_SORT,
                           C.MPSORT_DISABLE_SPARSE_ALLTOALLV | C.MPSORT_DISABLE_GATHER_SORT,
                           C.MPSORT_REQUIRE_SPARSE_ALLTOALLV, C.MPSORT_VERIFY_CHECKSUM | C.MPSORT_DISABLE_GATHER_SORT]))
    inplace = bool(outsizes == sizes and rng.integers(0, 2))
    return dict(p=p, E=E, offset=offset, width=width, nwords=nwords, signed=signed, sizes=sizes, outsizes=outsizes,
                dist=dist, opts=opts, inplace=inplace), recs


def run_case(par, recs, nccl, more=()):
    """`more`: further (recs, outsizes) rounds sorted on the SAME communicator after the first (buffers grow, peers'
    mappings change, sequence numbers advance)"""
    p, E = par["p"], par["E"]
    desc = O.Desc(par["offset"], par["width"], par["nwords"], par["signed"], 0)
    rounds = [(recs, par["outsizes"])] + list(more)
    exps = [O.numpy_sort(rc, desc, osz) for rc, osz in rounds]
    inss = [[r.copy() for r in rc] for rc, _ in rounds]
    outss = [ins if (par["inplace"] and [len(x) for x in ins] == list(osz)) else [np.zeros((osz[k], E), np.uint8) for k in range(p)]
             for ins, (_, osz) in zip(inss, rounds)]
    d = C.RadixDesc(par["offset"], par["width"], par["nwords"], par["signed"], 0)
    lib.mpsort_mpi_unset_options(-1)
    lib.mpsort_mpi_set_options(par["opts"])

    def sort(h, r):
        for ins, outs in zip(inss, outss):
            lib.mpsort_mpi_newarray_desc_impl(ins[r].ctypes.data, len(ins[r]), outs[r].ctypes.data, len(outs[r]), E,
                                              ctypes.byref(d), h, 0, b"fuzz")
        if r == 0:
            st = C.last_stats(h, p)
            for k in ("used_gather", "record_mode", "rebased", "p2p_exchange", "dense_exchange"):
                SEEN[k] = SEEN.get(k, 0) + (1 if st[k] else 0)
            SEEN["parts>1"] = SEEN.get("parts>1", 0) + (1 if st["exchange_phases"] > 1 else 0)
            SEEN["merge"] = SEEN.get("merge", 0) + (1 if st["second_sort_merge_tiles"] else 0)
            SEEN["resort"] = SEEN.get("resort", 0) + (1 if st["second_sort_passes"] else 0)

    if nccl and p > 1:
        uid = ctypes.create_string_buffer(C.MPSORT_UNIQUE_ID_BYTES)
        assert lib.mpsort_comm_get_unique_id(uid) == 0
        errs = []

        def body(r):
            try:
                h = ctypes.c_void_p(lib.mpsort_comm_init_rank(r, p, uid, r % NDEV))
                sort(h, r)
                lib.mpsort_comm_destroy(h)
            except BaseException as e:  # noqa: B902
                errs.append(e)
        ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(p)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(120)
        if errs or any(t.is_alive() for t in ts):
            return False
    else:
        mpsort.run_local(p, lambda comm: sort(comm.handle, comm.rank), timeout=120)
    return all(np.array_equal(outs[k], exp[k]) for outs, exp in zip(outss, exps) for k in range(p))


def main():
    seed, ncases = int(sys.argv[1]), int(sys.argv[2])
    nccl = len(sys.argv) > 3 and sys.argv[3] == "nccl"
    rng = np.random.default_rng(seed)
    for i in range(ncases):
        par, recs = make_case(rng)
        if nccl and REAL_GPU and par["p"] > NDEV:
            continue
        q = int(rng.choice([1, 1, 2, 3, 4]))
        os.environ["MPSORT_EXCHANGE_PHASES"] = str(q)
        os.environ["MPSORT_PHASES_MIN_RECORDS"] = "1"
        par["parts"] = q
        par["switches"] = [k for k in SWITCHES if rng.integers(0, 5) == 0]
        for k in SWITCHES:
            os.environ.pop(k, None)
        for k in par["switches"]:
            os.environ[k] = "1"
        os.environ["MPSORT_P2P_CE"] = str(int(rng.choice([1, 1, 0, 3])))
        # what the copies of a later exchange part wait for (exchange_p2p): the default of p, or one of the three gates
        os.environ.pop("MPSORT_CHAINED_PARTS", None)
        if i % 4:
            os.environ["MPSORT_CHAINED_PARTS"] = str(i % 4 - 1)
        par["chained"] = os.environ.get("MPSORT_CHAINED_PARTS", "default")
        more = []
        if rng.integers(0, 3) == 0:
            # the same communicator sorts two more inputs of other sizes (same record and key shape)
            for _ in range(2):
                sizes = [int(rng.integers(0, 3 * max(par["sizes"]) + 2)) for _ in range(par["p"])]
                tot = sum(sizes)
                cuts = np.sort(rng.integers(0, tot + 1, par["p"] - 1)) if par["p"] > 1 else np.array([], dtype=np.int64)
                osz = [int(x) for x in np.diff(np.concatenate([[0], cuts, [tot]]))]
                rc = [rng.integers(0, 4, size=(n, par["E"]), dtype=np.uint8) * rng.integers(0, 256, size=(1, par["E"]), dtype=np.uint8) for n in sizes]
                more.append((rc, osz))
            par["more_sizes"] = [[len(x) for x in rc] for rc, _ in more]
        if os.environ.get("FUZZ_VERBOSE"):
            print("case %d:" % i, par, flush=True)
        if not run_case(par, recs, nccl, more):
            print("FUZZ FAILED at case %d of seed %d:" % (i, seed), par)
            return 1
    print("FUZZ OK: %d cases, seed %d, %s ranks; paths taken:" % (ncases, seed, "NCCL-thread" if nccl else "in-process"), SEEN)
    return 0


if __name__ == "__main__":
    sys.exit(main())
