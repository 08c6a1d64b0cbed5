"""World-size-2 (and 3) multi-process test on CPU (gloo): the host-side logic of the
N > 1 path. Each process plays one rank: it sorts its share with numpy (standing in
for the device sort, which needs a GPU), computes its local splitter counts, the
ranks all-gather them over torch.distributed/gloo, and every rank runs the PRODUCT's
layout solver (mpsort_solve_layout / mpsort_cumulative_counts in libmpsort-b200.so,
pure host C) to get the exchange matrix. The resulting redistribution must equal the
oracle's output and the reference algorithm's SendCount matrix, incl. ties that
straddle a splitter (filled from lower source ranks first, mpsort-mpi.c:704-724)."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, distinct, q):
    sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import importlib.util
    spec = importlib.util.spec_from_file_location("_capi", os.path.join(ROOT, "mp-sort_b200", "mpsort", "_capi.py"))
    C = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(C)
    import mpsort_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = world
        rng = np.random.default_rng(1234)           # same stream on every rank: everyone knows all inputs
        sizes = [int(rng.integers(50, 400)) for _ in range(p)]
        outsizes = sizes[::-1]
        dt = np.dtype([("key", "u8"), ("tag", "u8")])
        arrays = []
        for r in range(p):
            a = np.zeros(sizes[r], dtype=dt)
            a["key"] = rng.integers(0, distinct, size=sizes[r]) if distinct else rng.integers(0, 1 << 62, size=sizes[r])
            a["tag"] = (r << 40) + np.arange(sizes[r])
            arrays.append(a)
        mine = arrays[rank][np.argsort(arrays[rank]["key"], kind="stable")]    # "FirstSort"
        # cumulative output counts by the product's host code
        Cc = (ctypes.c_int64 * (p + 1))()
        C.lib.mpsort_cumulative_counts(p, (ctypes.c_int64 * p)(*outsizes), Cc)
        # splitters = keys at global ranks C[b]-1 (what the device descent finds); agree on them via gloo
        gathered = [None] * p
        dist.all_gather_object(gathered, mine["key"].tolist())
        allk = np.sort(np.concatenate([np.array(g, dtype=np.uint64) for g in gathered]))
        split = [allk[Cc[b] - 1] if Cc[b] > 0 else allk[0] for b in range(1, p)]
        row = [int(np.searchsorted(mine["key"], s, "left")) for s in split] + \
              [int(np.searchsorted(mine["key"], s, "right")) for s in split]
        rows = [None] * p
        dist.all_gather_object(rows, row)                                    # "LayDistr"
        clt = (ctypes.c_int64 * (p * (p - 1)))(*[rows[j][i] for j in range(p) for i in range(p - 1)])
        cle = (ctypes.c_int64 * (p * (p - 1)))(*[rows[j][p - 1 + i] for j in range(p) for i in range(p - 1)])
        cut = (ctypes.c_int64 * (p * (p + 1)))()
        rc = C.lib.mpsort_solve_layout(p, Cc, clt, cle, (ctypes.c_int64 * p)(*sizes), cut)    # "LaySolve"
        assert rc == 0
        sendcounts = [[cut[j * (p + 1) + k + 1] - cut[j * (p + 1) + k] for k in range(p)] for j in range(p)]
        # "Exchange": my slices to everyone
        pieces = [mine[cut[rank * (p + 1) + k]:cut[rank * (p + 1) + k + 1]].tobytes() for k in range(p)]
        allpieces = [None] * p
        dist.all_gather_object(allpieces, pieces)
        recv = np.concatenate([np.frombuffer(allpieces[j][rank], dtype=dt) for j in range(p)])
        out = recv[np.argsort(recv["key"], kind="stable")]                  # "SecondSort"
        desc = O.Desc(0, 8, 1, 0, 0)
        exp, info = O.c_sort([O.as_bytes(a) for a in arrays], desc, outsizes, O.DISABLE_GATHER_SORT)
        ok = np.array_equal(O.as_bytes(out), exp[rank]) and sendcounts == info["sendcounts"].tolist()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,distinct", [(2, None), (2, 3), (3, 5)])
def test_layout_and_exchange_over_gloo(world, distinct):
    # stdlib multiprocessing: the parent must not import torch (it may already hold the
    # system NCCL through libmpsort-b200.so; torch wants its own newer one)
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + (distinct or 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, distinct, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert [r[1] for r in results] == [True] * world
    assert all(p.exitcode == 0 for p in procs)
