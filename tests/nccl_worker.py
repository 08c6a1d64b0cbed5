"""Worker of test_nccl_process_per_gpu: one process per GPU (torchrun env), NCCL
transport. Every rank generates all ranks' inputs with the oracle generator, sorts its
own share on its GPU and compares its output chunk with the oracle's."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402
import mpsort_oracle as O  # noqa: E402


def main():
    comm = mpsort.Comm.from_env()
    p, r = comm.size, comm.rank
    ok = True
    for kind, E, signed, n, tuning in ((0, 16, 0, 20000, []), (2, 48, 1, 30000, []), (1, 16, 0, 25000, []),
                                       (2, 48, 1, 30000, ["DISABLE_SPARSE_ALLTOALLV"]),
                                       (0, 16, 0, 50, ["REQUIRE_GATHER_SORT"]), (0, 16, 0, 50, [])):
        sizes = [n + 13 * k for k in range(p)]
        outsizes = sizes[::-1]
        recs = [O.generate(sizes[k], E, kind, 0x5EED0001, k, p) for k in range(p)]
        desc = O.Desc(0, 8, 1, signed, 0)
        exp = O.numpy_sort(recs, desc, outsizes)
        dt = np.dtype([("key", "i8" if signed else "u8"), ("rest", "u1", E - 8)])
        mine = recs[r].copy().view(dt).reshape(-1)
        out = np.zeros(outsizes[r], dtype=dt)
        mpsort.sort(mine, "key", out=out, comm=comm, tuning=tuning)
        good = np.array_equal(O.as_bytes(out), exp[r])
        # device-resident, in place
        dev = mpsort.DeviceArray.from_host(recs[r].copy().view(dt).reshape(-1), comm.device)
        if sizes == outsizes:
            mpsort.sort(dev, "key", comm=comm, tuning=tuning)
            good &= np.array_equal(O.as_bytes(dev.to_host()), exp[r])
        dev.free()
        allgood = comm.allgather(bool(good))
        if r == 0:
            print("kind", kind, "E", E, "tuning", tuning, "->", allgood, C.last_stats(comm.handle, p)["sendcounts"])
        ok &= all(allgood)
    comm.barrier()
    if r == 0:
        print("NCCL PARITY OK" if ok else "NCCL PARITY FAILED")
    comm.destroy()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
