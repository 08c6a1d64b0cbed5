"""Worker of tests/test_hostflow_mock.py and tests/test_gpu_parity.py: host buffers that move in chunks
(MPSORT_CHUNK_MIN_BYTES / MPSORT_CHUNK_BYTES lowered by the caller) -- input chunks behind which the histogram /
key extraction pass runs, output ranges that leave while the fix-up or the merge of the next exchange part still
runs -- one rank (record mode + hybrid, bare keys, index mode; out of place and in place) and four
rank threads (exchange in two parts). Every output byte for byte against the oracle's contract."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402
import mpsort_oracle as O  # noqa: E402
rng = np.random.default_rng(1)
ok = True
# single rank, record mode hybrid (n >= 2^22), host buffers out of place and in place; index mode 48-byte
for E, n, kind, signed in ((16, (1 << 22) + 777, 0, 0), (16, (1 << 22) + 5, 1, 0), (48, 300001, 2, 1), (8, (1 << 22) + 3, 0, 0)):
    if E == 8:
        a = rng.integers(0, 1 << 63, size=n, dtype=np.uint64).view(np.uint8).reshape(n, 8)
    else:
        a = O.generate(n, E, kind, 7, 0, 1)
    desc = O.Desc(0, 8, 1, signed, 0)
    exp = O.numpy_sort([a], desc)[0]
    comm = mpsort.Comm.self(0)
    out = np.zeros_like(a)
    C.lib.mpsort_mpi_newarray_desc_impl(a.ctypes.data, n, out.ctypes.data, n, E, C.byref(C.RadixDesc(0, 8, 1, signed, 0)), comm.handle, 0, b"chunk")
    st = C.last_stats(comm.handle, 1)
    good = np.array_equal(out, exp)
    b = a.copy()
    C.lib.mpsort_mpi_desc_impl(b.ctypes.data, n, E, C.byref(C.RadixDesc(0, 8, 1, signed, 0)), comm.handle, 0, b"chunk")
    good2 = np.array_equal(b, exp)
    print(E, n, kind, "->", good, good2, "hybrid", st["hybrid"], "passes", st["first_sort_passes"])
    ok &= good and good2
    comm.destroy()
# 4 ranks (threads), host buffers, exchange in 2 parts
p = 4
for E, kind, signed in ((16, 0, 0), (48, 2, 1)):
    sizes = [200000 + 17 * r for r in range(p)]; outs_n = sizes[::-1]
    recs = [O.generate(sizes[r], E, kind, 9, r, p) for r in range(p)]
    desc = O.Desc(0, 8, 1, signed, 0)
    exp = O.numpy_sort(recs, desc, outs_n)
    outs = [np.zeros((outs_n[r], E), np.uint8) for r in range(p)]
    d = C.RadixDesc(0, 8, 1, signed, 0)
    def work(comm):
        r = comm.rank
        C.lib.mpsort_mpi_newarray_desc_impl(recs[r].ctypes.data, len(recs[r]), outs[r].ctypes.data, len(outs[r]), E, C.byref(d), comm.handle, 0, b"chunk")
        return C.last_stats(comm.handle, p)
    stats = mpsort.run_local(p, work)
    good = all(np.array_equal(outs[r], exp[r]) for r in range(p))
    print("p=4", E, kind, "->", good, "parts", [s["exchange_phases"] for s in stats], "merge", [s["second_sort_merge_tiles"] for s in stats])
    ok &= good
print("CHUNK OK" if ok else "CHUNK FAILED")
sys.exit(0 if ok else 1)
