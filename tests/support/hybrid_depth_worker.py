"""Worker of tests/test_hostflow_mock.py: keys whose four top varying digits leave runs of 16 equal high
parts and whose fifth digit separates them (seven varying bytes), plus one run of 300 equal keys that goes
to the long-run work list. Prints the number of passes the record-mode hybrid took; exit 1 on a wrong order.
The same bytes must come out at any depth (every pass is stable)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402

n = 1 << 22
rng = np.random.default_rng(5)
i = rng.permutation(n).astype(np.uint64)
g = ((i // 16) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
rec = np.zeros(n, dtype=[("key", "u8"), ("tag", "u8")])
rec["key"] = (g << np.uint64(24)) | rng.integers(0, 1 << 24, n, dtype=np.uint64)
rec["tag"] = np.arange(n)
rec["key"][:600:2] = rec["key"][3]
exp = rec[np.argsort(rec["key"], kind="stable")]
comm = mpsort.Comm.self(0)
d = C.RadixDesc(0, 8, 1, 0, 0)
C.lib.mpsort_mpi_desc_impl(rec.ctypes.data, n, 16, ctypes.byref(d), comm.handle, 0, b"hybrid_depth")
st = C.last_stats(comm.handle, 1)
ok = np.array_equal(rec, exp)
print("passes=%d hybrid=%d long_runs=%d equal=%s" % (st["first_sort_passes"], st["hybrid"], st["hybrid_long_runs"], ok))
sys.exit(0 if ok else 1)
