"""Kernel tuning aid: the SecondSort merge ALONE, at full size, on ONE GPU.
p sorted runs of n/p synthetic records each (what a rank of a p-GPU sort holds after the
exchange: every run covers the rank's whole key range) -> mpsort_util_merge_runs, timed per
kernel class, verified (order, tie order = run order, multiset of records).
  python tools/merge_probe.py [p] [log2n] [elsize] [kind] [reps]
ncu sees the same kernels as an 8-GPU run does, without the 8 GPUs."""
import ctypes
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_capi", os.path.join(ROOT, "mp-sort_b200", "mpsort", "_capi.py"))
C = importlib.util.module_from_spec(spec)
spec.loader.exec_module(C)
lib = C.lib

p = int(sys.argv[1]) if len(sys.argv) > 1 else 8
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
E = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kind = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
n = 1 << log2n
per = n // p
comm = lib.mpsort_comm_self(0)
desc = C.RadixDesc(0, 8, 1, 1 if kind == 2 else 0, 0)
runs = lib.mpsort_util_dev_malloc(0, n * E)
out = lib.mpsort_util_dev_malloc(0, n * E)
rd = (ctypes.c_int64 * (p + 1))()
# PROBE_SHAPE=two: what a rank of a sort of MOSTLY SORTED input receives -- 99% from itself, 1% from one
# neighbour, nothing from the rest (the tile kernel then runs one merge round, not log2 p)
shape = os.environ.get("PROBE_SHAPE", "even")
sizes = [per] * (p - 1) + [n - per * (p - 1)]
if shape == "two" and p >= 2:
    sizes = [0] * p
    sizes[p // 2] = n - n // 100
    sizes[p // 2 - 1] = n // 100
for r in range(p):
    rd[r + 1] = rd[r] + sizes[r]
    ptr = ctypes.c_void_p(runs + rd[r] * E)
    cnt = rd[r + 1] - rd[r]
    if cnt == 0:
        continue
    # tags (rank << 40) + i: ties across runs must come out in run order
    lib.mpsort_util_generate_as(comm, ptr, cnt, E, kind, 0x5EED0001, r, p)
    lib.mpsort_mpi_desc_impl(ptr, cnt, E, ctypes.byref(desc), comm, 0, b"merge_probe")
h_in = C.multiset_hash(comm, runs, n, E)
assert lib.mpsort_util_merge_runs(comm, p, runs, rd, out, E, ctypes.byref(desc)) == 0, "radix fallback, not the merge"
lib.mpsort_util_kernel_timing(comm, 1)
e0 = lib.mpsort_util_event_create(comm)
e1 = lib.mpsort_util_event_create(comm)
lib.mpsort_util_event_record(comm, e0)
for _ in range(reps):
    lib.mpsort_util_merge_runs(comm, p, runs, rd, out, E, ctypes.byref(desc))
lib.mpsort_util_event_record(comm, e1)
ms = lib.mpsort_util_event_elapsed_ms(comm, e0, e1) / reps
kt = C.kernel_times(comm)
bad = lib.mpsort_util_check_sorted(comm, out, n, E, ctypes.byref(desc), 1, 8, None)
h_out = C.multiset_hash(comm, out, n, E)
st = C.last_stats(comm, 1)
print("%-36s merge[%s] p=%d n=2^%d E=%d kind=%d: %.3f ms per merge (wall, incl. sample sort + host syncs) = %.0f GB/s (2E per record)  "
      "bad=%d multiset_ok=%s | %s" % (
          os.path.basename(os.environ.get("MPSORT_LIB", "default")) + " " + " ".join("%s=%s" % (k[7:], v) for k, v in sorted(os.environ.items()) if k.startswith("MPSORT_") and k != "MPSORT_LIB"),
          shape, p, log2n, E, kind, ms, 2.0 * E * n / ms / 1e6, bad, h_in == h_out,
          "  ".join("%s %.3f/%d" % (k, v[0] / reps, v[1] // reps) for k, v in kt.items() if v[1])))
if bad or h_in != h_out:
    sys.exit(1)
