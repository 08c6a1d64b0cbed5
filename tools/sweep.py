"""Kernel tuning aid: time the kernel classes of one single-GPU sort for the library
given in $MPSORT_LIB (ctypes only). python tools/sweep.py [log2n] [elsize] [kind]"""
import ctypes
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_capi", os.path.join(ROOT, "mp-sort_b200", "mpsort", "_capi.py"))
C = importlib.util.module_from_spec(spec)
spec.loader.exec_module(C)
lib = C.lib

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << log2n
comm = lib.mpsort_comm_self(0)
desc = C.RadixDesc(0, 8, 1, 1 if kind == 2 else 0, 0)
din = lib.mpsort_util_dev_malloc(0, n * E)
dout = lib.mpsort_util_dev_malloc(0, n * E)
# SWEEP_AS="r,p": the records rank r of p would hold (mostly-sorted keys then span the range of a p-GPU run)
if os.environ.get("SWEEP_AS"):
    r_, p_ = [int(x) for x in os.environ["SWEEP_AS"].split(",")]
    lib.mpsort_util_generate_as(comm, din, n, E, kind, 0x5EED0001, r_, p_)
else:
    lib.mpsort_util_generate(comm, din, n, E, kind, 0x5EED0001)
for _ in range(2):
    lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, ctypes.byref(desc), comm, 0, b"sweep")
lib.mpsort_util_kernel_timing(comm, 1)
e0 = lib.mpsort_util_event_create(comm)
e1 = lib.mpsort_util_event_create(comm)
K = 4
lib.mpsort_util_event_record(comm, e0)
for _ in range(K):
    lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, ctypes.byref(desc), comm, 0, b"sweep")
lib.mpsort_util_event_record(comm, e1)
ms = lib.mpsort_util_event_elapsed_ms(comm, e0, e1) / K
kt = C.kernel_times(comm)
bad = lib.mpsort_util_check_sorted(comm, dout, n, E, ctypes.byref(desc), 1 if E >= 16 else 0, 8, None)
tag = " ".join("%s=%s" % (k[7:], v) for k, v in sorted(os.environ.items()) if k.startswith("MPSORT_") and k != "MPSORT_LIB")
st = C.last_stats(comm, 1)
print("%-28s n=2^%d E=%d kind=%d %s passes=%d hybrid=%d step %.3f ms  %.2f Grec/s  bad=%d  | " % (
    os.path.basename(os.environ.get("MPSORT_LIB", "default")), log2n, E, kind, tag, st["first_sort_passes"], st["hybrid"], ms, n / ms / 1e6, bad) +
    "  ".join("%s %.3f/%d" % (k, v[0] / K, v[1] // K) for k, v in kt.items() if v[1]))
