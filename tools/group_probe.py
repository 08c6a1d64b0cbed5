"""Kernel tuning aid: one multi-rank sort on ONE GPU (rank threads), so that ncu can
see the exchange-side kernels (merge, splitters). python tools/group_probe.py [p] [log2n] [elsize] [kind]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import importlib.util
import threading
spec = importlib.util.spec_from_file_location("_capi", os.path.join(ROOT, "mp-sort_b200", "mpsort", "_capi.py"))
C = importlib.util.module_from_spec(spec)
spec.loader.exec_module(C)

lib = C.lib
p = int(sys.argv[1]) if len(sys.argv) > 1 else 2
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 26
E = int(sys.argv[3]) if len(sys.argv) > 3 else 16
kind = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = 1 << log2n
desc = C.RadixDesc(0, 8, 1, 1 if kind == 2 else 0, 0)


class _Comm(object):
    def __init__(self, h):
        self.handle = h


def work(comm):
    din = lib.mpsort_util_dev_malloc(0, n * E)
    dout = lib.mpsort_util_dev_malloc(0, n * E)
    lib.mpsort_util_generate(comm.handle, din, n, E, kind, 0x5EED0001)
    for it in range(3):
        if it == 1:
            lib.mpsort_util_kernel_timing(comm.handle, 1)
        lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, ctypes.byref(desc), comm.handle, 0, b"probe")
    kt = C.kernel_times(comm.handle)
    bad = lib.mpsort_util_check_sorted(comm.handle, dout, n, E, ctypes.byref(desc), 1, 8, None)
    lib.mpsort_util_dev_free(0, din)
    lib.mpsort_util_dev_free(0, dout)
    return kt, bad, C.last_run()


devs = (ctypes.c_int * p)(*([0] * p))
comms = (ctypes.c_void_p * p)()
assert lib.mpsort_comm_init_local_group(p, devs, comms) == 0
res = [None] * p
def body(r):
    res[r] = work(_Comm(ctypes.c_void_p(comms[r])))
th = [threading.Thread(target=body, args=(r,)) for r in range(p)]
[t.start() for t in th]
[t.join() for t in th]
kt, bad, phases = res[0]
print("p=%d n=2^%d E=%d kind=%d bad=%s" % (p, log2n, E, kind, [r[1] for r in res]))
print("  kernels(ms/sort):", "  ".join("%s %.3f/%d" % (k, v[0] / 2, v[1] // 2) for k, v in kt.items() if v[1]))
print("  phases(ms):", [(k, round(v * 1e3, 2)) for k, v in phases if v > 2e-4])
