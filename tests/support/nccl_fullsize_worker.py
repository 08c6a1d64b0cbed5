"""Worker of test_nccl_full_size_bytes_equal_the_reference: one process per GPU (torchrun env), the production
transport at FULL size -- 2^LOG2N 16-byte records per GPU (default 28: what bench.py sorts), byte for byte against
the unmodified reference (oracle/_ref/bench16, R MPI-shim ranks on the host cores, run by rank 0 while the others
wait). The reference's R ranks generate R chunks; GPU g holds chunks g*R/G .. (g+1)*R/G - 1 in order, so the
rank-order concatenation -- and with it the output contract -- is the same on both sides."""
import ctypes
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402
import mpsort_oracle as O  # noqa: E402

lib = C.lib


def main():
    comm = mpsort.Comm.from_env()
    G, g = comm.size, comm.rank
    log2n = int(os.environ.get("LOG2N", "28"))
    kind = int(os.environ.get("KIND", "0"))
    E = 16
    cores = os.cpu_count() or 1
    R = G
    while R * 2 <= min(cores, 32):
        R *= 2
    n = 1 << log2n                       # per GPU
    per = n * G // R                     # per reference rank
    cpg = R // G                         # reference chunks per GPU
    buf = lib.mpsort_util_dev_malloc(comm.device, n * E)
    out = lib.mpsort_util_dev_malloc(comm.device, n * E)
    for j in range(cpg):
        lib.mpsort_util_generate_as(comm.handle, ctypes.c_void_p(buf + j * per * E), per, E, kind, 0x5EED0001, g * cpg + j, R)
    desc = C.RadixDesc(0, 8, 1, 0, 0)
    lib.mpsort_mpi_newarray_desc_impl(buf, n, out, n, E, ctypes.byref(desc), comm.handle, 0, b"nccl_fullsize")
    st = C.last_stats(comm.handle, G)
    outdir = comm.bcast(tempfile.mkdtemp(prefix="mpsort_fullsize_") if g == 0 else None)
    secs = None
    if g == 0:
        secs = O.run_bench16(R, per, elsize=E, kind=kind, reps=1, timeout=3000, outdir=outdir)["best_seconds"]
    comm.barrier()
    good = True
    got = np.empty((per, E), np.uint8)
    for j in range(cpg):
        lib.mpsort_util_memcpy(comm.device, got.ctypes.data, ctypes.c_void_p(out + j * per * E), per * E)
        exp = np.fromfile(os.path.join(outdir, "out.%d" % (g * cpg + j)), dtype=np.uint8).reshape(per, E)
        good = good and bool(np.array_equal(got, exp))
    allgood = comm.allgather(good)
    stats = comm.allgather((st["p2p_exchange"], st["exchange_phases"], st["second_sort_merge_tiles"], st["own_slices_in_place"], st["bytes_sent_remote"]))
    comm.barrier()
    if g == 0:
        shutil.rmtree(outdir, ignore_errors=True)
        print("kind", kind, "records per GPU 2^%d" % log2n, "GPUs", G, "reference ranks", R, "%.1f s" % secs, "->", allgood,
              "(p2p, parts, merge tiles, own slices in place, bytes sent)", stats)
        print("NCCL FULL SIZE OK" if all(allgood) else "NCCL FULL SIZE FAILED")
    lib.mpsort_util_dev_free(comm.device, buf)
    lib.mpsort_util_dev_free(comm.device, out)
    comm.destroy()
    return 0 if all(allgood) else 1


if __name__ == "__main__":
    sys.exit(main())
