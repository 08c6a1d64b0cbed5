"""Python face of the oracle (TEST INFRASTRUCTURE ONLY -- see mpsort_oracle.h).

Three independent checkers of the distributed sort, all CPU:

  numpy_sort   the output contract in one sentence: the stable sort of
               the rank-order concatenation by the unsigned radix key, cut into the
               requested output sizes.
  c_sort       oracle/mpsort_oracle.c: restatement of the reference's algorithm
               (bisection splitters, greedy layout, two stable local sorts).
  ref_sort     the UNMODIFIED reference (oracle/_ref/ref_driver, built from
               /root/reference by oracle/Makefile) under the single-host MPI shim.

Records are passed as uint8 arrays of shape [n, elsize] (use `as_bytes`).
Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may
import this module.
"""
import ctypes
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libmpsort_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = os.environ.get("MPSORT_REFERENCE_SRC", "/root/reference")

DISABLE_SPARSE_ALLTOALLV = 1 << 1
DISABLE_GATHER_SORT = 1 << 3
REQUIRE_GATHER_SORT = 1 << 4
REQUIRE_SPARSE_ALLTOALLV = 1 << 6
TUNING_BITS = {
    "DISABLE_SPARSE_ALLTOALLV": DISABLE_SPARSE_ALLTOALLV,
    "DISABLE_GATHER_SORT": DISABLE_GATHER_SORT,
    "REQUIRE_GATHER_SORT": REQUIRE_GATHER_SORT,
    "REQUIRE_SPARSE_ALLTOALLV": REQUIRE_SPARSE_ALLTOALLV,
}


class Desc(ctypes.Structure):
    """struct oracle_desc: offset, width, nwords, is_signed, raw"""
    _fields_ = [("offset", ctypes.c_size_t), ("width", ctypes.c_uint32), ("nwords", ctypes.c_uint32),
                ("is_signed", ctypes.c_int32), ("raw", ctypes.c_int32)]

    def astuple(self):
        return (self.offset, self.width, self.nwords, self.is_signed, self.raw)


def build(ref=True):
    """make -C oracle (the C restatement always; oracle/_ref when the reference
    sources are present)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "REF=" + REFERENCE_SRC] + targets, check=True,
                   stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _lib = ctypes.CDLL(ORACLE_SO)
        _lib.oracle_radix_sort.restype = None
        _lib.oracle_radix_sort.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(Desc)]
        _lib.oracle_mpsort.restype = ctypes.c_int
        _lib.oracle_mpsort.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
                                       ctypes.c_size_t, ctypes.POINTER(Desc), ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int)]
        _lib.oracle_checksum.restype = ctypes.c_uint64
        _lib.oracle_checksum.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _lib.oracle_generate.restype = None
        _lib.oracle_generate.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    return _lib


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("mpirun-shim", "ref_driver", "bench16"))


def as_bytes(a):
    """any C-contiguous 1-d (struct) array -> uint8 view [n, itemsize]"""
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(len(a), a.dtype.itemsize)


def key_words(rec, desc):
    """radix words of every record as unsigned integers, list of nwords arrays
    (word nwords-1 most significant); the binding.pyx:81-121 rule."""
    offset, width, nwords, is_signed, _raw = desc.astuple()
    out = []
    for w in range(nwords):
        b = np.ascontiguousarray(rec[:, offset + w * width: offset + (w + 1) * width])
        v = b.view("<u%d" % width).reshape(len(rec)).astype(np.uint64)
        if is_signed:
            v = v ^ np.uint64(1 << (8 * width - 1))
        out.append(v)
    return out


def numpy_sort(recs, desc, outsizes=None):
    """The contract: stable sort of the rank-order concatenation, cut by outsizes."""
    if outsizes is None:
        outsizes = [len(r) for r in recs]
    elsize = recs[0].shape[1]
    allrec = np.concatenate(recs, axis=0) if len(recs) else np.zeros((0, elsize), np.uint8)
    assert sum(outsizes) == len(allrec)
    order = np.lexsort(tuple(key_words(allrec, desc)))   # lexsort is stable; last key is primary
    s = allrec[order]
    cuts = np.cumsum([0] + list(outsizes))
    return [s[cuts[i]:cuts[i + 1]].copy() for i in range(len(outsizes))]


def generate(n, elsize, kind, seed, rank, nranks):
    out = np.zeros((n, elsize), np.uint8)
    lib().oracle_generate(out.ctypes.data, n, elsize, kind, seed, rank, nranks)
    return out


def checksum(rec):
    rec = np.ascontiguousarray(rec)
    return int(lib().oracle_checksum(rec.ctypes.data, rec.nbytes))


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic), oracle/synth.h:mix64"""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def multiset_hash(rec):
    """(sum, xor) over the records of h(record): the order-independent whole-record hash the
    property checks use (mpsort_util_multiset_hash). h folds the record's 8-byte little-endian
    words (the last one zero-padded) through mix64 from a fixed seed."""
    rec = np.ascontiguousarray(rec)
    n, elsize = rec.shape
    nw = (elsize + 7) // 8
    padded = np.zeros((n, nw * 8), np.uint8)
    padded[:, :elsize] = rec
    words = padded.view("<u8")
    h = np.full(n, 0x243F6A8885A308D3, dtype=np.uint64)
    for k in range(nw):
        h = _mix64(h ^ words[:, k])
    with np.errstate(over="ignore"):
        s = int(np.add.reduce(h, dtype=np.uint64)) if n else 0
    x = int(np.bitwise_xor.reduce(h)) if n else 0
    return s, x


def c_radix_sort(rec, desc):
    out = np.ascontiguousarray(rec).copy()
    lib().oracle_radix_sort(out.ctypes.data, len(out), out.shape[1], ctypes.byref(desc))
    return out


def c_sort(recs, desc, outsizes=None, options=0, inplace=False):
    """oracle_mpsort. Returns (outs, info); info has sendcounts [p,p], rounds, nleaders."""
    p = len(recs)
    if outsizes is None:
        outsizes = [len(r) for r in recs]
    elsize = recs[0].shape[1]
    ins = [np.ascontiguousarray(r).copy() for r in recs]
    outs = ins if inplace else [np.zeros((outsizes[i], elsize), np.uint8) for i in range(p)]
    bases = (ctypes.c_void_p * p)(*[a.ctypes.data for a in ins])
    obases = (ctypes.c_void_p * p)(*[a.ctypes.data for a in outs])
    n = (ctypes.c_size_t * p)(*[len(a) for a in ins])
    on = (ctypes.c_size_t * p)(*outsizes)
    sc = (ctypes.c_int64 * (p * p))()
    rounds = ctypes.c_int(0)
    nlead = ctypes.c_int(0)
    rc = lib().oracle_mpsort(p, bases, n, obases, on, elsize, ctypes.byref(desc), options, sc,
                             ctypes.byref(rounds), ctypes.byref(nlead))
    if rc != 0:
        raise RuntimeError("oracle_mpsort failed with code %d" % rc)
    info = {"sendcounts": np.array(list(sc), dtype=np.int64).reshape(p, p), "rounds": rounds.value,
            "nleaders": nlead.value}
    return outs, info


def ref_sort(recs, desc, outsizes=None, options=0, inplace=False, timeout=300):
    """Run the unmodified reference under the MPI shim, one process per rank."""
    if not have_ref():
        raise RuntimeError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
    p = len(recs)
    if outsizes is None:
        outsizes = [len(r) for r in recs]
    elsize = recs[0].shape[1]
    d = tempfile.mkdtemp(prefix="mpsort_ref_")
    try:
        for r in range(p):
            np.ascontiguousarray(recs[r]).tofile(os.path.join(d, "in.%d" % r))
            with open(os.path.join(d, "outn.%d" % r), "w") as f:
                f.write("%d\n" % outsizes[r])
        offset, width, nwords, is_signed, raw = desc.astuple()
        cmd = [os.path.join(REF_DIR, "mpirun-shim"), "-np", str(p), os.path.join(REF_DIR, "ref_driver"), d,
               str(elsize), str(offset), str(width), str(nwords), str(is_signed), str(raw), str(options),
               "1" if inplace else "0"]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        if proc.returncode != 0:
            raise RuntimeError("reference run failed (%d): %s" % (proc.returncode, proc.stderr.decode()[-2000:]))
        outs = []
        for r in range(p):
            n = len(recs[r]) if inplace else outsizes[r]
            outs.append(np.fromfile(os.path.join(d, "out.%d" % r), dtype=np.uint8).reshape(n, elsize))
        return outs
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_bench16(np_ranks, n_per_rank, elsize=16, kind=0, reps=1, timeout=900, warmups=0, budget_s=0.0, outdir=None, total=0):
    """The reference's CPU run of the benchmark workload (bench-mpi's sibling for struct
    records). Returns dict(best_seconds, mean_seconds, records_per_second, phases, np, ...).
    warmups: repetitions left out of mean_seconds; budget_s: stop repeating after that many
    seconds of sorting; outdir: every rank writes its sorted output there as out.<rank>;
    total > 0: the ranks share `total` records as evenly as possible (n_per_rank is ignored)."""
    cmd = [os.path.join(REF_DIR, "mpirun-shim"), "-np", str(np_ranks), os.path.join(REF_DIR, "bench16"),
           "-k", str(kind), "-e", str(elsize), "-r", str(reps), "-w", str(warmups), "-t", "%g" % budget_s]
    if outdir:
        cmd += ["-o", outdir]
    if total:
        cmd += ["-T", str(int(total))]
    cmd.append(str(n_per_rank))
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    if proc.returncode != 0:
        raise RuntimeError("bench16 failed (%d): %s" % (proc.returncode, proc.stderr.decode()[-2000:]))
    res = {"np": np_ranks, "phases": {}, "totals": []}
    for line in proc.stdout.decode().splitlines():
        if line.startswith("MPSort total time:"):
            res["totals"].append(float(line.split(":", 1)[1]))
            continue
        if line.startswith("BENCH16"):
            for tok in line.split()[1:]:
                k, v = tok.split("=")
                res[k] = float(v) if "." in v else int(v)
        elif ":" in line:
            k, v = line.split(":", 1)
            try:
                res["phases"][k.strip()] = float(v)
            except ValueError:
                pass
    return res
