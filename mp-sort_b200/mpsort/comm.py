"""Communicators for the Python surface.

The reference takes an mpi4py communicator (binding.pyx:168-180) and its tests use
comm.rank/size/allgather/allreduce/bcast/scatter/barrier (test_mpsort.py:8-20).
There is no MPI on the target, so `Comm` wraps the library's own `mpsort_comm_t`
(NCCL over NVLink, or an in-process group of rank threads) and offers those same
methods, implemented with the library's small host collectives.
"""
import ctypes
import os
import pickle
import threading
import time

from . import _capi as C

lib = C.lib


class Comm(object):
    """One rank's view of a communicator (same method names as mpi4py's Comm)."""

    def __init__(self, handle, owner=True):
        self.handle = ctypes.c_void_p(handle) if not isinstance(handle, ctypes.c_void_p) else handle
        self._owner = owner
        self.rank = lib.mpsort_comm_rank(self.handle)
        self.size = lib.mpsort_comm_size(self.handle)
        self.device = lib.mpsort_comm_device(self.handle)
        self.numa = None

    # ---- construction --------------------------------------------------
    @classmethod
    def self(cls, device=0):
        """size-1 communicator on one GPU (no NCCL)."""
        return cls(lib.mpsort_comm_self(device))

    @classmethod
    def local_group(cls, size, devices=None):
        """`size` ranks driven by `size` threads of this process; devices[i] is the
        CUDA device of rank i (default: round-robin over the visible devices)."""
        ndev = lib.mpsort_util_device_count()
        if ndev < 1:
            raise RuntimeError("mpsort: no CUDA device available; there is no CPU fallback")
        if devices is None:
            devices = [r % ndev for r in range(size)]
        devs = (ctypes.c_int * size)(*devices)
        comms = (ctypes.c_void_p * size)()
        rc = lib.mpsort_comm_init_local_group(size, devs, comms)
        if rc != 0:
            raise RuntimeError("mpsort_comm_init_local_group failed with %d" % rc)
        return [cls(comms[r]) for r in range(size)]

    @classmethod
    def from_env(cls):
        """One process per GPU under torchrun / any launcher that exports RANK,
        WORLD_SIZE and LOCAL_RANK. The NCCL unique id travels through a file in
        $MPSORT_RENDEZVOUS_DIR (default /tmp) keyed by the launcher's pid, so no
        torch / MPI import is needed. Single node (the 8 GPUs of one box)."""
        rank = int(os.environ.get("RANK", "0"))
        size = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        numa = bind_to_device_numa_node(local)
        if size <= 1:
            comm = cls.self(local)
            comm.numa = numa
            return comm
        uid = ctypes.create_string_buffer(C.MPSORT_UNIQUE_ID_BYTES)
        path = _rendezvous_path()
        if rank == 0:
            if lib.mpsort_comm_get_unique_id(uid) != 0:
                raise RuntimeError("ncclGetUniqueId failed")
            tmp = path + ".tmp.%d" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(uid.raw)
            os.rename(tmp, path)
        else:
            deadline = time.time() + float(os.environ.get("MPSORT_RENDEZVOUS_TIMEOUT", "300"))
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise RuntimeError("mpsort: rendezvous file %s did not appear" % path)
                time.sleep(0.01)
            with open(path, "rb") as f:
                uid.raw = f.read(C.MPSORT_UNIQUE_ID_BYTES)
        comm = cls(lib.mpsort_comm_init_rank(rank, size, uid, local))
        comm.numa = numa
        comm.barrier()
        if rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass
        return comm

    def destroy(self):
        if self._owner and self.handle:
            lib.mpsort_comm_destroy(self.handle)
            self.handle = None

    # ---- mpi4py-style host collectives ------------------------------------
    def barrier(self):
        lib.mpsort_comm_barrier(self.handle)

    Barrier = barrier

    def allgather(self, obj):
        if self.size == 1:
            return [obj]
        blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        mine = ctypes.c_uint64(len(blob))
        sizes = (ctypes.c_uint64 * self.size)()
        lib.mpsort_comm_allgather_host(self.handle, ctypes.byref(mine), sizes, 8)
        counts = (ctypes.c_size_t * self.size)(*list(sizes))
        total = sum(sizes)
        recv = ctypes.create_string_buffer(max(total, 1))
        send = ctypes.create_string_buffer(blob, max(len(blob), 1))
        lib.mpsort_comm_allgatherv_host(self.handle, send, len(blob), recv, counts)
        out, off = [], 0
        raw = recv.raw
        for r in range(self.size):
            out.append(pickle.loads(raw[off:off + sizes[r]]))
            off += sizes[r]
        return out

    def allreduce(self, value, op=None):
        vals = self.allgather(value)
        if op is not None:
            acc = vals[0]
            for v in vals[1:]:
                acc = op(acc, v)
            return acc
        acc = vals[0]
        for v in vals[1:]:
            acc = acc + v
        return acc

    def bcast(self, obj, root=0):
        return self.allgather(obj if self.rank == root else None)[root]

    def scatter(self, objs, root=0):
        return self.allgather(objs if self.rank == root else None)[root][self.rank]

    def gather(self, obj, root=0):
        vals = self.allgather(obj)
        return vals if self.rank == root else None

    def __repr__(self):
        return "<mpsort.Comm rank %d of %d on cuda:%d>" % (self.rank, self.size, self.device)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_device_numa_node(device):
    """One process per GPU: run this process on the CPUs of the NUMA node its GPU hangs off, so that the
    host buffers it allocates and pins afterwards (first touch) are local to that GPU's PCIe root. With eight
    ranks staging 4 GiB each way, buffers that all land on one socket leave half of the GPUs copying across
    the socket interconnect. MPSORT_NO_NUMA_BIND=1 leaves the affinity alone. Returns a short description
    (None when nothing was done: single node, no sysfs entry, or an affinity the launcher already narrowed)."""
    if os.environ.get("MPSORT_NO_NUMA_BIND"):
        return None
    try:
        buf = ctypes.create_string_buffer(64)
        if lib.mpsort_util_device_pci_bus_id(device, buf, 64) != 0:
            return None
        bus = buf.value.decode().lower()
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        have = os.sched_getaffinity(0)
        if len(have) < (os.cpu_count() or 0):
            return "affinity already set by the launcher (%d cpus); left alone" % len(have)
        use = cpus & have
        if not use or use == have:
            return None
        os.sched_setaffinity(0, use)
        return "cuda:%d %s -> NUMA node %d (%d cpus)" % (device, bus, node, len(use))
    except (OSError, ValueError, AttributeError):
        return None


_rdzv_seq = [0]


def _rendezvous_path():
    explicit = os.environ.get("MPSORT_RENDEZVOUS_FILE")
    _rdzv_seq[0] += 1
    if explicit:
        return "%s.%d" % (explicit, _rdzv_seq[0])
    # all workers of one launch share their parent (the torchrun agent / the shell)
    ppid = os.getppid()
    start = "0"
    try:
        with open("/proc/%d/stat" % ppid) as f:
            start = f.read().rsplit(")", 1)[1].split()[19]
    except (OSError, IndexError):
        pass
    d = os.environ.get("MPSORT_RENDEZVOUS_DIR", "/tmp")
    return os.path.join(d, "mpsort_rdzv_%d_%s_%s_%d" % (ppid, start, os.environ.get("MASTER_PORT", "0"), _rdzv_seq[0]))


_world = [None]


def world():
    """The default communicator (the reference's MPI.COMM_WORLD, binding.pyx:169-170):
    built from the launcher's environment on first use."""
    if _world[0] is None:
        _world[0] = Comm.from_env()
    return _world[0]


def run_local(size, fn, devices=None, timeout=600):
    """Run fn(comm) on `size` rank-threads of an in-process group and return the list
    of results in rank order (how the reference's `mpirun -n 4 pytest` cases run on a
    box with fewer GPUs). Exceptions are re-raised in the caller."""
    comms = Comm.local_group(size, devices)
    results = [None] * size
    errors = [None] * size

    def body(r):
        try:
            results[r] = fn(comms[r])
        except BaseException as e:  # noqa: B902 -- reported below
            errors[r] = e

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(size)]
    for t in threads:
        t.start()
    deadline = time.time() + timeout
    # poll: a rank that raised while the others sit in a collective must not cost the whole timeout.
    # After its exception the others get a short grace period (they may be about to finish), then the
    # group is abandoned: its threads are daemons blocked inside the library, its communicators are
    # NOT destroyed (a destroy would wait for them) -- the group is unusable from then on.
    failed_at = None
    while any(t.is_alive() for t in threads):
        now = time.time()
        if failed_at is None and any(e is not None for e in errors):
            failed_at = now
        if now > deadline or (failed_at is not None and now > failed_at + 5.0):
            break
        time.sleep(0.002)
    hung = [t for t in threads if t.is_alive()]
    for e in errors:
        if e is not None:
            if not hung:
                for c in comms:
                    c.destroy()
            raise e
    if hung:
        raise RuntimeError("mpsort.run_local: %d rank threads did not finish in %g s" % (len(hung), timeout))
    for c in comms:
        c.destroy()
    return results
