"""Opt-in GPU checks of NON-DEFAULT switches (MPSORT_TEST_CANDIDATES=1), so that the default suite states what the
shipped paths do. Round 1 wrote candidates here without a GPU at hand; round 2 measured them all
(profiles/r02_call1..., r02_call_n2..., r02_call_n8...): the bucket merge, the persistent record pass, the pipelined pack and the
split copies were slower or no faster and are gone; the peer-memory splitter kernel, the fused pack + exchange kernel and the
five-pass hybrid depth became defaults (one process per GPU). What is left here:

  MPSORT_PEER_SPLITTER=1     the splitter kernel also for rank THREADS of one process (default there: all-reduce per level)
  MPSORT_NO_FUSED_PACK=1 / MPSORT_NO_PEER_SPLITTER=1 over NCCL processes (needs >= 2 GPUs): the fallbacks of the defaults
  randomised cases as NCCL ranks (needs >= 2 GPUs)
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MPSORT_TEST_CANDIDATES") != "1",
                                 reason="candidate kernels: set MPSORT_TEST_CANDIDATES=1")]

WORKER = r"""
import sys, os
sys.path.insert(0, os.path.join(%(root)r, "mp-sort_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, ctypes
import mpsort
from mpsort import _capi as C
import mpsort_oracle as O
lib = C.lib
ok = True
for p, n, E, kind, signed in ((4, 60000, 16, 0, 0), (8, 40000, 16, 1, 0), (3, 50000, 48, 2, 1), (2, 70000, 24, 3, 0), (8, 30000, 16, 0, 0)):
    sizes = [n + 17 * r for r in range(p)]
    outsizes = sizes[::-1]
    recs = [O.generate(sizes[r], E, kind, 0x5EED0001, r, p) for r in range(p)]
    desc = O.Desc(0, 8, 1, signed, 0)
    exp = O.numpy_sort(recs, desc, outsizes)
    outs = [np.zeros((outsizes[r], E), np.uint8) for r in range(p)]
    d = C.RadixDesc(0, 8, 1, signed, 0)
    lib.mpsort_mpi_unset_options(-1)
    lib.mpsort_mpi_set_options(C.MPSORT_DISABLE_GATHER_SORT)
    def work(comm):
        r = comm.rank
        lib.mpsort_mpi_newarray_desc_impl(recs[r].ctypes.data, len(recs[r]), outs[r].ctypes.data, len(outs[r]), E,
                                          ctypes.byref(d), comm.handle, 0, b"cand")
        return C.last_stats(comm.handle, p)
    stats = mpsort.run_local(p, work)
    good = all(np.array_equal(outs[r], exp[r]) for r in range(p))
    print("p", p, "E", E, "kind", kind, "->", good, "merge tiles", [s["second_sort_merge_tiles"] for s in stats])
    ok &= good and all(s["second_sort_merge_tiles"] > 0 for s in stats)
print("CANDIDATE OK" if ok else "CANDIDATE FAILED")
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("switch", ["MPSORT_PEER_SPLITTER"])
def test_candidate_matches_oracle(switch):
    """bit-exact against the oracle on 2-8 rank threads, 16/24/48-byte records, all three key kinds"""
    env = dict(os.environ, **{switch: "1"})
    rc = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, timeout=900,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"CANDIDATE OK" in rc.stdout, rc.stdout.decode()[-4000:]


WORKER_PROPS = r"""
import sys, os
sys.path.insert(0, os.path.join(%(root)r, "mp-sort_b200"))
import ctypes
import mpsort
from mpsort import _capi as C
lib = C.lib
p, n, E, kind = 4, 1 << %(log2n)d, %(E)d, %(kind)d
desc = C.RadixDesc(0, 8, 1, 1 if kind == 2 else 0, 0)
res = [None] * p
def work(comm):
    r = comm.rank
    buf = lib.mpsort_util_dev_malloc(0, n * E)
    out = lib.mpsort_util_dev_malloc(0, n * E)
    lib.mpsort_util_generate(comm.handle, buf, n, E, kind, 0x5EED0001)
    s1 = lib.mpsort_util_checksum(comm.handle, buf, n * E)
    lib.mpsort_mpi_newarray_desc_impl(buf, n, out, n, E, ctypes.byref(desc), comm.handle, 0, b"cand")
    st = C.last_stats(comm.handle, p)
    s2 = lib.mpsort_util_checksum(comm.handle, out, n * E)
    fl = (ctypes.c_uint64 * 2)()
    bad = lib.mpsort_util_check_sorted(comm.handle, out, n, E, ctypes.byref(desc), 1, 8, fl)
    res[r] = (s1, s2, bad, fl[0], fl[1], st)
mpsort.run_local(p, work)
mask = (1 << 64) - 1
ok = (sum(x[0] for x in res) & mask) == (sum(x[1] for x in res) & mask) and all(x[2] == 0 for x in res)
ok = ok and all(res[r - 1][4] <= res[r][3] for r in range(1, p))
print("phases", [x[5]["exchange_phases"] for x in res], "merge tiles", [x[5]["second_sort_merge_tiles"] for x in res],
      "record mode", [x[5]["record_mode"] for x in res],
      "passes", [x[5]["first_sort_passes"] for x in res])
ok = ok and all(x[5]["exchange_phases"] == 2 for x in res)
if %(passes)d:
    # which key bytes vary is a per-rank fact (the swapped keys of a neighbour decide): at least one rank
    ok = ok and any(x[5]["first_sort_passes"] == %(passes)d for x in res)
print("CANDIDATE OK" if ok else "CANDIDATE FAILED")
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("E,kind,extra", [(16, 0, {"MPSORT_PEER_SPLITTER": "1"})])
def test_candidates_at_2_22_records_per_rank_by_properties(E, kind, extra):
    """4 rank threads x 2^22 records, exchange in two parts: global order, tie order (tags), checksum
    of checksums -- with the pipelined pack (index mode) or the peer splitter kernel switched on"""
    env = dict(os.environ, MPSORT_EXCHANGE_PHASES="2", **extra)
    rc = subprocess.run([sys.executable, "-c", WORKER_PROPS % {"root": ROOT, "E": E, "kind": kind, "log2n": 22, "passes": 0}],
                        env=env, timeout=900, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"CANDIDATE OK" in rc.stdout, rc.stdout.decode()[-4000:]


@pytest.mark.parametrize("switch", ["MPSORT_NO_FUSED_PACK", "MPSORT_NO_PEER_SPLITTER"])
def test_candidates_over_nccl_processes(switch):
    """one process per GPU (mapped peer memory): the NCCL worker of the default suite (48-byte and
    16-byte records, uneven sizes, all tunings) with one of the defaults switched off"""
    sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
    from mpsort import _capi as C
    ngpu = min(C.lib.mpsort_util_device_count(), 8)
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519", **{switch: "1"})
    rc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu),
                         "--master-addr", "127.0.0.1", "--master-port", "29519",
                         os.path.join(ROOT, "tests", "nccl_worker.py")], env=env, timeout=600,
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"NCCL PARITY OK" in rc.stdout, rc.stdout.decode()[-4000:]


@pytest.mark.parametrize("transport", [[], ["nccl"]], ids=["rank-threads", "nccl"])
def test_randomised_cases_on_the_gpu(transport):
    """tests/support/hostflow_fuzz.py (the generator of the CPU host-flow tests) against the REAL library: 300 random
    cases on rank threads of one GPU, and as NCCL ranks (threads of one process, one GPU each; cases with more ranks than
    GPUs are left out) -- shipped switches only, 1-4 exchange parts on tiny inputs -- byte for byte against the oracle's
    contract. Opt-in like the rest of this file: written without a GPU at hand."""
    if transport:
        sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
        from mpsort import _capi as C
        if C.lib.mpsort_util_device_count() < 2:
            pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, FUZZ_SWITCHES="shipped")
    env.pop("MPSORT_LIB", None)
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), "20261019", "300"] + transport,
                        env=env, timeout=1200, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert rc.returncode == 0 and b"FUZZ OK" in rc.stdout, rc.stdout.decode()[-4000:]
