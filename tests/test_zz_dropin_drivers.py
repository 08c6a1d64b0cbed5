"""The reference's OWN C drivers -- main-mpi.c and bench-mpi.c (its C smoke test and
its benchmark) -- compiled UNMODIFIED where they lie under /root/reference, with the reference's own
mpsort.h and mp-mpiu.h, against tests/dropin/mpi.h (`typedef struct mpsort_comm * MPI_Comm`) and linked
against the product library: the reference's declarations of mpsort_mpi_impl / mpsort_mpi_newarray_impl /
the option functions / mpsort_mpi_report_last_run / MPIU_Set_verbose_malloc ARE the exported symbols.
Each driver checks itself (checksum, local order, neighbour order) and exits non-zero or prints `fail`.

CPU part: linked against the mock device build (host flow only, rank threads). GPU part: the binaries that
build() made from /root/reference while it was there (tests/dropin/_build/, shipped with the snapshot),
against libmpsort-b200.so: one rank, rank threads on one GPU, and one process per GPU when there are several.

Collected last on purpose (file name): written in a GPU-less session, first GPU run is the round-end suite."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

DROPIN = os.path.join(ROOT, "tests", "dropin")
BUILD = os.path.join(DROPIN, "_build")
REF = "/root/reference"

CASES = [("main-mpi", ["100000"]), ("main-mpi", ["7"]), ("main-mpi", ["0"]),
         ("bench-mpi", ["-g", "200000"]), ("bench-mpi", ["-s", "50000"]), ("bench-mpi", ["-G", "3000"]),
         ("bench-mpi", ["-g", "-s", "-b", "4", "30000"]), ("bench-mpi", ["-b", "10", "-g", "100000"]), ("bench-mpi", ["1"])]


def run_driver(exe, args, threads=None, launcher=None, env=None, keep=()):
    env = dict(env or os.environ)
    env = {k: v for k, v in env.items() if not k.startswith("MPSORT_") and (not k.startswith("MOCK_") or k in keep)}
    if threads:
        env["MPSORT_DROPIN_THREADS"] = str(threads)
    cmd = [exe] + args
    if launcher:
        cmd = [sys.executable, os.path.join(DROPIN, "launch.py"), "-np", str(launcher)] + cmd
    rc = subprocess.run(cmd, env=env, timeout=600, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    out = rc.stdout.decode()
    assert rc.returncode == 0 and "fail" not in out and "inconsistent" not in out, "%s\n%s" % (" ".join(cmd), out[-3000:])
    return out


@pytest.fixture(scope="module")
def mock_drivers():
    if not os.path.isdir(REF):
        pytest.skip("the reference sources are not on this box")
    sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
    import hostmock
    hostmock.build()
    subprocess.run(["make", "-C", DROPIN, "mock"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return BUILD


@pytest.mark.parametrize("driver,args", CASES, ids=lambda v: v if isinstance(v, str) else "_".join(v))
def test_reference_drivers_host_flow_on_the_mock(mock_drivers, driver, args):
    exe = os.path.join(mock_drivers, driver + ".mock")
    run_driver(exe, args)                      # one rank
    out = run_driver(exe, args, threads=5)     # five rank threads
    run_driver(exe, args, threads=3, env=dict(os.environ, MOCK_ASYNC="1"), keep=("MOCK_ASYNC",))    # stream operations deferred
    if driver == "bench-mpi":
        assert "NTask = 5" in out and "MPSort total time" in out
    if args[-1] not in ("0", "1", "7", "3000"):
        assert "FirstSort:" in out and "SecondSort:" in out        # mpsort_mpi_report_last_run, the reference's phase names


@pytest.mark.gpu
@pytest.mark.parametrize("driver,args", CASES, ids=lambda v: v if isinstance(v, str) else "_".join(v))
def test_reference_drivers_on_the_gpu(driver, args):
    exe = os.path.join(BUILD, driver)
    if not os.path.exists(exe):
        pytest.skip("tests/dropin/_build/%s was not built (needs /root/reference at build time)" % driver)
    run_driver(exe, args)
    out = run_driver(exe, args, threads=4)
    if driver == "bench-mpi":
        assert "NTask = 4" in out and "MPSort total time" in out


@pytest.mark.gpu
def test_reference_drivers_one_process_per_gpu():
    sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
    from mpsort import _capi as C
    ngpu = min(C.lib.mpsort_util_device_count(), 8)
    exe = os.path.join(BUILD, "bench-mpi")
    if ngpu < 2 or not os.path.exists(exe):
        pytest.skip("needs >= 2 GPUs and the prebuilt drivers")
    out = run_driver(exe, ["-g", "4000000"], launcher=ngpu)
    assert "NTask = %d" % ngpu in out
    run_driver(os.path.join(BUILD, "main-mpi"), ["1000000"], launcher=ngpu)
