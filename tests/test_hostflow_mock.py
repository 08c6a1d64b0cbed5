"""Host orchestration of libmpsort-b200.so WITHOUT a GPU: mpsort_host.c / mpsort_comm.c /
mpsort_layout.c / mpsort_util.c, compiled unchanged, run against tests/native/mock_device.c
(kernel ABI, CUDA runtime and NCCL restated as plain loops / no-ops). What passes here is the
C host logic -- phases, splitter targets, layout over virtual ranks, exchange parts, gather
path, callback entry points, arena bookkeeping, the Cython binding -- NOT the kernels: those
are checked by the same test files on the B200 (`-m gpu`) through the real library.

The GPU test files are re-run in a subprocess with MPSORT_LIB pointing at the mock build, so
every host-side path the GPU suite drives is driven here too, and a host bug shows up on the
CPU box instead of at the end of a round."""
import json
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
import hostmock  # noqa: E402


@pytest.fixture(scope="module")
def mock_env():
    so = hostmock.build()
    env = dict(os.environ, MPSORT_LIB=so, MPSORT_ALLOW_MOCK_DEVICE="1")
    for k in list(env):
        if k.startswith("MPSORT_") and k not in ("MPSORT_LIB", "MPSORT_ALLOW_MOCK_DEVICE"):
            del env[k]
    return env


def run_py(env, code_or_args, timeout=900, **extra):
    args = [sys.executable] + (code_or_args if isinstance(code_or_args, list) else ["-c", code_or_args])
    return subprocess.run(args, cwd=ROOT, env=dict(env, **extra), timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def test_mock_provides_every_symbol_the_host_code_uses(mock_env):
    """linked with -z defs: a new CUDA / NCCL / kernel-ABI call in the host files fails the build here"""
    assert os.path.exists(mock_env["MPSORT_LIB"])


# the full-size tests take minutes as CPU loops; their host flow is the same as the 2^22 cases below
BIG = ["tests/test_gpu_parity.py::test_full_size_config_b_by_properties",
       "tests/test_gpu_parity.py::test_full_size_config_b_bytes_equal_the_reference",
       "tests/test_gpu_parity.py::test_full_size_configs_4_and_5_bytes_equal_the_reference",
       "tests/test_gpu_parity.py::test_large_particles48_and_mostly_sorted_by_properties"]


def test_gpu_suite_host_flow_on_the_mock(mock_env):
    """tests/test_python_api.py (the reference's own test file re-hosted) and test_zz_callback_api.py, `-m gpu`, against
    the mock build, every stream operation run at once (test_gpu_parity.py joins them in
    test_gpu_suite_host_flow_with_deferred_stream_operations below)"""
    args = ["-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_python_api.py", "tests/test_zz_callback_api.py"]
    rc = run_py(mock_env, args, timeout=1500)
    out = rc.stdout.decode()
    assert rc.returncode == 0, out[-4000:]
    m = re.search(r"(\d+) passed", out)
    assert m and int(m.group(1)) >= 60 and "failed" not in out, out[-2000:]


def _candidates():
    import test_zz_candidates as T
    return T


@pytest.mark.parametrize("extra", [{}, {"MPSORT_EXCHANGE_PHASES": "2"}, {"MPSORT_EXCHANGE_PHASES": "3"},
                                   {"MPSORT_NO_MERGE": "1"}, {"MPSORT_NO_REC16": "1", "MPSORT_NO_REBASE": "1"}, {"MPSORT_PEER_SPLITTER": "1"},
                                   {"MPSORT_NO_HYBRID": "1", "MPSORT_NO_HIST4": "1"}],
                         ids=lambda e: "+".join(sorted(e)) or "default")
def test_switches_leave_the_result_bit_exact(mock_env, extra):
    """2-8 rank threads, 16/24/48-byte records, the three key kinds, uneven sizes: the oracle's bytes
    whatever host-side switch is set"""
    code = _candidates().WORKER % {"root": ROOT}
    if "MPSORT_NO_MERGE" in extra:      # the second sort is then a radix re-sort: no merge tiles to count
        code = code.replace('all(s["second_sort_merge_tiles"] > 0 for s in stats)', 'all(s["second_sort_passes"] > 0 for s in stats)')
        assert "second_sort_passes" in code
    rc = run_py(mock_env, code, **extra)
    assert rc.returncode == 0 and b"CANDIDATE OK" in rc.stdout, rc.stdout.decode()[-4000:]


@pytest.mark.parametrize("E,kind,extra", [(16, 0, {}), (48, 2, {}), (48, 2, {"MPSORT_EXCHANGE_PHASES": "3"}),
                                          (24, 3, {"MPSORT_EXCHANGE_PHASES": "4"})])
def test_exchange_in_parts_at_2_22_records_per_rank(mock_env, E, kind, extra):
    """4 rank threads x 2^22 records (the smallest size at which the exchange is cut into parts): order, tie order,
    checksum of checksums, and that the parts were really taken"""
    code = _candidates().WORKER_PROPS % {"root": ROOT, "E": E, "kind": kind, "log2n": 22, "passes": 0}
    if extra.get("MPSORT_EXCHANGE_PHASES", "2") != "2":
        code = code.replace('x[5]["exchange_phases"] == 2', 'x[5]["exchange_phases"] == %s' % extra["MPSORT_EXCHANGE_PHASES"])
    rc = run_py(mock_env, code, **dict({"MPSORT_EXCHANGE_PHASES": "2"}, **extra))
    assert rc.returncode == 0 and b"CANDIDATE OK" in rc.stdout, rc.stdout.decode()[-4000:]


@pytest.mark.parametrize("seed", ["", "3", "11"], ids=["immediate", "deferred-3", "deferred-11"])
def test_host_buffers_in_chunks(mock_env, seed):
    """a host input that arrives in chunks behind which the histogram pass runs, a host output that leaves
    range by range beside the fix-up / the merge of the next exchange part -- with tiny chunks, every stream operation
    run at once and deferred + randomly interleaved (a reader that does not wait for its chunk shows here)"""
    extra = {"MPSORT_CHUNK_MIN_BYTES": "4096", "MPSORT_CHUNK_BYTES": "1000000", "MPSORT_EXCHANGE_PHASES": "2",
             "MPSORT_PHASES_MIN_RECORDS": "1000"}
    if seed:
        extra["MOCK_ASYNC"] = seed
    rc = run_py(mock_env, [os.path.join(ROOT, "tests", "support", "chunk_worker.py")], **extra)
    assert rc.returncode == 0 and b"CHUNK OK" in rc.stdout, rc.stdout.decode()[-4000:]


def test_bench_gpu_arm_runs_end_to_end_on_the_mock(mock_env):
    """bench.py itself (not a stand-in for it): device-resident leg, e2e leg through mpsort.sort, verification,
    roofline bookkeeping, one JSON line. Numbers are meaningless here; the control flow is what is checked."""
    rc = subprocess.run([sys.executable, "bench.py", "--log2n", "16", "--steps", "2", "--warmup", "1", "--no-cpu-baseline",
                         "--preflight-log2n", "15", "--extra-steps", "1"],
                        cwd=ROOT, env=mock_env, timeout=600, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert rc.returncode == 0, rc.stderr.decode()[-3000:]
    lines = [l for l in rc.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 16) * 16 and d["roofline"]["bound"] == "hbm"
    # the pre-flight ran all three workloads bit-exactly against the oracle; the other two workloads were timed and verified
    pf = d["parity_preflight"]
    assert pf["ok"] and [c["workload"] for c in pf["cases"]] == ["uniform16", "mostly_sorted16", "particles48"]
    assert all(c["device_resident_equals_oracle"] and c["host_api_equals_oracle"] and c["generator_equals_oracle"] for c in pf["cases"])
    assert set(d["workloads"]) == {"mostly_sorted16", "particles48"} and all(w["value"] > 0 for w in d["workloads"].values())
    assert d["workloads"]["particles48"]["local_sort"]["record_mode"] is False


# ---- the NCCL transport (mpsort_comm_init_rank), one rank per thread of the worker ------------------
NCCL_WORKER = os.path.join(ROOT, "tests", "support", "nccl_threads_worker.py")


@pytest.mark.parametrize("p,extra,p2p", [(4, {}, 1), (2, {}, 1), (7, {"MOCK_NO_IPC": "1"}, 0), (4, {"MPSORT_NO_P2P": "1"}, 0),
                                         (3, {"MPSORT_P2P_PULL": "1"}, 1), (4, {"MPSORT_P2P_CE": "0"}, 1), (5, {"MPSORT_P2P_CE": "7"}, 1),
                                         (6, {"MPSORT_NO_PEER_SPLITTER": "1"}, 1), (3, {"MPSORT_NO_FUSED_PACK": "1"}, 1), (3, {"MOCK_NO_IPC": "1", "MPSORT_NO_PEER_SPLITTER": "1"}, 0)],
                         ids=lambda v: "+".join("%s=%s" % kv for kv in sorted(v.items())) or "default" if isinstance(v, dict) else str(v))
def test_nccl_transport_small_cases(mock_env, p, extra, p2p):
    """the cases of tests/nccl_worker.py (which needs >= 2 GPUs) over every exchange transport of mpsort_comm.c:
    DMA copies into mapped peer buffers (default), peer stores from the copy kernel, pull, grouped ncclSend/ncclRecv
    by choice and as the fallback when the buffers cannot be mapped; sparse and dense; the gather path; empty input"""
    rc = run_py(mock_env, [NCCL_WORKER, str(p)], EXPECT_P2P=str(p2p), **extra)
    assert rc.returncode == 0 and b"NCCL THREADS OK" in rc.stdout, rc.stdout.decode()[-4000:]


SMALLER = {"BIG_LOG2N": "18", "MPSORT_PHASES_MIN_RECORDS": "1000"}      # the candidates: same flow at 2^18 records per rank


@pytest.mark.parametrize("extra,phases", [({}, 4), (dict(SMALLER, MPSORT_NO_FUSED_PACK="1", MPSORT_EXCHANGE_PHASES="3"), 3),
                                          (dict(SMALLER, MPSORT_NO_PEER_SPLITTER="1", MPSORT_EXCHANGE_PHASES="2"), 2)],
                         ids=lambda v: "+".join("%s=%s" % kv for kv in sorted(v.items())) or "default" if isinstance(v, dict) else str(v))
def test_nccl_transport_exchange_in_parts(mock_env, extra, phases):
    """3 ranks x 2^22 records (16-byte uniform keys, 48-byte records with duplicates): the default of one process per
    GPU -- four exchange parts over mapped peer buffers with the merge of a part beside the transfer of the next, the
    fused pack for 48-byte records, the peer-memory splitter kernel, taken from 2^22 records per rank on -- and the
    switches that turn each of them off"""
    rc = run_py(mock_env, [NCCL_WORKER, "3", "big"], EXPECT_PHASES=str(phases), **extra)
    assert rc.returncode == 0 and b"NCCL THREADS OK" in rc.stdout, rc.stdout.decode()[-4000:]


# ---- rank threads under ThreadSanitizer ------------------------------------------------------------
def test_rank_threads_have_no_data_races(tmp_path):
    """tests/native/hostflow_threads.c + the host files + the mock, all built with -fsanitize=thread: rank threads
    of an in-process group and NCCL ranks as threads, two sorts each, 16- and 48-byte records with cross-rank ties"""
    exe = str(tmp_path / "hostflow_tsan")
    try:
        hostmock.compile_and_link(exe, flags=("-fsanitize=thread",), shared=False,
                                  extra_sources=[os.path.join(ROOT, "tests", "native", "hostflow_threads.c")])
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime on this box")
    env = {k: v for k, v in os.environ.items() if not k.startswith("MPSORT_")}
    # plain, and with deferred stream operations + the exchange in two parts: a peer's deferred copy into my receive buffer
    # and my merge of it are then ordered by nothing but the host code's own synchronisation
    deferred = dict(env, MOCK_ASYNC="1", MPSORT_EXCHANGE_PHASES="2", MPSORT_PHASES_MIN_RECORDS="1")
    for e, args in ((env, ["4", "30000", "16"]), (env, ["3", "20000", "48"]), (env, ["7", "5000", "24"]),
                    (deferred, ["4", "30000", "16"]), (deferred, ["3", "20000", "48"])):
        rc = subprocess.run([exe] + args, env=e, timeout=600, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        out = rc.stdout.decode()
        assert rc.returncode == 0 and "THREADS OK" in out and "ThreadSanitizer" not in out, out[-4000:]


# ---- the host files under AddressSanitizer + UBSan ---------------------------------------------------
def test_host_code_under_address_and_ub_sanitizers(tmp_path, mock_env):
    """the mock build again with -fsanitize=address,undefined ("device memory" is malloc here, so an arena slot that is
    too small for what a kernel's contract writes is a heap overflow): NCCL-transport cases and the callback entry
    points, sanitizer runtimes preloaded into the python worker"""
    so = hostmock.compile_and_link(str(tmp_path / "libmpsort-hostmock-asan.so"), flags=("-fsanitize=address,undefined", "-fno-omit-frame-pointer"))
    pre = [subprocess.run(["gcc", "-print-file-name=" + n], stdout=subprocess.PIPE).stdout.decode().strip() for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(x) and os.path.exists(x) for x in pre):
        pytest.skip("sanitizer runtimes not found")
    env = dict(mock_env, MPSORT_LIB=so, LD_PRELOAD=" ".join(pre), ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    rc = run_py(env, [NCCL_WORKER, "5"])
    out = rc.stdout.decode()
    assert rc.returncode == 0 and "NCCL THREADS OK" in out and "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
    # randomised cases incl. several sorts per communicator: receive buffers grow, peers' mappings of the old ones go stale
    # ... and with the stream operations deferred (a deferred copy into a buffer that was freed meanwhile would show here)
    for extra in ({}, {"MOCK_ASYNC": "5"}):
        rc = run_py(env, [os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), "77", "60", "nccl"], **extra)
        out = rc.stdout.decode()
        assert rc.returncode == 0 and "FUZZ OK" in out and "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]
    rc = run_py(env, ["-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-s", "tests/test_zz_callback_api.py",
                      "tests/test_gpu_parity.py", "-k", "callback or golden or other_key_shapes or radix_sort_desc"])
    out = rc.stdout.decode()
    assert rc.returncode == 0 and "AddressSanitizer" not in out and "runtime error" not in out, out[-4000:]


@pytest.mark.parametrize("extra,passes", [({}, 5), ({"MPSORT_NO_HYBRID5": "1"}, 4), ({"MPSORT_NO_HYBRID": "1"}, 7)])
def test_hybrid_depth_decisions(mock_env, extra, passes):
    """record mode, 2^22 keys with runs of 16 equal high parts that a fifth digit separates: five passes + fix-up
    (the predictor's choice), four when the fifth is switched off, seven plain passes without the hybrid; one long
    run through the work list; same bytes"""
    rc = run_py(mock_env, [os.path.join(ROOT, "tests", "support", "hybrid_depth_worker.py")], **extra)
    out = rc.stdout.decode()
    assert rc.returncode == 0 and ("passes=%d " % passes) in out and "equal=True" in out, out[-2000:]
    if passes != 7:
        assert "hybrid=1 long_runs=1" in out


# ---- randomised cases ---------------------------------------------------------------------------------
@pytest.mark.parametrize("seed,transport", [(20261017, []), (20261018, ["nccl"])], ids=["in-process", "nccl-threads"])
def test_randomised_host_flow_cases(mock_env, seed, transport):
    """tests/support/hostflow_fuzz.py: 250 random cases per transport -- ranks, sizes with zeros, output layouts, record and
    key shapes, key distributions (duplicates, all equal, sorted, narrow signed range), options, 1-4 exchange parts on tiny
    inputs (MPSORT_PHASES_MIN_RECORDS=1), host-side switches -- each compared byte for byte with the oracle's contract"""
    rc = run_py(mock_env, [os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), str(seed), "250"] + transport)
    out = rc.stdout.decode()
    assert rc.returncode == 0 and "FUZZ OK" in out, out[-3000:]


# ---- stream semantics: deferred, randomly interleaved execution of the stream operations -----------------
@pytest.mark.parametrize("transport", [[], ["nccl"]], ids=["in-process", "nccl-threads"])
@pytest.mark.parametrize("async_seed", ["1", "2"])
def test_randomised_cases_with_deferred_stream_operations(mock_env, transport, async_seed):
    """tests/native/mock_async.cpp (MOCK_ASYNC=<seed>): copies, memsets, events, NCCL calls and kernel launches only run
    when a synchronisation, an event wait or the seeded scheduler makes them, so a consumer that does not wait for its
    producer gets stale bytes. 150 random cases per transport and seed, every switch incl. the candidates' stream juggling."""
    rc = run_py(mock_env, [os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), "4242", "150"] + transport, MOCK_ASYNC=async_seed)
    out = rc.stdout.decode()
    assert rc.returncode == 0 and "FUZZ OK" in out, out[-3000:]


def test_gpu_suite_host_flow_with_deferred_stream_operations(mock_env):
    """the GPU test files again (see test_gpu_suite_host_flow_on_the_mock), stream operations deferred"""
    args = ["-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "tests/test_python_api.py",
            "tests/test_gpu_parity.py", "tests/test_zz_callback_api.py"]
    for b in BIG:
        args += ["--deselect", b]
    rc = run_py(mock_env, args, timeout=1500, MOCK_ASYNC="3")
    out = rc.stdout.decode()
    m = re.search(r"(\d+) passed", out)
    assert rc.returncode == 0 and m and int(m.group(1)) >= 150 and "failed" not in out, out[-4000:]


MUTANTS = [
    # the bug this model found in round 1: the copy stream of the own slice was only synchronised when p == 8
    ("mpsort_comm.c", "for (k = 0; k < 8; k++) if ((used >> k) & 1u) CUDA_OK(c, cudaStreamWaitEvent(c->p2p.ce_stream[k], gate, 0));",
     "for (k = 0; k < 8 && k < p; k++) if ((used >> k) & 1u) CUDA_OK(c, cudaStreamWaitEvent(c->p2p.ce_stream[k], gate, 0));", ("3", "200", "1")),
    # the merge of a part does not wait for its transfer
    ("mpsort_host.c", "if (Q > 1) CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->phase_ev[late ? Q - 1 : q], 0));", "/* mutant */", ("3", "200", "1")),
    # the copies of an exchange do not wait for the send buffer (the event of the first part is never recorded)
    ("mpsort_comm.c", "CUDA_OK(c, cudaEventRecord(now, c->stream));", "/* mutant */", ("3", "200", "1")),
    # the caller's stream does not wait for the merges on the second stream
    ("mpsort_host.c", "CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->phase_ev[MPS_MAX_RANKS], 0));", "/* mutant */", ("3", "200", "1")),
]


@pytest.mark.parametrize("fname,old,new,fuzz", MUTANTS, ids=["own-slice-stream", "merge-before-transfer", "copies-before-send-buffer", "return-before-merge"])
def test_stream_model_catches_seeded_synchronisation_bugs(tmp_path, mock_env, fname, old, new, fuzz):
    """mutation check of the checker: a copy of the host sources with ONE synchronisation removed must fail the
    randomised cases under MOCK_ASYNC (and these four do, with every seed tried)"""
    import shutil
    csrc = str(tmp_path / "csrc")
    shutil.copytree(hostmock.CSRC, csrc, ignore=shutil.ignore_patterns("kernels"))
    path = os.path.join(csrc, fname)
    src = open(path).read()
    assert src.count(old) == 1, "the mutated line moved: update MUTANTS"
    open(path, "w").write(src.replace(old, new))
    saved = hostmock.CSRC, list(hostmock.INCLUDES)
    try:
        hostmock.CSRC = csrc
        hostmock.INCLUDES[1] = "-I" + csrc
        so = hostmock.compile_and_link(str(tmp_path / "libmutant.so"))
    finally:
        hostmock.CSRC = saved[0]
        hostmock.INCLUDES[:] = saved[1]
    rc = run_py(dict(mock_env, MPSORT_LIB=so), [os.path.join(ROOT, "tests", "support", "hostflow_fuzz.py"), fuzz[0], fuzz[1], "nccl"], MOCK_ASYNC=fuzz[2])
    assert rc.returncode != 0 and b"FUZZ OK" not in rc.stdout, "the mutant went unnoticed"


def test_mock_build_is_not_a_cpu_path_of_the_product(mock_env):
    """pointing MPSORT_LIB at the mock build is not enough: the Python package refuses it without the tests' second switch"""
    env = {k: v for k, v in mock_env.items() if k != "MPSORT_ALLOW_MOCK_DEVICE"}
    rc = run_py(env, "import sys; sys.path.insert(0, %r); import mpsort" % os.path.join(ROOT, "mp-sort_b200"))
    assert rc.returncode != 0 and b"no CPU fallback" in rc.stdout, rc.stdout.decode()[-2000:]
