#!/usr/bin/env python
"""bench.py -- the distributed histogram sort of mpsort-b200 on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N = 1: this process)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
  python bench.py --impl reference ...    the reference's own CPU path (oracle/_ref:
                                          unmodified MP-sort sources under the MPI shim)

A "step" is one complete sort of a batch of synthetic records (BASELINE.json):
  N = 1   configs[1]: 2^28 uniform u64-keyed 16-byte records on one GPU
  N > 1   configs[2] shape: 2^28 records per GPU (weak scaling), full all-to-all
`value` times K steps with the input already resident in HBM; `e2e` times the same
sort through the public Python API with pinned HOST buffers (H2D + sort + D2H).
Inputs (4 GiB per GPU) are far larger than L2 (126 MB), so no L2 flush is needed.

Before anything is timed, `parity_preflight` sorts small inputs of all three workloads
(uneven sizes, one empty rank, uneven output sizes, exchange in two parts) through the SAME
communicator and transport and compares every rank's output byte for byte with the oracle
(oracle/mpsort_oracle.py:numpy_sort, the reference's output contract): a mismatch ends the
run with a non-zero exit code. Every timed workload is then verified at full size by
properties: order, tie order, rank boundaries, and an order-independent 64-bit multiset hash
of whole records taken on the device before and after.

After the headline workload, BASELINE.json configs[3] and [4] (`mostly_sorted16`,
`particles48`) run for a few steps each and are reported under `workloads`.
torch is used only as the process launcher for N > 1; workers never import it.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))

SEED = 0x5EED0001
KINDS = {"uniform16": (0, 16, 0), "mostly_sorted16": (1, 16, 0), "particles48": (2, 48, 1)}
MASK64 = (1 << 64) - 1
TUNING_BITS = (1 << 1) | (1 << 3) | (1 << 4) | (1 << 6)     # the reference's four; MPSORT_VERIFY_CHECKSUM stays


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class Clocks(object):
    """nvidia-smi sampling: started early (it takes ~0.2 s to come up), sampled every
    20 ms with timestamps; only samples inside [mark_start, mark_stop] are used."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(t[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(t[1]), float(t[2]), [nm for k, nm in enumerate(names) if t[5 + k].lower().startswith("active")]))
            except ValueError:
                continue
        self.f.close()
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        use = inside if inside else rows[-3:]
        if use:
            sm = sorted(r[1] for r in use)
            reasons = set()
            for r in use:
                reasons.update(r[3])
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in use), reasons=sorted(reasons),
                       samples=len(use), samples_in_timed_region=len(inside))
        return out


def workload_name(workload, log2n):
    """the `config.workload` string both arms report"""
    _, E, signed = KINDS[workload]
    return ("%s: 2^%d %d-byte records per GPU, %s key, device-resident in/out of place"
            % (workload, log2n, E, "i64" if signed else "u64"))


def make_config(workload, log2n, n_gpus):
    """`config` of both arms: the same dict for the same workload (arm-specific facts -- the
    transport, the CPU sample -- are reported beside it, not inside)"""
    _, E, _ = KINDS[workload]
    n = 1 << log2n
    return {"workload": workload_name(workload, log2n), "records_per_gpu": n, "elsize": E,
            "l2": "inputs (%.1f GiB per GPU) larger than L2; no flush" % (n * E / 2.0**30),
            "baseline_config": "configs[1]" if n_gpus == 1 else "configs[2] shape at %d GPUs" % n_gpus}


# ----------------------------------------------------------------------------------------------
# the reference arm

def run_reference(args, workload, rank):
    """The reference arm: unmodified MP-sort (oracle/_ref/bench16 = bench-mpi's sibling for struct
    records -> mpsort_mpi_newarray) on the host cores, one MPI-shim rank per core.
    N = 1: the SAME workload as the GPU arm, all 2^log2n records per sort, mean over the sorts
    that fit --ref-budget-s seconds (after one warm-up sort). N > 1: the GPU arm's workload is
    N x 2^log2n records, more than the reference arm is given time for; each step sorts a bounded
    sample (--ref-records, default 2^24) -- records/s of a merge sort falls slowly with n
    (~log n), so the sample flatters the CPU."""
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mpsort_oracle as O
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs the reference sources at build time)"}))
        return 0
    kind, elsize, _ = KINDS[workload]
    cores = os.cpu_count() or 1
    np_ranks = max(1, min(cores, 64))
    full = args.gpus == 1 and args.ref_records is None
    total = (1 << args.log2n) if full else int(args.ref_records or (1 << 24))
    reps = args.steps + args.warmup
    warm = 1 if reps > 1 else 0
    t0 = time.time()
    r = O.run_bench16(np_ranks, 0, elsize=elsize, kind=kind, reps=reps, timeout=3000, warmups=warm,
                      budget_s=args.ref_budget_s, total=total)
    wall = time.time() - t0
    value = r["mean_records_per_second"]
    sample = ("%d records per sort (%d MPI-shim ranks = host threads), %s, unmodified reference mpsort_mpi_newarray, "
              "gcc -O2; mean of %d sorts after %d warm-up sort(s) (%d of the %d requested repetitions fit %.0f s)"
              % (total, np_ranks, workload, r["reps_timed"], warm, r["reps_done"], reps, args.ref_budget_s))
    config = make_config(workload, args.log2n, args.gpus)
    if not full:
        config["records_per_step"] = total
        config["note"] = ("each step sorts a bounded sample of the GPU arm's workload (same generator, same record "
                          "layout) with the reference's CPU implementation")
    line = {
        "impl": "reference", "metric": "sorted records/s", "value": value, "unit": "records/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["mean_seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "gb_per_s": value * elsize / 1e9,
        "config": config,
        "same_workload_as_gpu_arm": bool(full),
        "sorts_timed": r["reps_timed"], "sorts_run": r["reps_done"], "seconds_per_sort": r["totals"],
        "best_records_per_second": r["records_per_second"],
        "cpu_baseline": {"value": value, "unit": "records/s", "cores": np_ranks, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases_s": r["phases"], "wall_s": wall, "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# parity pre-flight

def parity_preflight(comm, lib, C, mpsort, np, log2n=22):
    """Bit-exact check of the production path before anything is timed: every workload, small,
    through this communicator (NCCL ranks + IPC-mapped DMA exchange at N > 1), against the oracle.
    Collective; returns the report dict or raises SystemExit on every rank."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mpsort_oracle as O
    p, me, dev = comm.size, comm.rank, comm.device
    saved = os.environ.get("MPSORT_PHASES_MIN_RECORDS")
    os.environ["MPSORT_PHASES_MIN_RECORDS"] = "1"           # the two-part exchange + overlapped merge also for small inputs
    cases, ok_all = [], True
    t0 = time.time()
    try:
        # uniform16 at >= 2^22 records per rank takes the hybrid sort (four passes + run fix-up), as the headline does
        small = min(300000, (1 << log2n) // 8 + 3000)
        for workload, base in (("uniform16", (1 << log2n) + 4321), ("mostly_sorted16", small), ("particles48", small * 2 // 3)):
            kind, E, signed = KINDS[workload]
            sizes = [base + 1237 * r for r in range(p)]
            if p > 1:
                sizes[1] = 0                                   # one rank without input
            total = sum(sizes)
            outsizes = [total // p + 1000 * (2 * r - (p - 1)) for r in range(p)]
            outsizes[-1] += total - sum(outsizes)
            assert min(outsizes) >= 0 and sum(outsizes) == total
            recs = [O.generate(sizes[r], E, kind, SEED + 7, r, p) for r in range(p)]
            odesc = O.Desc(0, 8, 1, signed, 0)
            exp = O.numpy_sort(recs, odesc, outsizes)[me]
            desc = C.RadixDesc(0, 8, 1, signed, 0)
            n, on = sizes[me], outsizes[me]
            din = lib.mpsort_util_dev_malloc(dev, n * E)
            dout = lib.mpsort_util_dev_malloc(dev, on * E)
            # the device generator must agree with the oracle's restatement of it
            lib.mpsort_util_generate_as(comm.handle, din, n, E, kind, SEED + 7, me, p)
            got_in = np.zeros((n, E), np.uint8)
            lib.mpsort_util_memcpy(dev, got_in.ctypes.data, din, n * E)
            gen_ok = bool(np.array_equal(got_in, recs[me]))
            # device-resident, out of place (what `value` times)
            lib.mpsort_mpi_unset_options(TUNING_BITS)
            lib.mpsort_mpi_newarray_desc_impl(din, n, dout, on, E, C.byref(desc), comm.handle, 0, b"bench.py:preflight")
            st = C.last_stats(comm.handle, p)
            got = np.zeros((on, E), np.uint8)
            lib.mpsort_util_memcpy(dev, got.ctypes.data, dout, on * E)
            dev_ok = bool(np.array_equal(got, exp))
            lib.mpsort_util_dev_free(dev, din)
            lib.mpsort_util_dev_free(dev, dout)
            # host buffers through the public Python API (what `e2e` times)
            dt = np.dtype([("key", "i8" if signed else "u8"), ("rest", "u1", E - 8)])
            hin = recs[me].copy().view(dt).reshape(-1)
            hout = np.zeros(on, dtype=dt)
            mpsort.sort(hin, "key", out=hout, comm=comm)
            api_ok = bool(np.array_equal(hout.view(np.uint8).reshape(on, E), exp))
            case = {"workload": workload, "records": total, "sizes": sizes, "outsizes": outsizes,
                    "generator_equals_oracle": gen_ok, "device_resident_equals_oracle": dev_ok,
                    "host_api_equals_oracle": api_ok,
                    "exchange_parts": st["exchange_phases"], "p2p_exchange": st["p2p_exchange"],
                    "merge_tiles": st["second_sort_merge_tiles"], "hybrid": st["hybrid"], "record_mode": st["record_mode"],
                    "own_slices_merged_in_place": st.get("own_slices_in_place", 0)}
            oks = comm.allgather((gen_ok, dev_ok, api_ok, st["second_sort_merge_tiles"]))
            case["all_ranks_ok"] = all(a and b and c for a, b, c, _ in oks)
            case["merge_tiles_all_ranks"] = [o[3] for o in oks]
            ok_all = ok_all and case["all_ranks_ok"]
            cases.append(case)
    finally:
        if saved is None:
            del os.environ["MPSORT_PHASES_MIN_RECORDS"]
        else:
            os.environ["MPSORT_PHASES_MIN_RECORDS"] = saved
    report = {"ok": ok_all, "ranks": p, "checker": "oracle numpy_sort (the reference's output contract), byte for byte on every rank",
              "transport": ("none (1 GPU)" if p == 1 else
                            ("peer copies over NVLink (CUDA IPC mapped receive buffers, DMA engines)" if cases[0]["p2p_exchange"]
                             else "NCCL grouped send/recv over NVLink")),
              "seconds": round(time.time() - t0, 2), "cases": cases}
    if not ok_all:
        if me == 0:
            print(json.dumps({"parity_preflight": report}))
        raise SystemExit("bench.py: parity pre-flight FAILED on rank %d of %d: %r" % (me, p, [
            (c["workload"], c["generator_equals_oracle"], c["device_resident_equals_oracle"], c["host_api_equals_oracle"]) for c in cases]))
    return report


# ----------------------------------------------------------------------------------------------
# one workload, device-resident

def combine_hashes(pairs):
    s = x = 0
    for a, b in pairs:
        s = (s + a) & MASK64
        x ^= b
    return s, x


def measure(comm, lib, C, workload, log2n, K, W, clocks=None):
    """W warm-up sorts, K timed sorts of one workload with the input resident in HBM; verified.
    Returns (dict, din, dout, hash_out, firstlast) -- the buffers stay allocated for the e2e leg."""
    dev = comm.device
    kind, E, signed = KINDS[workload]
    n = 1 << log2n
    desc = C.RadixDesc(0, 8, 1, signed, 0)
    din = lib.mpsort_util_dev_malloc(dev, n * E)
    dout = lib.mpsort_util_dev_malloc(dev, n * E)
    lib.mpsort_util_generate(comm.handle, din, n, E, kind, SEED)
    hash_in = C.multiset_hash(comm.handle, din, n, E)
    lib.mpsort_mpi_unset_options(TUNING_BITS)

    def step():
        lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, C.byref(desc), comm.handle, 0, b"bench.py")

    def maxall(x):
        return max(comm.allgather(float(x)))

    for _ in range(W):
        step()
    e0 = lib.mpsort_util_event_create(comm.handle)
    e1 = lib.mpsort_util_event_create(comm.handle)
    lib.mpsort_util_kernel_timing(comm.handle, 1)
    lib.mpsort_util_launch_count(1)
    comm.barrier()
    lib.mpsort_util_stream_sync(comm.handle)
    if clocks:
        clocks.mark_start()
    lib.mpsort_util_event_record(comm.handle, e0)
    for _ in range(K):
        step()
    lib.mpsort_util_event_record(comm.handle, e1)
    lib.mpsort_util_stream_sync(comm.handle)
    if clocks:
        clocks.mark_stop()
    comm.barrier()
    ms_total = maxall(lib.mpsort_util_event_elapsed_ms(comm.handle, e0, e1))
    launches = int(lib.mpsort_util_launch_count(0))
    ktimes = C.kernel_times(comm.handle)
    lib.mpsort_util_kernel_timing(comm.handle, 0)
    phases = C.last_run()
    stats = C.last_stats(comm.handle, comm.size)
    lib.mpsort_util_event_destroy(e0)
    lib.mpsort_util_event_destroy(e1)

    # ---- the result is checked, every run: order + tie order + the multiset of records preserved
    fl = (ctypes.c_uint64 * 2)()
    bad = lib.mpsort_util_check_sorted(comm.handle, dout, n, E, C.byref(desc), 1, 8, fl)
    hash_out = C.multiset_hash(comm.handle, dout, n, E)
    alls = comm.allgather((hash_in, hash_out, int(bad), int(fl[0]), int(fl[1])))
    ok = combine_hashes([a[0] for a in alls]) == combine_hashes([a[1] for a in alls]) and all(a[2] == 0 for a in alls)
    flip = (1 << 63) if signed else 0
    for a, b in zip(alls[:-1], alls[1:]):
        ok = ok and (a[4] ^ flip) <= (b[3] ^ flip)       # last key of rank r <= first key of rank r+1
    if not ok:
        raise SystemExit("bench.py: the sorted output of %s failed verification: %r" % (workload, alls))

    ms_step = ms_total / K
    total_records = n * comm.size
    value = total_records / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (the onesweep pass that ran), measured live above
    peak, peak_src = peaks()
    # record mode (16-byte records carried through the passes) moves 32 B per record and
    # pass, index mode (u64 key + u32 index) 24 B; whichever ran is the dominant kernel
    cand = [("onesweep_pass_rec16", "onesweep_rec16_kernel", 32.0), ("onesweep_pass", "onesweep_kernel", 24.0)]
    cand = [(ktimes.get(c[0], (0.0, 0)), c) for c in cand]
    (ms_sweep, n_sweep), (kclass, kname, bytes_per_item) = max(cand, key=lambda x: x[0][0])
    roofline = None
    if n_sweep:
        alg_bytes = bytes_per_item * n   # per launch: every record (or key+index pair) read once, written once
        ach = alg_bytes / (ms_sweep / n_sweep * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "launches_timed": n_sweep,
                    "avg_launch_ms": ms_sweep / n_sweep,
                    "share_of_step": ms_sweep / (K * ms_step)}
        prof = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
        if os.path.exists(prof) and log2n == 28:
            try:
                with open(prof) as f:
                    t = json.load(f).get(kname, {})
                roofline["traffic"] = t.get("dram_bytes_per_launch")
                roofline["traffic_source"] = t.get("source", "ncu --set full capture of the same kernel and size, profiles/")
            except Exception:
                pass
    kernels = {}
    algo = {"onesweep_pass": 24.0 * n, "onesweep_pass_rec16": 32.0 * n, "gather_records": (2.0 * E + 4) * n,
            "merge_runs": 2.0 * E * n, "extract_hist": float(E) * n, "hybrid_fixup": float(E) * n}
    for name, (ms, cnt) in ktimes.items():
        if cnt:
            kernels[name] = {"ms_per_step": ms / K, "launches_per_step": cnt / K}
            if name in algo:
                # one launch of a pass / gather moves the whole array; the merge, the histogram pass and the
                # fix-up are several launches (sample sort, bounds, tiles, one set per exchange part) per array
                per_array_ms = ms / cnt if name in ("onesweep_pass", "onesweep_pass_rec16", "gather_records") else ms / K
                kernels[name]["achieved_gbs"] = algo[name] / (per_array_ms * 1e-3) / 1e9
                kernels[name]["frac_of_peak"] = kernels[name]["achieved_gbs"] / peak
    if stats.get("record_mode"):
        local_sort_bytes = (E + 32.0 * stats["first_sort_passes"] + (E if stats.get("hybrid") else 0)) * n
    else:
        local_sort_bytes = (3.0 * E + 12 + 24.0 * stats["first_sort_passes"]) * n
    ex_ms = kernels.get("exchange", {}).get("ms_per_step")
    first_ms = sum(v for k, v in phases if k == "FirstSort") * 1e3
    out = {
        "value": value, "ms_per_step": ms_step, "gb_per_s": value * E / 1e9, "steps": K, "warmup": W,
        "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
        "transport": "none (1 GPU)" if comm.size == 1 else
                     (("peer copies over NVLink (CUDA IPC mapped receive buffers, DMA engines)" if stats.get("p2p_exchange")
                       else "NCCL grouped send/recv over NVLink") + ", %d part(s)" % max(1, stats.get("exchange_phases", 1))),
        "local_sort": {"algorithmic_bytes": local_sort_bytes, "passes": stats["first_sort_passes"],
                       "record_mode": bool(stats.get("record_mode")), "hybrid": bool(stats.get("hybrid")),
                       "rebased": bool(stats.get("rebased")), "merge_tiles": stats.get("second_sort_merge_tiles"),
                       "second_sort_passes": stats["second_sort_passes"], "ms": first_ms,
                       "frac_of_peak": (local_sort_bytes / (first_ms * 1e-3) / 1e9 / peak) if first_ms > 0 else None},
        "phases_ms": [[k, v * 1e3] for k, v in phases],
        "exchange": {"bytes_sent_remote_rank0": stats["bytes_sent_remote"], "ms": ex_ms,
                     "gb_per_s_per_gpu": (stats["bytes_sent_remote"] / (ex_ms * 1e-3) / 1e9) if ex_ms and stats["bytes_sent_remote"] else None,
                     "frac_of_nvlink_900": (stats["bytes_sent_remote"] / (ex_ms * 1e-3) / 1e9 / 900.0) if ex_ms and stats["bytes_sent_remote"] else None,
                     "splitter_rounds": stats["splitter_rounds"],
                     "own_slices_merged_in_place": stats.get("own_slices_in_place", 0)},
        "verified": "order + tie order + rank boundaries + 64-bit multiset hash (sum and xor of mix64-folded records) "
                    "of every rank's input and output",
    }
    return out, din, dout, hash_out, (int(fl[0]), int(fl[1]))


def measure_e2e(comm, lib, C, mpsort, np, workload, log2n, K, din, hash_out, firstlast):
    """the same sort through the public Python API with pinned HOST buffers: H2D + sort + D2H
    inside the timed region, every step"""
    dev = comm.device
    kind, E, signed = KINDS[workload]
    n = 1 << log2n
    flip = (1 << 63) if signed else 0
    dt = np.dtype([("key", "i8" if signed else "u8"), ("rest", "u1", E - 8)])
    hin_p = lib.mpsort_util_host_malloc_pinned(n * E)
    hout_p = lib.mpsort_util_host_malloc_pinned(n * E) if hin_p else None
    pinned_here = bool(hin_p and hout_p)
    if pinned_here:
        hin = np.ctypeslib.as_array(ctypes.cast(hin_p, ctypes.POINTER(ctypes.c_uint8)), shape=(n * E,)).view(dt)
        hout = np.ctypeslib.as_array(ctypes.cast(hout_p, ctypes.POINTER(ctypes.c_uint8)), shape=(n * E,)).view(dt)
    else:
        # the box refused to pin 2 x n*E bytes: pageable numpy arrays (slower copies, same call)
        lib.mpsort_util_host_free_pinned(hin_p)
        try:
            hin = np.empty(n, dtype=dt)
            hout = np.empty(n, dtype=dt)
            hin_p, hout_p = hin.ctypes.data, hout.ctypes.data
        except MemoryError:
            hin = hout = None
    pinned = min(comm.allgather(pinned_here))
    have_host = min(comm.allgather(hin is not None))
    if not have_host:
        return {"value": None, "unit": "records/s", "error": "no host memory for 2 x %d bytes per rank" % (n * E)}
    lib.mpsort_util_memcpy(dev, hin_p, din, n * E)
    mpsort.sort(hin, "key", out=hout, comm=comm)                    # warm-up
    e0 = lib.mpsort_util_event_create(comm.handle)
    e1 = lib.mpsort_util_event_create(comm.handle)
    comm.barrier()
    t0 = time.perf_counter()
    lib.mpsort_util_event_record(comm.handle, e0)
    ke = max(1, min(K, 3))
    for _ in range(ke):
        mpsort.sort(hin, "key", out=hout, comm=comm)
    lib.mpsort_util_event_record(comm.handle, e1)
    lib.mpsort_util_stream_sync(comm.handle)
    wall = time.perf_counter() - t0
    comm.barrier()
    ms_e2e = max(comm.allgather(float(max(lib.mpsort_util_event_elapsed_ms(comm.handle, e0, e1), wall * 1e3)))) / ke
    phases = C.last_run()
    # the host result must equal the device result: same multiset of records on every rank, same first key
    same = C.multiset_hash(comm.handle, hout_p, n, E) == hash_out
    first_ok = (int(hout.view(np.uint8)[:8].view("<u8")[0]) ^ flip) == firstlast[0]
    if not (same and first_ok):
        raise SystemExit("bench.py: host-buffer result differs from the device-resident result")
    total_records = n * comm.size
    e2e = {"value": total_records / (ms_e2e * 1e-3), "unit": "records/s",
           "h2d_bytes_per_step": n * E * comm.size, "d2h_bytes_per_step": n * E * comm.size,
           "bytes_per_gpu_each_way": n * E, "ms_per_step": ms_e2e,
           "gb_per_s_per_gpu_each_way_if_copies_only": 2.0 * n * E / (ms_e2e * 1e-3) / 1e9 / 2.0,
           "api": "mpsort.sort(numpy %s host array, 'key', out=host array, comm)" % ("pinned" if pinned else "pageable"),
           "steps": ke, "phases_ms": [[k, v * 1e3] for k, v in phases if v > 2e-4],
           "numa": getattr(comm, "numa", None)}
    del hin, hout
    if pinned_here:
        lib.mpsort_util_host_free_pinned(hin_p)
        lib.mpsort_util_host_free_pinned(hout_p)
    lib.mpsort_util_event_destroy(e0)
    lib.mpsort_util_event_destroy(e1)
    return e2e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uniform16", choices=sorted(KINDS))
    ap.add_argument("--log2n", type=int, default=28, help="records per GPU = 2^log2n")
    ap.add_argument("--ref-records", type=float, default=None,
                    help="total records of one sort of the CPU reference (default: the full 2^log2n at --gpus 1, 2^24 otherwise)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="reference arm: stop repeating sorts after this many seconds of sorting")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-preflight", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true",
                    help="skip BASELINE.json configs[3], [4] (mostly_sorted16, particles48) after the headline workload")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--preflight-log2n", type=int, default=22, help="records per rank of the pre-flight's uniform16 case")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, args.workload, rank)
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch ourselves under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))

    # NCCL prints its version banner to stdout when NCCL_DEBUG is set: keep stdout for the JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import numpy as np
    import mpsort                       # raises if libmpsort-b200.so / the binding are missing
    from mpsort import _capi as C
    lib = C.lib
    if lib.mpsort_util_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; mpsort-b200 has no CPU fallback")

    comm = mpsort.Comm.from_env()
    clocks = Clocks(comm.device) if comm.rank == 0 else None
    K, W = args.steps, args.warmup

    preflight = None
    if not args.no_preflight:
        preflight = parity_preflight(comm, lib, C, mpsort, np, args.preflight_log2n)

    # ---- the headline workload: device-resident `value`, then `e2e` with host buffers
    head, din, dout, hash_out, fl = measure(comm, lib, C, args.workload, args.log2n, K, W, clocks)
    clk = clocks.stop() if clocks else None
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(comm, lib, C, mpsort, np, args.workload, args.log2n, K, din, hash_out, fl)
    lib.mpsort_util_dev_free(comm.device, din)
    lib.mpsort_util_dev_free(comm.device, dout)

    # ---- BASELINE.json configs[3] and [4]: a few steps each, same verification
    extra = {}
    if not args.no_extra_workloads:
        for w in ("mostly_sorted16", "particles48"):
            if w == args.workload:
                continue
            r, a, b, _, _ = measure(comm, lib, C, w, args.log2n, max(1, args.extra_steps), 2)
            lib.mpsort_util_dev_free(comm.device, a)
            lib.mpsort_util_dev_free(comm.device, b)
            r["config"] = make_config(w, args.log2n, comm.size)
            extra[w] = r

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference itself; one full-size sort by default
    cpu = None
    if comm.rank == 0 and comm.size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mpsort_oracle as O
        kind, E, _ = KINDS[args.workload]
        if O.have_ref():
            cores = max(1, min(os.cpu_count() or 1, 64))
            total = int(args.ref_records) if args.ref_records else (1 << args.log2n)
            r = O.run_bench16(cores, 0, elsize=E, kind=kind, reps=1, timeout=1800, total=total)
            cpu = {"value": r["records_per_second"], "unit": "records/s", "cores": cores, "kind": "reference",
                   "sample": "one sort of %d records (%s; %d MPI-shim ranks = host threads) through the unmodified reference "
                             "mpsort_mpi_newarray (oracle/_ref/bench16, gcc -O2)%s"
                             % (total, args.workload, cores, ": the GPU arm's full workload" if total == (1 << args.log2n) else ""),
                   "seconds": r["best_seconds"], "phases_s": r["phases"]}
        else:
            cpu = {"value": None, "unit": "records/s", "cores": 0, "kind": "reference",
                   "sample": "oracle/_ref missing on this box"}

    if comm.rank == 0:
        line = {
            "metric": "sorted records/s", "value": head["value"], "unit": "records/s", "n_gpus": comm.size,
            "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "gb_per_s": head["gb_per_s"],
            "config": make_config(args.workload, args.log2n, comm.size),
            "transport": head["transport"],
            "clocks": clk, "e2e": e2e, "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "cpu_baseline": cpu,
            "parity_preflight": preflight,
            "kernels": head["kernels"], "local_sort": head["local_sort"], "phases_ms": head["phases_ms"],
            "exchange": head["exchange"], "verified": head["verified"],
            "workloads": extra,
        }
        print(json.dumps(line))
    comm.barrier()
    comm.destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
