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


def run_reference(args, workload, rank):
    """the reference arm: unmodified MP-sort (oracle/_ref/bench16 = bench-mpi's sibling
    for struct records) on the host cores, one MPI-shim rank per core, on a bounded
    sample of the same workload"""
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
    total = int(args.ref_records)
    per_rank = max(1, total // np_ranks)
    reps = args.steps + args.warmup
    t0 = time.time()
    r = O.run_bench16(np_ranks, per_rank, elsize=elsize, kind=kind, reps=reps, timeout=3000)
    wall = time.time() - t0
    value = r["records_per_second"]
    sample = ("%d records (%d per rank x %d MPI-shim ranks), %s, unmodified reference mpsort_mpi_newarray, gcc -O2; "
              "best of %d runs" % (per_rank * np_ranks, per_rank, np_ranks, workload, reps))
    line = {
        "impl": "reference", "metric": "sorted records/s", "value": value, "unit": "records/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["best_seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "gb_per_s": value * elsize / 1e9,
        "config": {"workload": workload_name(workload, args.log2n), "records_per_gpu": 1 << args.log2n, "elsize": elsize,
                   "baseline_config": "configs[1]" if args.gpus == 1 else "configs[2] shape at %d GPUs" % args.gpus,
                   "records_per_step": per_rank * np_ranks,
                   "note": "each step sorts a bounded sample of the GPU arm's workload (same generator, same record "
                           "layout) with the reference's CPU implementation; no GPU"},
        "cpu_baseline": {"value": value, "unit": "records/s", "cores": np_ranks, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases_s": r["phases"], "wall_s": wall, "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uniform16", choices=sorted(KINDS))
    ap.add_argument("--log2n", type=int, default=28, help="records per GPU = 2^log2n")
    ap.add_argument("--ref-records", type=float, default=float(1 << 24),
                    help="total records of the CPU reference sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    dev = comm.device
    clocks = Clocks(dev) if comm.rank == 0 else None
    kind, E, signed = KINDS[args.workload]
    n = 1 << args.log2n
    desc = C.RadixDesc(0, 8, 1, signed, 0)
    K, W = args.steps, args.warmup

    din = lib.mpsort_util_dev_malloc(dev, n * E)
    dout = lib.mpsort_util_dev_malloc(dev, n * E)
    lib.mpsort_util_generate(comm.handle, din, n, E, kind, SEED)
    sum_in = lib.mpsort_util_checksum(comm.handle, din, n * E)
    lib.mpsort_mpi_unset_options(-1)

    def step():
        lib.mpsort_mpi_newarray_desc_impl(din, n, dout, n, E, C.byref(desc), comm.handle, 0, b"bench.py")

    def maxall(x):
        return max(comm.allgather(float(x)))

    # ---- device-resident: `value`
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
    clk = clocks.stop() if clocks else None
    launches = int(lib.mpsort_util_launch_count(0))
    ktimes = C.kernel_times(comm.handle)
    lib.mpsort_util_kernel_timing(comm.handle, 0)
    phases = C.last_run()
    stats = C.last_stats(comm.handle, comm.size)

    # ---- the result is checked, every run: order + tie order + bytes preserved
    fl = (ctypes.c_uint64 * 2)()
    bad = lib.mpsort_util_check_sorted(comm.handle, dout, n, E, C.byref(desc), 1, 8, fl)
    sum_out = lib.mpsort_util_checksum(comm.handle, dout, n * E)
    sums = comm.allgather((int(sum_in), int(sum_out), int(bad), int(fl[0]), int(fl[1])))
    mask = (1 << 64) - 1
    ok = (sum(s[0] for s in sums) & mask) == (sum(s[1] for s in sums) & mask) and all(s[2] == 0 for s in sums)
    flip = (1 << 63) if signed else 0
    for a, b in zip(sums[:-1], sums[1:]):
        ok = ok and (a[4] ^ flip) <= (b[3] ^ flip)       # last key of rank r <= first key of rank r+1
    if not ok:
        raise SystemExit("bench.py: the sorted output failed verification: %r" % (sums,))

    ms_step = ms_total / K
    total_records = n * comm.size
    value = total_records / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (the onesweep pass), measured live above
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
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    roofline["traffic"] = json.load(f).get(kname, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
    kernels = {}
    passes = stats["first_sort_passes"] + stats["second_sort_passes"]
    algo = {"onesweep_pass": 24.0 * n, "onesweep_pass_rec16": 32.0 * n, "gather_records": (2.0 * E + 4) * n,
            "merge_runs": 2.0 * E * n}
    for name, (ms, cnt) in ktimes.items():
        if cnt:
            kernels[name] = {"ms_per_step": ms / K, "launches_per_step": cnt / K}
            if name in algo:
                # one launch of a pass / gather moves the whole array; the merge is many small
                # launches (sample sort, bounds, tiles, one set per exchange part) per array
                per_array_ms = ms / K if name == "merge_runs" else ms / cnt
                kernels[name]["achieved_gbs"] = algo[name] / (per_array_ms * 1e-3) / 1e9
                kernels[name]["frac_of_peak"] = kernels[name]["achieved_gbs"] / peak
    if stats.get("record_mode"):
        local_sort_bytes = (E + 32.0 * stats["first_sort_passes"]) * n
    else:
        local_sort_bytes = (3.0 * E + 12 + 24.0 * stats["first_sort_passes"]) * n

    # ---- end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
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
    if not args.no_e2e and not have_host:
        e2e = {"value": None, "unit": "records/s", "error": "no host memory for 2 x %d bytes per rank" % (n * E)}
    elif not args.no_e2e:
        lib.mpsort_util_memcpy(dev, hin_p, din, n * E)
        mpsort.sort(hin, "key", out=hout, comm=comm)                    # warm-up
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
        ms_e2e = maxall(max(lib.mpsort_util_event_elapsed_ms(comm.handle, e0, e1), wall * 1e3)) / ke
        # the host result must equal the device result
        lib.mpsort_util_memcpy(dev, din, hout_p, n * E)
        same = lib.mpsort_util_checksum(comm.handle, din, n * E) == sum_out
        first_last_ok = (int(hout.view(np.uint8)[:8].view("<u8")[0]) ^ flip) == int(fl[0])
        if not (same and first_last_ok):
            raise SystemExit("bench.py: host-buffer result differs from the device-resident result")
        e2e = {"value": total_records / (ms_e2e * 1e-3), "unit": "records/s",
               "h2d_bytes_per_step": n * E * comm.size, "d2h_bytes_per_step": n * E * comm.size,
               "bytes_per_gpu_each_way": n * E, "ms_per_step": ms_e2e,
               "api": "mpsort.sort(numpy %s host array, 'key', out=host array, comm)" % ("pinned" if pinned else "pageable"),
               "steps": ke}
        del hin, hout
        if pinned_here:
            lib.mpsort_util_host_free_pinned(hin_p)
            lib.mpsort_util_host_free_pinned(hout_p)

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference itself on a bounded sample
    cpu = None
    if comm.rank == 0 and comm.size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mpsort_oracle as O
        if O.have_ref():
            cores = max(1, min(os.cpu_count() or 1, 64))
            per_rank = max(1, int(args.ref_records) // cores)
            r = O.run_bench16(cores, per_rank, elsize=E, kind=kind, reps=1, timeout=1800)
            cpu = {"value": r["records_per_second"], "unit": "records/s", "cores": cores, "kind": "reference",
                   "sample": "%d records (%d per rank x %d MPI-shim ranks) of %s through the unmodified reference "
                             "mpsort_mpi_newarray (oracle/_ref/bench16, gcc -O2)" % (per_rank * cores, per_rank, cores, args.workload),
                   "phases_s": r["phases"]}
        else:
            cpu = {"value": None, "unit": "records/s", "cores": 0, "kind": "reference",
                   "sample": "oracle/_ref missing on this box"}

    if comm.rank == 0:
        line = {
            "metric": "sorted records/s", "value": value, "unit": "records/s", "n_gpus": comm.size,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "gb_per_s": value * E / 1e9,
            "config": {"workload": workload_name(args.workload, args.log2n),
                       "records_per_gpu": n, "elsize": E, "l2": "inputs (%.1f GiB per GPU) larger than L2; no flush" % (n * E / 2.0**30),
                       "transport": "none (1 GPU)" if comm.size == 1 else
                                    (("peer copies over NVLink (CUDA IPC mapped receive buffers, DMA engines)" if stats.get("p2p_exchange")
                                      else "NCCL grouped send/recv over NVLink") + ", %d part(s)" % max(1, stats.get("exchange_phases", 1))),
                       "baseline_config": "configs[1]" if comm.size == 1 else "configs[2] shape at %d GPUs" % comm.size},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels,
            "local_sort": {"algorithmic_bytes": local_sort_bytes, "passes": stats["first_sort_passes"],
                           "record_mode": bool(stats.get("record_mode")), "merge_tiles": stats.get("second_sort_merge_tiles"),
                           "second_sort_passes": stats["second_sort_passes"],
                           "ms": sum(v for k, v in phases if k == "FirstSort") * 1e3},
            "phases_ms": [[k, v * 1e3] for k, v in phases],
            "exchange": {"bytes_sent_remote_rank0": stats["bytes_sent_remote"],
                         "ms": kernels.get("exchange", {}).get("ms_per_step"),
                         "gb_per_s_per_gpu": (stats["bytes_sent_remote"] / (kernels["exchange"]["ms_per_step"] * 1e-3) / 1e9)
                         if "exchange" in kernels and stats["bytes_sent_remote"] else None},
            "verified": "order + tie order + byte checksum of every rank's output",
        }
        print(json.dumps(line))
    lib.mpsort_util_dev_free(dev, din)
    lib.mpsort_util_dev_free(dev, dout)
    comm.barrier()
    comm.destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
