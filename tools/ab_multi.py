"""A/B of run-time switches inside ONE launch (an 8-GPU launch costs 8x its wall time: start-up, NCCL
initialisation and buffer allocation are paid once for all variants).
   torchrun ... tools/ab_multi.py --gpus N [--log2n 28] [--steps 5] SPEC [SPEC ...]
SPEC = workload:ENV=VAL,ENV=VAL   ("workload:-" = the defaults). Only switches that are read on every call
take effect (MPSORT_EXCHANGE_PHASES, MPSORT_PEER_SPLITTER, MPSORT_FUSED_PACK, MPSORT_PACK_PIPELINE, MPSORT_NO_HYBRID5,
MPSORT_NO_MERGE ...), not those fixed when the communicator is made (MPSORT_NO_P2P, MPSORT_P2P_CE)."""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--e2e", action="store_true", help="also time the host-buffer path of every variant")
ap.add_argument("specs", nargs="+")
args = ap.parse_args()
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import numpy as np  # noqa: E402
import mpsort  # noqa: E402
from mpsort import _capi as C  # noqa: E402
lib = C.lib
comm = mpsort.Comm.from_env()
assert comm.size == args.gpus, (comm.size, args.gpus)
for spec_ in args.specs:
    workload, _, envs = spec_.partition(":")
    pairs = [kv.split("=", 1) for kv in envs.split(",") if kv and kv != "-"]
    for k, v in pairs:
        os.environ[k] = v
    comm.barrier()
    r, din, dout, hash_out, fl = bench.measure(comm, lib, C, workload, args.log2n, args.steps, 2)
    e2e = None
    if args.e2e:
        e2e = bench.measure_e2e(comm, lib, C, mpsort, np, workload, args.log2n, 2, din, hash_out, fl)
    lib.mpsort_util_dev_free(comm.device, din)
    lib.mpsort_util_dev_free(comm.device, dout)
    for k, _ in pairs:
        del os.environ[k]
    if comm.rank == 0:
        print("== %s  %s" % (workload, envs))
        print("   ms/step %.3f  value %.2f Grec/s  %s" % (r["ms_per_step"], r["value"] / 1e9, r["transport"]))
        print("   phases", [(k, round(v, 2)) for k, v in r["phases_ms"] if v > 0.1])
        print("   exch", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["exchange"].items()})
        print("   kern", {k: round(v["ms_per_step"], 2) for k, v in r["kernels"].items()})
        if e2e:
            print("   e2e %.1f ms  %.2f Grec/s  %s" % (e2e["ms_per_step"], e2e["value"] / 1e9, [(k, round(v, 1)) for k, v in e2e["phases_ms"]]))
        sys.stdout.flush()
comm.barrier()
comm.destroy()
