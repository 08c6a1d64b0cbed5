"""Dry run of bench.py's GPU arm WITHOUT a GPU (test support for tests/test_bench_contract.py): the
library is replaced by a stand-in that returns plausible numbers, so that the control flow and the
JSON line of bench.py are exercised on a CPU-only box. Prints the JSON line. `--pinfail` makes the
pinned allocation fail (pageable fallback of the e2e leg)."""
import sys, types, ctypes, json, io, contextlib
sys.argv = ["bench.py", "--log2n", "10", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-preflight"] + sys.argv[1:]
PIN_FAIL = "--pinfail" in sys.argv
if PIN_FAIL: sys.argv.remove("--pinfail")
SIZE = 1
if "--size" in sys.argv:
    i = sys.argv.index("--size"); SIZE = int(sys.argv[i + 1]); del sys.argv[i:i + 2]
    sys.argv += ["--gpus", str(SIZE)]
    import os as _os
    _os.environ["WORLD_SIZE"] = str(SIZE); _os.environ["RANK"] = "0"
keep = []
class Lib:
    def __getattr__(self, name):
        def f(*a):
            if name == "mpsort_util_device_count": return 1
            if name in ("mpsort_util_dev_malloc",): return 0x1000
            if name == "mpsort_util_host_malloc_pinned":
                if PIN_FAIL: return None
                b = ctypes.create_string_buffer(a[0]); keep.append(b); return ctypes.addressof(b)
            if name == "mpsort_util_checksum": return 12345
            if name == "mpsort_util_check_sorted": return 0
            if name == "mpsort_util_event_elapsed_ms": return 20.0
            if name == "mpsort_util_launch_count": return 68
            if name in ("mpsort_util_event_create",): return 1
            return None
        return f
capi = types.ModuleType("mpsort._capi")
capi.lib = Lib()
capi.byref = ctypes.byref
class RadixDesc(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_size_t), ("width", ctypes.c_uint32), ("nwords", ctypes.c_uint32), ("is_signed", ctypes.c_int32), ("reserved", ctypes.c_int32)]
capi.RadixDesc = RadixDesc
capi.kernel_times = lambda h: {"extract_hist": (1.4, 8), "onesweep_pass_rec16": (16.0, 8), "hybrid_fixup": (2.4, 18), "merge_runs": (10.0, 48), "exchange": (11.0, 4)}
capi.last_run = lambda: [("FirstSort", 0.010), ("Exchange", 0.005)]
capi.last_stats = lambda h, s: {"first_sort_passes": 4, "second_sort_passes": 0, "record_mode": 1, "second_sort_merge_tiles": 10, "bytes_sent_remote": 1000, "p2p_exchange": 1, "exchange_phases": 2, "hybrid": 1, "rebased": 0, "splitter_rounds": 8}
capi.multiset_hash = lambda h, base, n, E: (777, 888)
mp = types.ModuleType("mpsort")
class Comm:
    rank, size, device, handle = 0, SIZE, 0, 1
    @classmethod
    def from_env(cls): return cls()
    def allgather(self, x): return [x] * SIZE
    def barrier(self): pass
    def destroy(self): pass
mp.Comm = Comm
mp.sort = lambda src, key, out=None, comm=None: out.view("u1").fill(0)
mp._capi = capi
sys.modules["mpsort"] = mp; sys.modules["mpsort._capi"] = capi
import importlib.util
import os
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
b.Clocks = lambda dev: types.SimpleNamespace(mark_start=lambda: None, mark_stop=lambda: None, stop=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    rc = b.main()
print(buf.getvalue().strip().splitlines()[-1])
