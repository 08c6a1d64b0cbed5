"""ctypes view of libmpsort-b200.so (the C ABI declared in include/mpsort.h and
include/mpsort_util.h).

The library is the product; this module only declares prototypes. There is no
fallback: if the shared object is missing, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPSORT_LIB") or os.path.join(os.path.dirname(_HERE), "libmpsort-b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "mpsort-b200: %s is missing. Build it with `make -C %s` (or "
        "`python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback."
        % (LIB_PATH, os.path.dirname(_HERE)))

# libmpsort-b200.so links libnccl.so.2 dynamically. A process that ALSO imports PyTorch must
# end up with one NCCL: either import torch first (its bundled, newer NCCL is then shared), or
# point MPSORT_NCCL_LIB at the libnccl.so.2 to preload (e.g. site-packages/nvidia/nccl/lib/).
_nccl = os.environ.get("MPSORT_NCCL_LIB")
if _nccl:
    ctypes.CDLL(_nccl, mode=ctypes.RTLD_GLOBAL)

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

# tests/native/ builds a stand-in for the GPU side (the C host code against a mock device) for CPU-only host-flow
# tests. It is not a CPU path of the product: loading it takes MPSORT_LIB pointing at it AND this second switch.
if hasattr(lib, "mocksync_cudaFree") and os.environ.get("MPSORT_ALLOW_MOCK_DEVICE") != "1":
    raise ImportError("mpsort-b200: %s is the mock-device build of the test suite, not the product library; "
                      "there is no CPU fallback (tests set MPSORT_ALLOW_MOCK_DEVICE=1)" % LIB_PATH)

c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
c_int = ctypes.c_int
c_u64 = ctypes.c_uint64
c_i64 = ctypes.c_int64

MPSORT_DISABLE_SPARSE_ALLTOALLV = 1 << 1
MPSORT_DISABLE_GATHER_SORT = 1 << 3
MPSORT_REQUIRE_GATHER_SORT = 1 << 4
MPSORT_REQUIRE_SPARSE_ALLTOALLV = 1 << 6
MPSORT_VERIFY_CHECKSUM = 1 << 8
MPSORT_UNIQUE_ID_BYTES = 128


class RadixDesc(ctypes.Structure):
    """struct mpsort_radix_desc"""
    _fields_ = [("offset", c_size_t), ("width", ctypes.c_uint32), ("nwords", ctypes.c_uint32),
                ("is_signed", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class LastStats(ctypes.Structure):
    """struct mpsort_last_stats"""
    _fields_ = [("nmemb", c_u64), ("outnmemb", c_u64), ("elsize", c_u64),
                ("key_words", ctypes.c_uint32), ("first_sort_passes", ctypes.c_uint32),
                ("second_sort_passes", ctypes.c_uint32), ("splitter_rounds", ctypes.c_uint32),
                ("used_gather", ctypes.c_uint32), ("dense_exchange", ctypes.c_uint32),
                ("bytes_sent_remote", c_u64), ("second_sort_merge_tiles", ctypes.c_uint32),
                ("record_mode", ctypes.c_uint32), ("hybrid", ctypes.c_uint32),
                ("hybrid_long_runs", ctypes.c_uint32), ("rebased", ctypes.c_uint32),
                ("p2p_exchange", ctypes.c_uint32), ("exchange_phases", ctypes.c_uint32),
                ("own_slices_in_place", ctypes.c_uint32)]


def _proto(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


# ---- include/mpsort.h ------------------------------------------------------
_proto("mpsort_mpi_set_options", None, c_int)
_proto("mpsort_mpi_has_options", c_int, c_int)
_proto("mpsort_mpi_unset_options", None, c_int)
_proto("mpsort_comm_get_unique_id", c_int, c_void_p)
_proto("mpsort_comm_init_rank", c_void_p, c_int, c_int, c_void_p, c_int)
_proto("mpsort_comm_self", c_void_p, c_int)
_proto("mpsort_comm_init_local_group", c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_void_p))
_proto("mpsort_comm_destroy", None, c_void_p)
_proto("mpsort_comm_rank", c_int, c_void_p)
_proto("mpsort_comm_size", c_int, c_void_p)
_proto("mpsort_comm_device", c_int, c_void_p)
_proto("mpsort_comm_stream", c_void_p, c_void_p)
_proto("mpsort_comm_barrier", None, c_void_p)
_proto("mpsort_comm_allgather_host", None, c_void_p, c_void_p, c_void_p, c_size_t)
_proto("mpsort_comm_allgatherv_host", None, c_void_p, c_void_p, c_size_t, c_void_p, ctypes.POINTER(c_size_t))
_proto("mpsort_mpi_desc_impl", None, c_void_p, c_size_t, c_size_t, ctypes.POINTER(RadixDesc),
       c_void_p, c_int, ctypes.c_char_p)
_proto("mpsort_mpi_newarray_desc_impl", None, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t,
       ctypes.POINTER(RadixDesc), c_void_p, c_int, ctypes.c_char_p)
_proto("radix_sort_desc", None, c_void_p, c_size_t, c_size_t, ctypes.POINTER(RadixDesc), c_int)
# the reference's own signatures: host radix(ptr, radix_out, arg) callback + rsize + arg
RadixFunc = ctypes.CFUNCTYPE(None, c_void_p, c_void_p, c_void_p)
_proto("mpsort_mpi_impl", None, c_void_p, c_size_t, c_size_t, RadixFunc, c_size_t, c_void_p,
       c_void_p, c_int, ctypes.c_char_p)
_proto("mpsort_mpi_newarray_impl", None, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, RadixFunc, c_size_t,
       c_void_p, c_void_p, c_int, ctypes.c_char_p)
_proto("radix_sort", None, c_void_p, c_size_t, c_size_t, RadixFunc, c_size_t, c_void_p)
_proto("mpsort_callback_desc", c_int, c_size_t, ctypes.POINTER(RadixDesc), ctypes.POINTER(c_size_t))
_proto("mpsort_callback_pack", None, c_void_p, c_size_t, c_size_t, RadixFunc, c_size_t, c_void_p, c_void_p)
_proto("mpsort_callback_unpack", None, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p)
_proto("mpsort_release_cached", None)
_proto("mpsort_mpi_report_last_run", None)
_proto("mpsort_mpi_get_last_run", c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double), c_int)
_proto("mpsort_comm_last_stats", None, c_void_p, ctypes.POINTER(LastStats), ctypes.POINTER(c_i64), c_int)
_proto("MPIU_Set_verbose_malloc", None, c_void_p)
# mpiu_set_malloc takes C function pointers; declared loosely
lib.mpiu_set_malloc.restype = None

# ---- host arithmetic exported for CPU unit tests ----------------------------
_proto("mpsort_solve_layout", c_int, c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
       ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64))
_proto("mpsort_cumulative_counts", None, c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64))
_proto("mpsort_key_range", c_int, c_int, ctypes.c_uint32, ctypes.POINTER(c_i64),
       ctypes.POINTER(c_u64), ctypes.POINTER(c_u64), ctypes.POINTER(c_u64),
       ctypes.POINTER(c_u64), ctypes.POINTER(c_u64))

# ---- include/mpsort_util.h --------------------------------------------------
_proto("mpsort_util_device_count", c_int)
_proto("mpsort_util_device_pci_bus_id", c_int, c_int, ctypes.c_char_p, c_int)
_proto("mpsort_util_dev_malloc", c_void_p, c_int, c_size_t)
_proto("mpsort_util_dev_free", None, c_int, c_void_p)
_proto("mpsort_util_host_malloc_pinned", c_void_p, c_size_t)
_proto("mpsort_util_host_free_pinned", None, c_void_p)
_proto("mpsort_util_memcpy", None, c_int, c_void_p, c_void_p, c_size_t)
_proto("mpsort_util_dev_memset", None, c_int, c_void_p, c_int, c_size_t)
_proto("mpsort_util_generate", None, c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_u64)
_proto("mpsort_util_generate_as", None, c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_u64, c_u64, c_u64)
_proto("mpsort_util_multiset_hash", None, c_void_p, c_void_p, c_size_t, c_size_t, ctypes.POINTER(c_u64))
_proto("mpsort_util_check_sorted", c_u64, c_void_p, c_void_p, c_size_t, c_size_t,
       ctypes.POINTER(RadixDesc), c_int, c_size_t, ctypes.POINTER(c_u64))
_proto("mpsort_util_checksum", c_u64, c_void_p, c_void_p, c_size_t)
_proto("mpsort_util_event_create", c_void_p, c_void_p)
_proto("mpsort_util_event_record", None, c_void_p, c_void_p)
_proto("mpsort_util_event_elapsed_ms", ctypes.c_double, c_void_p, c_void_p, c_void_p)
_proto("mpsort_util_event_destroy", None, c_void_p)
_proto("mpsort_util_stream_sync", None, c_void_p)
_proto("mpsort_util_flush_l2", None, c_void_p)
_proto("mpsort_util_launch_count", c_u64, c_int)
_proto("mpsort_util_kernel_timing", None, c_void_p, c_int)
_proto("mpsort_util_kernel_times", c_int, c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double),
       ctypes.POINTER(c_u64), c_int)
_proto("mpsort_util_merge_runs", c_int, c_void_p, c_int, c_void_p, ctypes.POINTER(c_i64), c_void_p, c_size_t,
       ctypes.POINTER(RadixDesc))
_proto("mpsort_util_mem_info", None, c_int, ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t))


def last_run():
    """[(phase, seconds)] of the last sort, as mpsort_mpi_report_last_run prints."""
    names = (ctypes.c_char_p * 64)()
    secs = (ctypes.c_double * 64)()
    n = lib.mpsort_mpi_get_last_run(names, secs, 64)
    return [(names[i].decode(), secs[i]) for i in range(min(n, 64))]


def last_stats(comm_handle, size):
    st = LastStats()
    sc = (c_i64 * max(size, 1))()
    lib.mpsort_comm_last_stats(comm_handle, ctypes.byref(st), sc, size)
    d = {name: getattr(st, name) for name, _ in LastStats._fields_}
    d["sendcounts"] = list(sc)[:size]
    return d

byref = ctypes.byref


def multiset_hash(comm_handle, base, n, elsize):
    """(sum, xor) over the records of mix64-folded record words: mpsort_util_multiset_hash"""
    out = (c_u64 * 2)()
    lib.mpsort_util_multiset_hash(comm_handle, base, n, elsize, out)
    return int(out[0]), int(out[1])


def kernel_times(comm_handle):
    """{class: (milliseconds, launches)} accumulated since mpsort_util_kernel_timing(on)"""
    names = (ctypes.c_char_p * 16)()
    ms = (ctypes.c_double * 16)()
    cnt = (c_u64 * 16)()
    n = lib.mpsort_util_kernel_times(comm_handle, names, ms, cnt, 16)
    return {names[i].decode(): (ms[i], int(cnt[i])) for i in range(n)}
