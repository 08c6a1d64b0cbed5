"""Builds tests/native/_build/libmpsort-hostmock.so: the product's C host files, compiled
unchanged, linked against tests/native/mock_device.c + mock_async.cpp instead of the CUDA kernels,
the CUDA runtime and NCCL (TEST INFRASTRUCTURE; see the headers of those files for what they can and
cannot show). Processes that should use it set MPSORT_LIB to the returned path and MPSORT_ALLOW_MOCK_DEVICE=1 BEFORE importing
mpsort; the product itself never looks for it. MOCK_ASYNC=<seed> in the environment of such a
process turns on the deferred, randomly interleaved execution of stream operations."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "mp-sort_b200", "csrc")
NATIVE = os.path.join(ROOT, "tests", "native")
HOST_FILES = ["mpsort_host.c", "mpsort_comm.c", "mpsort_layout.c", "mpsort_util.c"]
OUT = os.path.join(NATIVE, "_build", "libmpsort-hostmock.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
INCLUDES = ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + os.path.join(ROOT, "oracle"), "-I" + CUDA_INC, "-I" + NATIVE]
DEPS = [os.path.join(CSRC, f) for f in HOST_FILES + ["mpsort_kernels.h", "mpsort_internal.h"]] + \
       [os.path.join(NATIVE, f) for f in ("mock_device.c", "mock_async.cpp", "mock_rename.h")] + \
       [os.path.join(ROOT, "include", "mpsort.h"), os.path.join(ROOT, "include", "mpsort_util.h"), os.path.join(ROOT, "oracle", "synth.h")]


def compile_and_link(out, flags=(), shared=True, extra_sources=()):
    """host files + mock (C part with its stream-ordered entry points renamed, C++ part exporting them) -> out.
    flags: e.g. ("-fsanitize=address,undefined",); extra_sources: C files that bring a main()."""
    objdir = out + ".obj"
    os.makedirs(objdir, exist_ok=True)
    common = ["-O1", "-g", "-Wall", "-fPIC"] + list(flags) + INCLUDES
    objs = []
    for src in [os.path.join(CSRC, f) for f in HOST_FILES] + [os.path.join(NATIVE, "mock_device.c")] + list(extra_sources):
        o = os.path.join(objdir, os.path.basename(src) + ".o")
        subprocess.run(["gcc", "-std=gnu11", "-DMOCK_WITH_ASYNC_LAYER"] + common + ["-c", src, "-o", o], check=True)
        objs.append(o)
    o = os.path.join(objdir, "mock_async.cpp.o")
    subprocess.run(["g++", "-std=c++17"] + common + ["-c", os.path.join(NATIVE, "mock_async.cpp"), "-o", o], check=True)
    objs.append(o)
    # -z defs: every CUDA / NCCL / kernel-ABI symbol the host code uses must be provided by the mock
    link = ["g++"] + list(flags) + (["-shared", "-Wl,-z,defs"] if shared else []) + ["-o", out] + objs + ["-lpthread", "-lm"]
    if any("sanitize" in f for f in flags) and shared:
        link.remove("-Wl,-z,defs")          # the sanitizer runtime is resolved at load time
    subprocess.run(link, check=True)
    return out


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    return compile_and_link(OUT)


if __name__ == "__main__":
    print(build())
