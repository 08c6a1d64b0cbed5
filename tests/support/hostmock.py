"""Builds tests/native/_build/libmpsort-hostmock.so: the product's C host files, compiled
unchanged, linked against tests/native/mock_device.c instead of the CUDA kernels, the CUDA
runtime and NCCL (TEST INFRASTRUCTURE; see the header of mock_device.c for what it can and
cannot show). Processes that should use it set MPSORT_LIB to the returned path BEFORE
importing mpsort; the product itself never looks for it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "mp-sort_b200", "csrc")
HOST_FILES = ["mpsort_host.c", "mpsort_comm.c", "mpsort_layout.c", "mpsort_util.c"]
OUT = os.path.join(ROOT, "tests", "native", "_build", "libmpsort-hostmock.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def build():
    srcs = [os.path.join(CSRC, f) for f in HOST_FILES] + [os.path.join(ROOT, "tests", "native", "mock_device.c")]
    deps = srcs + [os.path.join(CSRC, "mpsort_kernels.h"), os.path.join(CSRC, "mpsort_internal.h"),
                   os.path.join(ROOT, "include", "mpsort.h"), os.path.join(ROOT, "include", "mpsort_util.h"),
                   os.path.join(ROOT, "oracle", "synth.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -z defs: every CUDA / NCCL / kernel-ABI symbol the host code uses must be provided by the mock
    subprocess.run(["gcc", "-O1", "-g", "-Wall", "-fPIC", "-std=gnu11", "-shared", "-I" + os.path.join(ROOT, "include"),
                    "-I" + CSRC, "-I" + os.path.join(ROOT, "oracle"), "-I" + CUDA_INC, "-o", OUT] + srcs
                   + ["-lpthread", "-Wl,-z,defs"], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
