"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs on a CPU-only box: the oracle against the golden vectors, host
logic, the C-ABI export check. `-m gpu` are the parity tests proper, through the
C ABI / Python surface of libmpsort-b200.so, compared with the oracle."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mp-sort_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_exception_interact(node, call, report):
    """MPSORT_TEST_INSTAFAIL=1: the failure is written out at once, not in the summary at the end of the session --
    a run under compute-sanitizer that is killed by its time limit after a failure would lose it (tools/sanitize.sh)"""
    if os.environ.get("MPSORT_TEST_INSTAFAIL"):
        sys.stderr.write("\n==== FAILED %s\n%s\n" % (node.nodeid, str(report.longrepr)[-8000:]))
        sys.stderr.flush()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """build the product library and the oracle once per session if missing"""
    lib = os.path.join(ROOT, "mp-sort_b200", "libmpsort-b200.so")
    orc = os.path.join(ROOT, "oracle", "_build", "libmpsort_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()
    yield
    # radix_sort_desc / radix_sort cache one communicator per thread and device: give its arenas back, so
    # that a leak check of the suite (tools/sanitize.sh, compute-sanitizer --leak-check full) comes out clean
    if "mpsort" in sys.modules:
        try:
            sys.modules["mpsort"]._capi.lib.mpsort_release_cached()
        except Exception:
            pass


def load_golden(name):
    import numpy as np
    import mpsort_oracle as O
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["desc"] = O.Desc(*[int(v) for v in g["desc"]])
    g["elsize"] = int(g["elsize"])
    if "records" in g:
        cuts = np.cumsum([0] + [int(s) for s in g["sizes"]])
        g["recs"] = [g["records"][cuts[i]:cuts[i + 1]] for i in range(len(g["sizes"]))]
        ocuts = np.cumsum([0] + [int(s) for s in g["outsizes"]])
        g["exp"] = [g["expected"][ocuts[i]:ocuts[i + 1]] for i in range(len(g["outsizes"]))]
        g["outsizes"] = [int(s) for s in g["outsizes"]]
    return g


GOLDEN_CASES = ["issue7", "mismatched_zeros", "sort_struct", "ties",
                "synth_uniform16", "synth_mostly_sorted16", "synth_particles48"]
