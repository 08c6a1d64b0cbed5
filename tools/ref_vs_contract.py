"""One-off campaign (CPU, needs oracle/_ref = the unmodified reference under the MPI shim): the contract sentence of
SURVEY.md 8(a) -- oracle/mpsort_oracle.py: numpy_sort -- and the C restatement against THE REFERENCE ITSELF on the
random cases of tests/support/hostflow_fuzz.py (ranks, zero-size ranks, output layouts, record and key shapes,
duplicate-heavy / all-equal / sorted / narrow signed keys, option bits). Usage: python tools/ref_vs_contract.py SEED N"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "mp-sort_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
import mpsort_oracle as O  # noqa: E402
import hostflow_fuzz as F  # noqa: E402  (only its case generator; importing it loads the product library, or the mock build with MPSORT_LIB + MPSORT_ALLOW_MOCK_DEVICE=1)


def main():
    seed, n = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(seed)
    done = 0
    while done < n:
        par, recs = F.make_case(rng)
        if max(par["sizes"] + [0]) > 70000:
            continue                                    # the 2^20-record cases: too slow for a campaign of processes
        desc = O.Desc(par["offset"], par["width"], par["nwords"], par["signed"], 0)
        opts = par["opts"] & (O.DISABLE_GATHER_SORT | O.REQUIRE_GATHER_SORT | 2 | 64)
        ref = O.ref_sort(recs, desc, par["outsizes"], opts, inplace=par["inplace"])
        exp = O.numpy_sort(recs, desc, par["outsizes"])
        cs, _ = O.c_sort(recs, desc, par["outsizes"], opts & (O.DISABLE_GATHER_SORT | O.REQUIRE_GATHER_SORT))
        if not all(np.array_equal(a, b) for a, b in zip(ref, exp)) or not all(np.array_equal(a, b) for a, b in zip(ref, cs)):
            print("MISMATCH with the reference at case %d of seed %d:" % (done, seed), par)
            return 1
        done += 1
    print("REFERENCE == CONTRACT == C RESTATEMENT on %d random cases (seed %d)" % (n, seed))
    return 0


if __name__ == "__main__":
    sys.exit(main())
