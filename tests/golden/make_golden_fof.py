#!/usr/bin/env python3
"""Generate tests/golden/ref_fof.npz with the reference's OWN fof.c (fof_label_primary, compiled unmodified and reached by
including the file in oracle/ref_fof_driver.c): the MinID label of every particle after the primary friends-of-friends
linking, for a clustered box with a second particle type mixed in, a pair straddling the periodic boundary, and a
linking length of 0.2 mean spacings.  Run in the build container:
    make -C oracle ref && python tests/golden/make_golden_fof.py"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref as R          # noqa: E402
import domain_scenarios as DS        # noqa: E402


def main():
    D = R.RefDomain(arena_gib=1.0, nthreads=2)
    out = {}
    for k, (pos, ids, typ, box, ll) in enumerate(DS.fof_cases()):
        out["%d/minid" % k] = D.fof_primary(pos, ids, typ, box, ll)
    path = os.path.join(ROOT, "tests", "golden", "ref_fof.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
