#!/usr/bin/env python3
"""Generate tests/golden/ref_peano.npz: (1) the 64 known-answer Peano-Hilbert keys of the reference's own
libgadget/tests/test_peano.c:107 (read from that file), (2) PEANO() of its compiled utils/peano.c on random
positions incl. the box faces, (3) domain_get_topleaf (domain.h:71-78) over a randomly refined top tree.
Run in the build container:  make -C oracle ref && python tests/golden/make_golden_peano.py"""
import os
import re
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref as R          # noqa: E402
import domain_scenarios as DS        # noqa: E402


def main():
    src = open("/root/reference/libgadget/tests/test_peano.c").read()
    known = np.array([int(v) for v in re.search(r"result_keys\[\] = \{(.*?)\};", src, re.S).group(1).split(",")], np.uint64)
    r = R.Ref(arena_gib=1.0, nthreads=1)
    pos4, box4 = DS.peano_test_positions()
    assert np.array_equal(r.peano_keys(pos4, box4), known)          # the compiled reference reproduces its own test vector
    pos, box = DS.random_positions()
    keys = r.peano_keys(pos, box)
    top = DS.refined_toptree()
    leaf = r.topleaf(keys, *top)
    out = dict(known_keys=known, random_keys=keys, topleaf=leaf, ntop=np.int64(len(top[0])), nleaf=np.int64(leaf.max() + 1))
    path = os.path.join(ROOT, "tests", "golden", "ref_peano.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; top nodes", len(top[0]), "leaves used", len(np.unique(leaf)))


if __name__ == "__main__":
    main()
