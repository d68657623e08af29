"""Primary friends-of-friends linking (oracle only so far): oracle/oracle_fof.c against the reference's
OWN fof.c (tests/golden/ref_fof.npz, generator make_golden_fof.py; live where oracle/_ref/libref_domain.so exists).
There is no CUDA kernel for it yet: this pins the checker the kernel will be held to."""
import os
import sys
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import domain_scenarios as DS        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_fof.npz"))


def test_oracle_fof_primary_equals_reference():
    for k, (pos, ids, typ, box, ll) in enumerate(DS.fof_cases()):
        got = oracle.fof_primary(pos, ids, typ, box, ll)
        want = GOLD["%d/minid" % k]
        assert np.array_equal(got, want), k
        prim = typ == 1
        assert np.array_equal(got[~prim], ids[~prim])                       # other types are not linked
        assert got[-1] == got[-2] == min(ids[-1], ids[-2]) or got[-1] == got[-2]          # the pair across the periodic face
        # the label is the smallest ID of the group and is carried by one of its members
        for label in np.unique(got[prim])[:50]:
            m = prim & (got == label)
            assert ids[m].min() == label
    sizes = np.unique(GOLD["0/minid"], return_counts=True)[1]
    assert sizes.max() > 500 and (sizes == 1).sum() > 500                   # one big clump, many singles


@pytest.mark.skipif(not R.domain_available(), reason="oracle/_ref/libref_domain.so not built")
def test_oracle_fof_equals_reference_live():
    D = R.RefDomain(arena_gib=1.0, nthreads=2)
    if not hasattr(D.L, "ref_fof_primary"):
        pytest.skip("prebuilt libref_domain.so predates ref_fof_primary")
    rng = np.random.default_rng(77)
    n, box = 5000, 64.0
    pos = np.mod(32.0 + rng.standard_normal((n, 3)) * np.array([12.0, 3.0, 1.0]), box)
    ids = rng.permutation(n).astype(np.int64)
    typ = np.ones(n, np.uint8)
    for ll in (0.1, 0.4, 1.5):
        assert np.array_equal(oracle.fof_primary(pos, ids, typ, box, ll), D.fof_primary(pos, ids, typ, box, ll)), ll
