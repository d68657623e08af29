#!/usr/bin/env python3
"""Generate tests/golden/ref_peano.npz: (1) the 64 known-answer Peano-Hilbert keys of the reference's own
libgadget/tests/test_peano.c:107 (read from that file), (2) PEANO() of its compiled utils/peano.c on random
positions incl. the box faces, (3) domain_get_topleaf (domain.h:71-78) over a randomly refined top tree, (4) the per-leaf particle counts of
domain_compute_costs and (5) domain_assign_topleaves_balanced for 120 cost distributions and 1-16 tasks, both
file-static in domain.c and reached by including that file in oracle/ref_domain_driver.c, (6) the top tree of two ranks
through the reference's static stages (local refinement of a subsample, truncation, merge, global refinement, leaves), (7) the exchange list and plan of
every rank of a five-task case (exchange.c statics through oracle/ref_exchange_driver.c).
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
    # domain_compute_costs and domain_assign_topleaves_balanced, file-static in domain.c (oracle/ref_domain_driver.c)
    D = R.RefDomain(arena_gib=1.0, nthreads=2)
    nleaf = int(top[3].max()) + 1
    out["leaf_counts"] = D.leaf_counts(pos, box, top, nleaf, flags=DS.garbage_flags(len(pos)))
    tasks = []
    for ntask, cost in DS.assign_cases():
        t, order = D.assign_balanced(ntask, np.arange(len(cost), dtype=np.uint64) * 8, cost)
        per_leaf = np.zeros(len(cost), np.int32); per_leaf[order] = t
        tasks.append(per_leaf)
    out["assign_tasks"] = np.concatenate(tasks)
    # the top tree of two ranks through the reference's own static stages
    import oracle

    class RefTree:
        def __init__(self, maxnodes):
            self.maxnodes = maxnodes; self.nodes = None; self.size = 0
        tree = property(lambda self: self.nodes[:self.size])
        def local(self, arg):
            pos, bx, sub = arg
            rc, self.nodes, self.size = D.toptree_local(pos, bx, sub, self.maxnodes)
            return rc
        def truncate(self, a, b): self.size = D.toptree_truncate(self.nodes, self.size, a, b)
        def merge(self, other): self.size = D.toptree_merge(self.nodes, self.size, other.nodes); return 0
        def global_refine(self, a, b):
            rc, self.size = D.toptree_global_refine(self.nodes, self.size, a, b); return rc
        def leaves(self): return D.toptree_leaves(self.nodes, self.size)
    # the exchange plan of every rank (exchange.c statics through oracle/ref_exchange_driver.c)
    typ, fl, tl, tk, ntask = DS.exchange_case()
    for r in range(ntask):
        lst, togo, ng = D.exchange_plan(typ, fl, tl, tk, ntask, r)
        out["xplan/%d/list" % r] = lst; out["xplan/%d/togo" % r] = togo; out["xplan/%d/ngarbage" % r] = np.int64(ng)
    for k, case in enumerate(DS.TOPTREE_CASES):
        fields, lf, nl, sizes = DS.toptree_pipeline(RefTree, lambda p_, b_, s_: (p_, b_, s_), case)
        for f, v in fields.items():
            out["toptree/%d/%s" % (k, f)] = v
        out["toptree/%d/leaf" % k] = lf; out["toptree/%d/nleaf" % k] = np.int64(nl); out["toptree/%d/sizes" % k] = sizes
    path = os.path.join(ROOT, "tests", "golden", "ref_peano.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; top nodes", len(top[0]), "leaves used", len(np.unique(leaf)))


if __name__ == "__main__":
    main()
