#!/usr/bin/env python3
"""Generate tests/golden/ref_tree_top.npz: the reference's OWN forcetree.c / gravshort-tree.c (oracle/_ref/libref_tree.so,
single thread) building the tree below an ARBITRARY domain top tree (force_tree_create_topnodes, forcetree.c:654-687,869-934)
and walking it.  Particle sets are those of ref_tree_gravity.npz; the top trees are domain_scenarios.refined_toptree
(randomly refined, 97 and 321 nodes, leaves at depths 1-4).

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden_top.py
"""
import importlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref as R          # noqa: E402
import domain_scenarios as D         # noqa: E402

ics = importlib.import_module("mp-gadget_b200.ics")
G = 43.0071
TOPS = {"top97": dict(seed=4, nrefine=12), "top321": dict(seed=5, nrefine=40)}
CASES = ["uniform3000", "zeldovich16"]


def main():
    r = R.load(nthreads=1)
    assert r is not None, "build oracle/_ref first (make -C oracle ref)"
    gold = np.load(os.path.join(HERE, "ref_tree_gravity.npz"))
    out = {}
    for tname, kw in TOPS.items():
        top = D.refined_toptree(**kw)
        for k, a in zip(("daughter", "startkey", "shift", "leaf"), top):
            out["%s/%s" % (tname, k)] = a
        for name in CASES:
            pos, mass, box = gold[name + "/pos"], gold[name + "/mass"], float(gold[name + "/box"])
            oldacc = gold[name + "/oldacc"]
            r.tree_build_top(pos, mass, box, top)
            t = r.tree_export()
            for k in ("center", "len", "cofm", "mass", "nocc", "part", "toplevel"):
                out["%s/%s/tree/%s" % (tname, name, k)] = t[k]
            for usebh in (1, 0):
                par = ics.tree_params(box, len(mass), treeusebh=usebh, rcut=7.0)
                r.tree_build_top(pos, mass, box, top, oldacc=oldacc)
                acc, pot = r.grav_short_tree(par, G, int(gold[name + "/nmesh"]), 1.5)
                out["%s/%s/bh%d/acc" % (tname, name, usebh)] = acc
                out["%s/%s/bh%d/pot" % (tname, name, usebh)] = pot
    path = os.path.join(HERE, "ref_tree_top.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
