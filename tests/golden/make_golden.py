#!/usr/bin/env python3
"""Generate tests/golden/ref_tree_gravity.npz by running the reference's OWN C
(oracle/_ref/libref_tree.so = libgadget/forcetree.c, treewalk.c,
gravshort-tree.c, gravity.c compiled unmodified from /root/reference by
oracle/Makefile.ref, single thread) on small seeded inputs.

Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The fixture travels to the GPU box, where /root/reference does not exist.
"""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R          # noqa: E402

ics = importlib.import_module("mp-gadget_b200.ics")
G = 43.0071
ASMTH = 1.5


def cases():
    rng = np.random.default_rng(20261017)
    box = 8.0
    pos = ics.clustered_mix(4096, box, seed=0)
    yield "gsl4096", pos, np.ones(len(pos), np.float32), box, 48
    pos = rng.random((3000, 3)) * box
    yield "uniform3000", pos, (1 + rng.random(3000)).astype(np.float32), box, 48
    pos, mass = ics.zeldovich_lattice(16, 16.0, seed=5)
    yield "zeldovich16", pos, mass, 16.0, 48


def main():
    r = R.load(nthreads=1)
    assert r is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, pos, mass, box, nmesh in cases():
        n = len(mass)
        rng = np.random.default_rng(abs(hash(name)) % 2 ** 31 if False else len(name) * 1000 + n)
        oldacc = rng.standard_normal((n, 3)) * 400.0
        out[name + "/pos"] = pos
        out[name + "/mass"] = mass
        out[name + "/box"] = np.float64(box)
        out[name + "/nmesh"] = np.int32(nmesh)
        out[name + "/oldacc"] = oldacc
        for topdepth in (0, 1):
            r.tree_build(pos, mass, box, topdepth=topdepth)
            t = r.tree_export()
            for k in ("center", "len", "cofm", "mass", "nocc", "part"):
                out["%s/tree%d/%s" % (name, topdepth, k)] = t[k]
        for usebh in (1, 0):
            par = ics.tree_params(box, n, treeusebh=usebh, rcut=7.0)
            r.tree_build(pos, mass, box, oldacc=oldacc, topdepth=0)
            acc, pot = r.grav_short_tree(par, G, nmesh, ASMTH)
            out["%s/bh%d/acc" % (name, usebh)] = acc
            out["%s/bh%d/pot" % (name, usebh)] = pot
            out["%s/bh%d/par" % (name, usebh)] = np.array([par[k] for k in ("ErrTolForceAcc", "BHOpeningAngle", "MaxBHOpeningAngle",
                                                                            "TreeUseBH", "Rcut", "GravitySoftening", "rho0")])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tree_gravity.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
