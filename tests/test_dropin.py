"""Drop-in check of the reference-signature shim (mp-gadget_b200/host/
libgadget_shims.c): the reference's own fixture code, forcetree.c and treewalk.c
call grav_short_tree(act, pm, tree, NULL, rho0, Ti) exactly as run.c:547 /
tests/test_gravity.c:210-213 do, but the symbol is provided by the shim, which
forwards to libb200force.so.  Results land in P[i].FullTreeGravAccel /
P[i].Potential and must match the stock CPU reference (golden fixture)."""
import os
import numpy as np
import pytest

from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_tree_gravity.npz"))
PARKEYS = ("ErrTolForceAcc", "BHOpeningAngle", "MaxBHOpeningAngle", "TreeUseBH", "Rcut", "GravitySoftening", "rho0")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN), reason="oracle/_ref/libref_dropin.so not built")
@pytest.mark.parametrize("name", ["gsl4096", "zeldovich16"])
def test_reference_driver_calls_gpu_grav_short_tree(name):
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_DROPIN)
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    for usebh in (1, 0):
        v = GOLD["%s/bh%d/par" % (name, usebh)]
        par = dict(zip(PARKEYS, [float(x) for x in v]))
        par["TreeUseBH"] = int(par["TreeUseBH"])
        r.tree_build(pos, mass, box, oldacc=GOLD[name + "/oldacc"], topdepth=0)     # reference forcetree.c (CPU)
        acc, pot = r.grav_short_tree(par, 43.0071, int(GOLD[name + "/nmesh"]), 1.5)   # shim -> GPU
        racc, rpot = GOLD["%s/bh%d/acc" % (name, usebh)], GOLD["%s/bh%d/pot" % (name, usebh)]
        scale = np.sqrt((racc ** 2).sum(1)).mean()
        assert np.abs(acc - racc).max() < 1e-6 * scale
        assert np.abs(pot - rpot).max() < 1e-6 * np.abs(rpot).max()
