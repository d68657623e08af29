"""Golden vectors produced by the reference's OWN compiled C (tests/golden/
make_golden.py): the oracle restatement must reproduce them on CPU, and the
CUDA path must reproduce them on the GPU.  Also a live oracle-vs-reference
check when oracle/_ref is present (the build container)."""
import importlib
import os
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_tree_gravity.npz"))
CASES = ["gsl4096", "uniform3000", "zeldovich16"]
G = 43.0071
PARKEYS = ("ErrTolForceAcc", "BHOpeningAngle", "MaxBHOpeningAngle", "TreeUseBH", "Rcut", "GravitySoftening", "rho0")


def _par(name, usebh):
    v = GOLD["%s/bh%d/par" % (name, usebh)]
    p = dict(zip(PARKEYS, [float(x) for x in v]))
    p["TreeUseBH"] = int(p["TreeUseBH"])
    return p


def _check_tree(t, name, topdepth):
    g = lambda k: GOLD["%s/tree%d/%s" % (name, topdepth, k)]
    assert len(t["len"]) == len(g("len"))
    for k in ("center", "len", "nocc", "part", "mass", "cofm"):
        assert np.array_equal(t[k], g(k)), k


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("topdepth", [0, 1])
def test_oracle_tree_equals_reference(name, topdepth):
    """Node-for-node, bit-for-bit: geometry, leaf membership and order, moments
    (libgadget/forcetree.c compiled from the reference, single thread)."""
    ot = oracle.OracleTree(GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"]), toplevel_depth=topdepth)
    on = ot.nodes
    _check_tree({k: on[k] for k in ("center", "len", "nocc", "part", "mass", "cofm")}, name, topdepth)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("usebh", [1, 0])
def test_oracle_walk_equals_reference(name, usebh):
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    ot = oracle.OracleTree(pos, mass, box)
    acc, pot, _ = ot.grav_short_tree(_par(name, usebh), G, int(GOLD[name + "/nmesh"]), 1.5, oldacc=GOLD[name + "/oldacc"])
    racc, rpot = GOLD["%s/bh%d/acc" % (name, usebh)], GOLD["%s/bh%d/pot" % (name, usebh)]
    scale = np.sqrt((racc ** 2).sum(1)).mean()
    # the reference is built with -ffast-math (Options.mk.example:6): equal to rounding, not bit-for-bit
    assert np.abs(acc - racc).max() < 1e-12 * scale
    assert np.abs(pot - rpot).max() < 1e-12 * np.abs(rpot).max()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference_multithreaded():
    """The reference with all its OpenMP threads and a 64-leaf domain (per-thread
    subtrees + merge, forcetree.c:727-860) gives the oracle's accelerations to rounding."""
    ics = importlib.import_module("mp-gadget_b200.ics")
    pos, mass = ics.zeldovich_lattice(32, 32.0, seed=9)
    par = ics.tree_params(32.0, len(mass), treeusebh=1)
    r = R.load()
    acc = r.tree_gravity(pos, mass, 32.0, 96, 1.5, G, par, topdepth=2)
    ot = oracle.OracleTree(pos, mass, 32.0, toplevel_depth=2)
    oacc, _, _ = ot.grav_short_tree(par, G, 96, 1.5)
    assert np.abs(acc - oacc).max() < 1e-11 * np.sqrt((oacc ** 2).sum(1)).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("topdepth", [0, 1])
def test_gpu_tree_equals_reference(engine, name, topdepth):
    engine.set_particles(GOLD[name + "/pos"], GOLD[name + "/mass"])
    engine.force_tree_build(float(GOLD[name + "/box"]), toplevel_depth=topdepth)
    _check_tree(engine.tree_export(), name, topdepth)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("usebh", [1, 0])
def test_gpu_walk_equals_reference(engine, name, usebh):
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    engine.set_particles(pos, mass, oldacc=GOLD[name + "/oldacc"])
    engine.gravpm_init_periodic(box, 1.5, int(GOLD[name + "/nmesh"]), G)
    engine.force_tree_full(box)
    acc, pot, _ = engine.grav_short_tree(_par(name, usebh))
    racc, rpot = GOLD["%s/bh%d/acc" % (name, usebh)], GOLD["%s/bh%d/pot" % (name, usebh)]
    scale = np.sqrt((racc ** 2).sum(1)).mean()
    assert np.abs(acc - racc).max() < 1e-6 * scale          # north_star tolerance
    assert np.abs(acc - racc).max() < 1e-11 * scale         # what we actually achieve
    assert np.abs(pot - rpot).max() < 1e-11 * np.abs(rpot).max()


# ---- the tree below an ARBITRARY domain top tree (force_tree_create_topnodes, forcetree.c:654-687,869-934) -----------
GTOP = np.load(os.path.join(HERE, "golden", "ref_tree_top.npz"))
TOPS = ["top97", "top321"]
TOPCASES = ["uniform3000", "zeldovich16"]


def _top(tname):
    return tuple(GTOP["%s/%s" % (tname, k)] for k in ("daughter", "startkey", "shift", "leaf"))


def _check_tree_top(t, tname, name):
    g = lambda k: GTOP["%s/%s/tree/%s" % (tname, name, k)]
    assert len(t["len"]) == len(g("len"))
    for k in ("center", "len", "nocc", "part", "mass"):
        assert np.array_equal(t[k], g(k)), k
    # the reference is built with -ffast-math (Options.mk.example:6): the centres of mass of a few top-level nodes come
    # out one unit in the last place away from the IEEE quotient
    assert np.abs(t["cofm"] - g("cofm")).max() <= 4e-16 * float(GOLD[name + "/box"])


@pytest.mark.parametrize("name", TOPCASES)
@pytest.mark.parametrize("tname", TOPS)
def test_oracle_tree_below_topnodes_equals_reference(tname, name):
    ot = oracle.OracleTree(GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"]), top_daughter=_top(tname)[0])
    on = ot.nodes
    _check_tree_top({k: on[k] for k in ("center", "len", "nocc", "part", "mass", "cofm")}, tname, name)
    assert np.array_equal(on["toplevel"], GTOP["%s/%s/tree/toplevel" % (tname, name)])
    for usebh in (1, 0):
        acc, pot, _ = ot.grav_short_tree(_par(name, usebh), G, int(GOLD[name + "/nmesh"]), 1.5, oldacc=GOLD[name + "/oldacc"])
        racc = GTOP["%s/%s/bh%d/acc" % (tname, name, usebh)]
        assert np.abs(acc - racc).max() < 1e-12 * np.sqrt((racc ** 2).sum(1)).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("name", TOPCASES)
@pytest.mark.parametrize("tname", TOPS)
def test_gpu_tree_below_topnodes_equals_reference(engine, tname, name):
    """b200_tree_build(toplevel_depth = -1) below the top tree of b200_domain_set_topnodes: node for node the reference's
    tree; counts of the walk equal to the oracle's on the same tree; accelerations equal to the reference's."""
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    top = _top(tname)
    engine.set_particles(pos, mass, oldacc=GOLD[name + "/oldacc"])
    engine.domain_set_topnodes(top)
    engine.force_tree_build(box, toplevel_depth=-1)
    t = engine.tree_export()
    _check_tree_top(t, tname, name)
    ot = oracle.OracleTree(pos, mass, box, top_daughter=top[0])
    assert np.array_equal(t["cofm"], ot.nodes["cofm"])         # IEEE arithmetic on both sides: bit for bit
    engine.gravpm_init_periodic(box, 1.5, int(GOLD[name + "/nmesh"]), G)
    for usebh in (1, 0):
        acc, pot, cnt = engine.grav_short_tree(_par(name, usebh), want_counts=True)
        oacc, opot, ocnt = ot.grav_short_tree(_par(name, usebh), G, int(GOLD[name + "/nmesh"]), 1.5, oldacc=GOLD[name + "/oldacc"])
        for f in ("nodes_accepted", "nodes_opened", "nodes_discarded", "particles"):
            assert np.array_equal(cnt[f], ocnt[f]), f
        racc = GTOP["%s/%s/bh%d/acc" % (tname, name, usebh)]
        assert np.abs(acc - racc).max() < 1e-11 * np.sqrt((racc ** 2).sum(1)).mean()
