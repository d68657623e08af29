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


GOLD_SPH = np.load(os.path.join(HERE, "golden", "ref_sph.npz"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN_SPH), reason="oracle/_ref/libref_dropin_sph.so not built")
@pytest.mark.parametrize("name", ["clustered16", "zeldovich16"])
@pytest.mark.parametrize("kt,DI", [(1, 0), (2, 1)])
def test_reference_driver_calls_gpu_density_and_hydro(name, kt, DI):
    """The reference's fixture code (ref_driver.c, modelled on tests/test_density.c:55-152)
    calls density() and hydro_force() with the reference signatures; the symbols come from
    libgadget_sph_shims.c and run on the GPU.  Results land in P[]/SphP[] and must match the
    stock density.c / hydra.c (golden fixture)."""
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_DROPIN_SPH)
    g = lambda k: GOLD_SPH[name + "/" + k]
    d = r.sph_density(g("pos"), g("mass"), float(g("box")), g("h0"), vel=g("vel"), entropy=g("entropy"), kerneltype=kt,
                      init_hsml=False, DoEgyDensity=DI)
    h = r.sph_hydro(atime=0.5, hubble=0.2, dloga_bin=0.01, DensityIndependentSphOn=DI)
    key = "%s/k%d_di%d/" % (name, kt, DI)
    close = lambda a, b, tol: np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)
    for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel", "dthsml"):
        if k == "egywtdensity" and not DI:
            continue            # not written without DoEgyDensity (density.c:566-568)
        assert close(d[k], GOLD_SPH[key + k], 1e-11), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert close(h[k], GOLD_SPH[key + "hydro_" + k], 1e-10), k


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN_SPH) or not os.path.exists(R.SO), reason="oracle/_ref not built")
def test_shim_set_init_hsml_equals_reference():
    """set_init_hsml of the shim (walks the caller's host ForceTree) against density.c:700-749,
    through the converged smoothing lengths of tests/test_density.c's lattice case."""
    n1 = 16
    box = 8.0
    x = (np.arange(n1) + 0.5) * box / n1
    pos = np.stack(np.meshgrid(x, x, x, indexing="ij"), -1).reshape(-1, 3)
    mass = np.ones(len(pos), np.float32)
    out = []
    for so in (R.SO, R.SO_DROPIN_SPH):
        r = R.Ref(arena_gib=2.0, nthreads=2, so=so)
        out.append(r.sph_density(pos, mass, box, np.ones(len(pos)), kerneltype=1, init_hsml=True, meansep=1.0))
    assert np.abs(out[0]["hsml"] - out[1]["hsml"]).max() <= 1e-11 * out[0]["hsml"].max()
    assert np.abs(out[0]["density"] - out[1]["density"]).max() <= 1e-11 * out[0]["density"].max()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN_ALL), reason="oracle/_ref/libref_dropin_all.so not built")
def test_gpu_only_tree_mode():
    """forcetree.c replaced too (libgadget_forcetree_shims.c): force_tree_full /
    force_tree_rebuild_mask / force_tree_calc_moments only describe the tree, the octree exists
    in HBM only; the reference fixture code still gets the stock results."""
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_DROPIN_ALL)
    name = "gsl4096"
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    v = GOLD["%s/bh0/par" % name]
    par = dict(zip(PARKEYS, [float(x) for x in v])); par["TreeUseBH"] = int(par["TreeUseBH"])
    r.tree_build(pos, mass, box, oldacc=GOLD[name + "/oldacc"], topdepth=0)
    acc, pot = r.grav_short_tree(par, 43.0071, int(GOLD[name + "/nmesh"]), 1.5)
    racc, rpot = GOLD["%s/bh0/acc" % name], GOLD["%s/bh0/pot" % name]
    assert np.abs(acc - racc).max() < 1e-6 * np.sqrt((racc ** 2).sum(1)).mean()
    assert np.abs(pot - rpot).max() < 1e-6 * np.abs(rpot).max()
    g = lambda k: GOLD_SPH["zeldovich16/" + k]
    d = r.sph_density(g("pos"), g("mass"), float(g("box")), g("h0"), vel=g("vel"), entropy=g("entropy"), kerneltype=2,
                      init_hsml=False, DoEgyDensity=1)
    h = r.sph_hydro(atime=0.5, hubble=0.2, dloga_bin=0.01, DensityIndependentSphOn=1)
    close = lambda a, b, tol: np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)
    assert close(d["hsml"], GOLD_SPH["zeldovich16/k2_di1/hsml"], 1e-11)
    assert close(d["density"], GOLD_SPH["zeldovich16/k2_di1/density"], 1e-11)
    assert close(h["acc"], GOLD_SPH["zeldovich16/k2_di1/hydro_acc"], 1e-10)


GOLD_MIXED = np.load(os.path.join(HERE, "golden", "ref_sph_mixed.npz"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN_SPH), reason="oracle/_ref/libref_dropin_sph.so not built")
def test_reference_driver_mixed_timebin_step_through_shims():
    """The mixed-time-bin fixture (ActiveParticles list, per-bin DriftKickTimes) driven through the
    reference-signature density() / hydro_force() shims: the shim derives the per-bin factors with
    the reference's own init_kick_factor_data / dloga_from_dti / get_exact_drift_factor calls."""
    name = "zeldovich16"
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_DROPIN_SPH)
    g = lambda k: GOLD_SPH[name + "/" + k]
    m = lambda k: GOLD_MIXED[name + "/" + k]
    r.sph_density(g("pos"), g("mass"), float(g("box")), g("h0"), vel=g("vel"), entropy=g("entropy"), kerneltype=2,
                  init_hsml=False, DoEgyDensity=1)
    h0 = r.sph_hydro(atime=0.5, hubble=0.2, dloga_bin=0.01, DensityIndependentSphOn=1)
    tb = {k: GOLD_MIXED["tables/" + k] for k in ("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")}
    act, out = r.sph_mixed(m("bins"), int(GOLD_MIXED["Ti_Current"]), tb, 0.5, 0.2, 1, vel=m("vel_new"), fullacc=m("fullacc"),
                           hydroacc=h0["acc"], dtentropy=h0["dtentropy"])
    assert np.array_equal(act, m("active"))
    close = lambda a, b, tol: np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)
    for k in ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel", "dthsml"):
        assert close(out[k], m("mixed_" + k), 1e-11), k          # active: new values; inactive: untouched state
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert close(out[k], m("mixed_" + k), 1e-10), k


GOLD_PM = np.load(os.path.join(HERE, "golden", "ref_pm.npz"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN_PM), reason="oracle/_ref/libref_dropin_pm.so not built")
@pytest.mark.parametrize("name", ["zeldovich16_n32", "clustered16_n48"])
def test_reference_driver_calls_gpu_gravpm_force(name, tmp_path):
    """gravpm_init_periodic + gravpm_force with the reference signatures (run.c:330,522), provided by the
    shim and run on the GPU: P[i].GravPM / P[i].Potential equal the reference's own PM (golden), and the
    power-spectrum file the reference's powerspectrum_save writes from the GPU sums equals its own."""
    g = lambda k: GOLD_PM[name + "/" + k]
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_DROPIN_PM)
    gp, pot = r.gravpm_force(g("pos"), g("mass"), float(g("box")), int(g("nmesh")), float(g("asmth")), 43.0071, str(tmp_path))
    assert np.abs(gp - g("gravpm")).max() <= 1e-9 * np.abs(g("gravpm")).max()
    assert np.abs(pot - g("potential")).max() <= 1e-10 * np.abs(g("potential")).max()
    ps = np.loadtxt(os.path.join(str(tmp_path), "powerspectrum-1.0000.txt"))
    assert np.array_equal(ps[:, 2].astype(np.int64), g("ps_N"))
    assert np.all(np.abs(ps[:, 1] - g("ps_P")) <= 2e-5 * np.abs(g("ps_P")))
    assert np.all(np.abs(ps[:, 0] - g("ps_k")) <= 2e-5 * np.abs(g("ps_k")))


GTOP = np.load(os.path.join(HERE, "golden", "ref_tree_top.npz"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(R.SO_DROPIN), reason="oracle/_ref/libref_dropin.so not built")
@pytest.mark.parametrize("so", ["SO_DROPIN", "SO_DROPIN_ALL"])
def test_shims_build_below_the_domain_top_tree(so):
    """The device tree is built below the caller's domain top tree (forcetree.c:654-687), taken from the host tree's
    top-level nodes (grav_short_tree gets no DomainDecomp) or, in GPU-only tree mode, from ddecomp->TopNodes: with the
    same node set the walk takes the reference's decisions, so the accelerations agree to rounding, not just to 1e-6."""
    if not os.path.exists(getattr(R, so)):
        pytest.skip("not built")
    r = R.Ref(arena_gib=2.0, nthreads=2, so=getattr(R, so))
    name = "uniform3000"
    pos, mass, box = GOLD[name + "/pos"], GOLD[name + "/mass"], float(GOLD[name + "/box"])
    for tname in ("top97", "top321"):
        top = tuple(GTOP["%s/%s" % (tname, k)] for k in ("daughter", "startkey", "shift", "leaf"))
        for usebh in (1, 0):
            v = GOLD["%s/bh%d/par" % (name, usebh)]
            par = dict(zip(PARKEYS, [float(x) for x in v])); par["TreeUseBH"] = int(par["TreeUseBH"])
            r.tree_build_top(pos, mass, box, top, oldacc=GOLD[name + "/oldacc"])
            acc, pot = r.grav_short_tree(par, 43.0071, int(GOLD[name + "/nmesh"]), 1.5)
            racc = GTOP["%s/%s/bh%d/acc" % (tname, name, usebh)]
            assert np.abs(acc - racc).max() < 1e-11 * np.sqrt((racc ** 2).sum(1)).mean()
    # the 64-leaf domain of the reference's multi-threaded build
    r.tree_build(pos, mass, box, oldacc=GOLD[name + "/oldacc"], topdepth=2)
    acc, pot = r.grav_short_tree(par, 43.0071, int(GOLD[name + "/nmesh"]), 1.5)
    racc = GOLD["%s/bh0/acc" % name]
    assert np.abs(acc - racc).max() < 1e-6 * np.sqrt((racc ** 2).sum(1)).mean()
