"""SPH density (+ smoothing-length iteration) and hydro force: the oracle against
golden vectors from the reference's own compiled density.c / hydra.c, the
reference's test_density.c goldens, and the CUDA path against both."""
import importlib
import os
import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_sph.npz"))
CASES = ["lattice16", "clustered16", "zeldovich16"]
HYDRO = dict(atime=0.5, hubble=0.2, dloga_bin=0.01)
DENS_KEYS = ("hsml", "density", "egywtdensity", "dhsmlfac", "divvel", "curlvel", "dthsml")


def _inputs(name):
    g = lambda k: GOLD[name + "/" + k]
    return g("pos"), g("mass"), g("vel"), g("entropy"), float(g("box")), g("h0")


def _close(a, b, tol):
    return np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("kt", [1, 2])
@pytest.mark.parametrize("DI", [0, 1])
def test_oracle_sph_equals_reference(name, kt, DI):
    pos, mass, vel, ent, box, h0 = _inputs(name)
    n = len(mass)
    t = oracle.OracleTree(pos, mass, box, type=np.zeros(n, np.uint8), mask=1)
    sp = oracle.sph_params(KernelType=kt, MinGasHsml=0.006, DensityIndependentSphOn=DI, **HYDRO)
    d = oracle.density(t, sp, h0, vel=vel, entropy=ent, DoEgyDensity=DI)
    assert d["rc"] == 0
    key = "%s/k%d_di%d/" % (name, kt, DI)
    for k in DENS_KEYS:
        assert _close(d[k], GOLD[key + k], 1e-12), k
    h = oracle.hydro(t, sp, d, vel=vel, entropy=ent)
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k], GOLD[key + "hydro_" + k], 1e-11), k


def test_oracle_reference_test_density_goldens(ics):
    """tests/test_density.c:154-250: 32^3 gas, cubic spline, eta 1, MaxNumNgbDeviation 2,
    set_init_hsml: mean Hsml 0.501747 +- 1e-4 (lattice), 0.187515 +- 1e-3 (clustered, gsl mt19937 seed 0)."""
    box, nc = 8.0, 32
    n = nc ** 3
    sp = oracle.sph_params(KernelType=1, MinGasHsml=0.006, DensityIndependentSphOn=0)
    bg = np.random.MT19937()
    bg._legacy_seeding(4357)
    for pos, want, tol in ((ics.lattice(nc, box), 0.501747, 1e-4), (ics.clustered_mix_from(bg, n, box), 0.187515, 1e-3),
                           (ics.clustered_mix_from(bg, n, box), 0.187515, 1e-3)):
        t = oracle.OracleTree(pos, np.ones(n, np.float32), box, type=np.zeros(n, np.uint8), mask=1)
        h0 = oracle.set_init_hsml(t, 1, 1.0, box)
        d = oracle.density(t, sp, h0, vel=np.full((n, 3), 1.5))
        assert d["rc"] == 0
        assert abs(d["hsml"].mean() - want) < tol
        assert np.all(np.isfinite(d["hsml"])) and np.all(d["density"] > 0) and d["hsml"].min() >= 0.006


def _gpu_sph(b200, engine, pos, mass, vel, ent, box, h0, kt, DI):
    n = len(mass)
    engine.set_particles(pos, mass, type=np.zeros(n, np.uint8))
    engine.force_tree_build(box, mask=1)                       # force_tree_rebuild_mask(GASMASK), run.c:466
    engine.sph_set_gas(h0, vel=vel, entropy=ent)
    sp = b200.sph_params(KernelType=kt, MinGasHsml=0.006, DensityIndependentSphOn=DI, **HYDRO)
    d = engine.density(sp, update_hsml=1, DoEgyDensity=DI)
    h = engine.hydro_force(sp)
    return d, h


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("kt", [1, 2])
@pytest.mark.parametrize("DI", [0, 1])
def test_gpu_sph_equals_reference(b200, engine, name, kt, DI):
    """CUDA density + hydro vs the reference's own compiled density.c / hydra.c."""
    pos, mass, vel, ent, box, h0 = _inputs(name)
    d, h = _gpu_sph(b200, engine, pos, mass, vel, ent, box, h0, kt, DI)
    key = "%s/k%d_di%d/" % (name, kt, DI)
    for k in DENS_KEYS:
        assert _close(d[k], GOLD[key + k], 1e-11), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k], GOLD[key + "hydro_" + k], 1e-10), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_sph_neighbour_counts_equal_oracle(b200, engine, name):
    """Integer neighbour counts (r^2 <= h^2, treewalk.c:1233-1240), pass counts and
    hydro candidate counts are bit-exact against the oracle."""
    pos, mass, vel, ent, box, h0 = _inputs(name)
    n = len(mass)
    d, h = _gpu_sph(b200, engine, pos, mass, vel, ent, box, h0, 2, 1)
    t = oracle.OracleTree(pos, mass, box, type=np.zeros(n, np.uint8), mask=1)
    sp = oracle.sph_params(KernelType=2, MinGasHsml=0.006, DensityIndependentSphOn=1, **HYDRO)
    od = oracle.density(t, sp, h0, vel=vel, entropy=ent, DoEgyDensity=1)
    oh = oracle.hydro(t, sp, od, vel=vel, entropy=ent)
    assert np.array_equal(d["ninteract"], od["ninteract"])
    assert np.array_equal(d["niter"], od["niter"])
    assert np.array_equal(h["ninteract"], oh["ninteract"])
    assert _close(d["numngb"], od["numngb"], 1e-12)


@pytest.mark.gpu
def test_gpu_reference_test_density_goldens(b200, engine, ics):
    """tests/test_density.c goldens through the CUDA path (initial Hsml from the oracle's set_init_hsml)."""
    box, nc = 8.0, 32
    n = nc ** 3
    sp = b200.sph_params(KernelType=1, MinGasHsml=0.006, DensityIndependentSphOn=0)
    bg = np.random.MT19937()
    bg._legacy_seeding(4357)
    for pos, want, tol in ((ics.lattice(nc, box), 0.501747, 1e-4), (ics.clustered_mix_from(bg, n, box), 0.187515, 1e-3)):
        mass = np.ones(n, np.float32)
        t = oracle.OracleTree(pos, mass, box, type=np.zeros(n, np.uint8), mask=1)
        h0 = oracle.set_init_hsml(t, 1, 1.0, box)
        engine.set_particles(pos, mass, type=np.zeros(n, np.uint8))
        engine.force_tree_build(box, mask=1)
        engine.sph_set_gas(h0, vel=np.full((n, 3), 1.5))
        d = engine.density(sp)
        assert abs(d["hsml"].mean() - want) < tol
        assert np.all(d["density"] > 0) and d["hsml"].min() >= 0.006 and d["hsml"].max() <= box


MIXED = np.load(os.path.join(HERE, "golden", "ref_sph_mixed.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["clustered16", "zeldovich16"])
def test_gpu_sph_mixed_timebins_equal_reference(b200, engine, name):
    """A mixed-time-bin step (bins 2,3 active, 4,5 not; tests/golden/make_golden_sph_mixed.py) against
    the reference's own density.c / hydra.c: density + Hsml iteration and hydro force for the active
    particles only, per-bin kick/drift factors, inactive neighbours contributing their stale state."""
    pos, mass, vel, ent, box, h0 = _inputs(name)
    n = len(mass)
    m = lambda k: MIXED[name + "/" + k]
    tb = {k: MIXED["tables/" + k] for k in ("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")}
    bins, act = m("bins"), m("active")
    Ti = int(MIXED["Ti_Current"])
    active_bin = np.array([b <= 0 or Ti % (1 << b) == 0 for b in range(47)])
    tabs = dict(gravkick=tb["gravkick"][:47], hydrokick=tb["hydrokick"][:47], dloga_pred=tb["dloga_pred"][:47],
                drift=np.where(active_bin, 0.0, tb["drift"][:47]),          # hydra.c:178-186: no drift for active bins
                dloga_bin=tb["dloga_bin"][:47])
    engine.set_particles(pos, mass, type=np.zeros(n, np.uint8))
    engine.force_tree_build(box, mask=1)
    engine.sph_set_gas(m("sync_hsml"), vel=m("vel_new"), entropy=ent, dtentropy=m("sync_hydro_dtentropy"),
                       fullacc=m("fullacc"), hydroacc=m("sync_hydro_acc"))
    engine.sph_set_timebins(bins, bins, tabs)
    engine.sph_set_active(act)
    engine.sph_set_state(density=m("sync_density"), egywtdensity=m("sync_egywtdensity"), dhsmlfac=m("sync_dhsmlfac"),
                         divvel=m("sync_divvel"), curlvel=m("sync_curlvel"))
    sp = b200.sph_params(KernelType=2, MinGasHsml=0.006, DensityIndependentSphOn=1, atime=0.5, hubble=0.2,
                         pmkick=float(tb["gravkick"][47]))
    d = engine.density(sp, update_hsml=1, DoEgyDensity=1)
    h = engine.hydro_force(sp)
    for k in DENS_KEYS:
        assert _close(d[k][act], m("mixed_" + k)[act], 1e-11), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k][act], m("mixed_" + k)[act], 1e-10), k
    # inactive particles keep their state
    inact = np.setdiff1d(np.arange(n), act)
    assert np.array_equal(d["density"][inact], m("sync_density")[inact])
    assert np.array_equal(d["hsml"][inact], m("sync_hsml")[inact])


@pytest.mark.parametrize("name", ["clustered16", "zeldovich16"])
def test_oracle_sph_mixed_timebins_equal_reference(name):
    """The oracle's mixed-time-bin mode against the reference's own density.c / hydra.c on the
    mixed fixture (bins 2,3 active, 4,5 not; per-bin factors; stale state of inactive neighbours)."""
    pos, mass, vel, ent, box, h0 = _inputs(name)
    n = len(mass)
    m = lambda k: MIXED[name + "/" + k]
    tb = {k: MIXED["tables/" + k] for k in ("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")}
    bins, act = m("bins"), m("active")
    Ti = int(MIXED["Ti_Current"])
    active_bin = np.array([b <= 0 or Ti % (1 << b) == 0 for b in range(47)])
    tabs = dict(gravkick=tb["gravkick"][:47], hydrokick=tb["hydrokick"][:47], dloga_pred=tb["dloga_pred"][:47],
                drift=np.where(active_bin, 0.0, tb["drift"][:47]), dloga_bin=tb["dloga_bin"][:47])
    t = oracle.OracleTree(pos, mass, box, type=np.zeros(n, np.uint8), mask=1)
    sp = oracle.sph_params(KernelType=2, MinGasHsml=0.006, DensityIndependentSphOn=1, atime=0.5, hubble=0.2,
                           pmkick=float(tb["gravkick"][47]))
    state = {k: m("sync_" + k) for k in ("density", "egywtdensity", "dhsmlfac", "divvel", "curlvel", "dthsml")}
    oracle.sph_set_mixed(bins, bins, tabs, act, n=n)
    try:
        d = oracle.density(t, sp, m("sync_hsml"), vel=m("vel_new"), entropy=ent, dtentropy=m("sync_hydro_dtentropy"),
                           fullacc=m("fullacc"), hydroacc=m("sync_hydro_acc"), DoEgyDensity=1, state=state)
        assert d["rc"] == 0
        h = oracle.hydro(t, sp, d, vel=m("vel_new"), entropy=ent, dtentropy=m("sync_hydro_dtentropy"),
                         fullacc=m("fullacc"), hydroacc=m("sync_hydro_acc"))
    finally:
        oracle.sph_set_mixed()
    for k in DENS_KEYS:
        assert _close(d[k], m("mixed_" + k), 1e-12), k          # inactive particles: untouched state
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k][act], m("mixed_" + k)[act], 1e-11), k


# ---- the quartic spline (DENSITY_KERNEL_QUARTIC_SPLINE = 4, densitykernel.h:17-21) ----------------------------------
GOLD_Q = np.load(os.path.join(HERE, "golden", "ref_sph_quartic.npz"))


@pytest.mark.parametrize("DI", [0, 1])
def test_oracle_sph_quartic_equals_reference(DI):
    pos, mass, vel, ent, box, h0 = _inputs("zeldovich16")
    n = len(mass)
    t = oracle.OracleTree(pos, mass, box, type=np.zeros(n, np.uint8), mask=1)
    sp = oracle.sph_params(KernelType=4, MinGasHsml=0.006, DensityIndependentSphOn=DI, **HYDRO)
    d = oracle.density(t, sp, h0, vel=vel, entropy=ent, DoEgyDensity=DI)
    assert d["rc"] == 0
    key = "zeldovich16/k4_di%d/" % DI
    for k in DENS_KEYS:
        assert _close(d[k], GOLD_Q[key + k], 1e-12), k
    h = oracle.hydro(t, sp, d, vel=vel, entropy=ent)
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k], GOLD_Q[key + "hydro_" + k], 1e-11), k


@pytest.mark.gpu
@pytest.mark.parametrize("DI", [0, 1])
def test_gpu_sph_quartic_equals_reference(b200, engine, DI):
    pos, mass, vel, ent, box, h0 = _inputs("zeldovich16")
    d, h = _gpu_sph(b200, engine, pos, mass, vel, ent, box, h0, 4, DI)
    key = "zeldovich16/k4_di%d/" % DI
    for k in DENS_KEYS:
        assert _close(d[k], GOLD_Q[key + k], 1e-11), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert _close(h[k], GOLD_Q[key + "hydro_" + k], 1e-10), k
