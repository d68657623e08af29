"""The device-resident step loop (csrc/steploop.cu, b200_step_*) against the golden vectors of the
reference's own drift.c / timestep.c / timebinmgr.c (tests/golden/ref_step.npz) and the oracle.

First hardware run: round 2 (all green after working round a ptxas 12.9 miscompile in k_step_active_flags,
see the comment there).  Marker `gpu`."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import step_scenarios as SC          # noqa: E402
import test_step as TS               # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stepper(b200):
    import importlib
    try:
        e = b200.Engine(0)
    except Exception as ex:          # no CUDA device: nothing to run, and no CPU path to fall back to
        pytest.skip("no CUDA device (%s)" % ex)
    SL = importlib.import_module("mp-gadget_b200.steploop")
    O = TS.make_oracle()             # the checker also supplies the cosmology callables (the reference host's cosmology.c / timefac.c)
    cosmo = {k: float(TS.GOLD["cosmo/" + k]) for k in ("Omega0", "OmegaBaryon", "Hubble", "G")}
    ts = {k: float(TS.GOLD["tspar/" + k]) for k in TS.TSKEYS}
    S = SL.StepEngine(e, TS.GOLD["sync_loga"], O.factor, O.hubble, **cosmo, **ts)
    yield S
    e.close()


def test_gpu_primitives_equal_reference(stepper):
    """drift, active list + counts, sub-lists, the three half kicks: lists and counts bit-exact,
    positions / velocities / entropies to 2e-14 of the reference's own C."""
    out = SC.run_primitives(stepper, SC.primitives_inputs())
    TS.check_primitives(out)


def test_gpu_hierarchy_equals_reference(stepper):
    """Eight passes of the hierarchical KDK loop with the tree gravity on the GPU: time bins, kick times,
    PM step bit-exact; positions / velocities to 1e-9 (GPU tree gravity agrees with the reference to ~1e-11)."""
    rec = SC.run_hierarchy(stepper, SC.hierarchy_inputs())
    TS.check_hierarchy(rec, rtol=1e-9)


def test_gpu_gas_hierarchy_equals_reference(stepper):
    """Ten passes of the hierarchical loop with gas (hydro kicks, hydro time bins, Hsml prediction; hydro-only sub-steps)."""
    TS.check_gas(SC.run_gas_hierarchy(stepper, SC.gas_hierarchy_inputs()), rtol=1e-9)


def test_gpu_nonsplit_loop_equals_reference(stepper):
    """Six passes with SplitGravityTimestepsOn = 0: b200_step_grav_short_tree + half kicks + b200_step_find_timesteps."""
    TS.check_nonsplit(SC.run_nonsplit(stepper, SC.hierarchy_inputs(seed=15, n=1536)), rtol=1e-9)


def test_gpu_dropin_step_shims(stepper):
    """The reference's own loop with drift_all_particles, build_active_particles, the half kicks and the hierarchical
    gravity drivers redirected (ld --wrap) to host/libgadget_step_shims.c -> GPU (oracle/_ref/libref_dropin_step.so)."""
    from oracle import ref as R
    if not os.path.exists(R.SO_DROPIN_STEP):
        pytest.skip("oracle/_ref/libref_dropin_step.so not built")
    S = R.RefStep(nthreads=2, arena_gib=1.0, so=R.SO_DROPIN_STEP, **SC.TIMELINE)
    TS.check_primitives(SC.run_primitives(S, SC.primitives_inputs()))
    TS.check_hierarchy(SC.run_hierarchy(S, SC.hierarchy_inputs()), rtol=1e-9)
    TS.check_gas(SC.run_gas_hierarchy(S, SC.gas_hierarchy_inputs()), rtol=1e-9)


@pytest.mark.parametrize("name", ["clustered16", "zeldovich16"])
def test_gpu_gas_substep_device_resident(b200, name):
    """The mixed-time-bin density + hydro step of tests/test_sph.py (golden of the reference's own density.c / hydra.c,
    tests/golden/ref_sph_mixed.npz), but fed from the step state on the device: b200_step_set_state ->
    b200_step_build_active (must find the reference's active list) -> b200_step_sph_prepare -> b200_density ->
    b200_hydro_force -> b200_step_adopt_hydro, no per-particle host array in between."""
    import importlib
    import test_sph as SPH
    try:
        e = b200.Engine(0)
    except Exception as ex:
        pytest.skip("no CUDA device (%s)" % ex)
    SL = importlib.import_module("mp-gadget_b200.steploop")
    M = SPH.MIXED
    pos, mass, vel, ent, box, h0 = SPH._inputs(name)
    n = len(mass)
    m = lambda k: M[name + "/" + k]
    tb = {k: M["tables/" + k] for k in ("gravkick", "hydrokick", "drift", "dloga_pred", "dloga_bin")}
    bins, act = m("bins"), m("active")
    Ti = int(M["Ti_Current"])
    active_bin = np.array([b <= 0 or Ti % (1 << b) == 0 for b in range(47)])
    tabs = dict(gravkick=tb["gravkick"][:47], hydrokick=tb["hydrokick"][:47], dloga_pred=tb["dloga_pred"][:47],
                drift=np.where(active_bin, 0.0, tb["drift"][:47]), dloga_bin=tb["dloga_bin"][:47])
    S = SL.StepEngine(e, np.log([0.1, 1.0]), lambda k, a, b: 0.0, lambda a: 0.2)
    S.set_particles(pos, mass, np.zeros(n, np.uint8), box, vel=m("vel_new"), fullacc=m("fullacc"), bin_grav=bins, bin_hydro=bins,
                    hsml=m("sync_hsml"), hydroacc=m("sync_hydro_acc"), entropy=ent, dtentropy=m("sync_hydro_dtentropy"))
    scal = np.zeros(7, np.int64); scal[3] = Ti; scal[4] = 1 << 40; scal[0] = 1; scal[1] = 46          # not a PM step
    S.set_times(scal, np.zeros(47, np.int64), np.zeros(47, np.int64))
    lst, counts = S.build_active()
    assert np.array_equal(lst, act)
    e.force_tree_build(box, mask=1)
    S.sph_prepare(tabs)
    e.sph_set_state(density=m("sync_density"), egywtdensity=m("sync_egywtdensity"), dhsmlfac=m("sync_dhsmlfac"),
                    divvel=m("sync_divvel"), curlvel=m("sync_curlvel"))
    sp = b200.sph_params(KernelType=2, MinGasHsml=0.006, DensityIndependentSphOn=1, atime=0.5, hubble=0.2, pmkick=float(tb["gravkick"][47]))
    d = e.density(sp, update_hsml=1, DoEgyDensity=1)
    h = e.hydro_force(sp)
    for k in SPH.DENS_KEYS:
        assert SPH._close(d[k][act], m("mixed_" + k)[act], 1e-11), k
    for k in ("acc", "dtentropy", "maxsignalvel"):
        assert SPH._close(h[k][act], m("mixed_" + k)[act], 1e-10), k
    a = S.adopt_hydro()
    inact = np.setdiff1d(np.arange(n), act)
    assert np.array_equal(a["hydroacc"][act], h["acc"][act]) and np.array_equal(a["dtentropy"][act], h["dtentropy"][act])
    assert np.array_equal(a["maxsignalvel"][act], h["maxsignalvel"][act])
    assert np.array_equal(a["hydroacc"][inact], m("sync_hydro_acc")[inact])          # inactive gas keeps its stale state
    assert np.array_equal(S.get()["hsml"][act], d["hsml"][act])                      # the converged Hsml is the state's Hsml
    e.close()
