"""Step loop around the force computation: integer timeline, drift, active lists,
half kicks and the hierarchical gravity driver, pinned against the reference's OWN drift.c /
timestep.c / timebinmgr.c (tests/golden/ref_step.npz, generator make_golden_step.py).
CPU: the oracle restatement (oracle/oracle_step.c) against the golden and, when oracle/_ref/libref_step.so
is present, against the reference run live."""
import os
import sys
import numpy as np
import pytest

from oracle import ref as R
from oracle import step as OS

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import step_scenarios as SC          # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_step.npz"))
TSKEYS = ("ErrTolIntAccuracy", "MaxGasVel", "MaxSizeTimestep", "MinSizeTimestep", "MaxRMSDisplacementFac", "CourantFac")
RTOL = 2e-14          # the two sides differ only by the compiler's re-association (-ffast-math in the reference build)


def make_oracle():
    cosmo = {k: float(GOLD["cosmo/" + k]) for k in ("Omega0", "OmegaBaryon", "Hubble", "G")}
    ts = {k: float(GOLD["tspar/" + k]) for k in TSKEYS}
    return OS.StepOracle(GOLD["sync_loga"], **cosmo, **ts)


def close(a, b, rtol=RTOL):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


def check_primitives(out, gold=GOLD, rtol=RTOL):
    for k in ("active", "active_counts", "last_drift", "sublist36", "sublist37", "sublist41", "pmkick_times", "kick_times",
              "pm_active_counts", "pm_sublist38", "hydro_bad", "hydro_bins", "hydro_times", "find_bad", "find_bin_grav", "find_bin_hydro",
              "find_times", "findpm_bad", "findpm_bin_grav", "findpm_bin_hydro", "findpm_times"):
        assert np.array_equal(out[k], gold["prim/" + k]), k
    for k in ("ddrift", "drift_pos", "drift_hsml", "halfkick_vel", "halfkick_entropy", "hydrokick_vel", "hydrokick_entropy", "pmkick_vel"):
        assert close(out[k], gold["prim/" + k], rtol), k


def check_hierarchy(rec, gold=GOLD, rtol=1e-12):
    for s, r in enumerate(rec):
        for k in ("info", "scal", "kick", "last", "bin_grav"):
            assert np.array_equal(r[k], gold["hier/%d/%s" % (s, k)]), (s, k)
        assert r["bad"] == int(gold["hier/%d/bad" % s])
        if s in SC.HIER_KEEP:
            for k in ("pos", "vel", "fullacc"):
                assert close(r[k], gold["hier/%d/%s" % (s, k)], rtol), (s, k)


def test_oracle_timeline_equals_reference():
    O = make_oracle()
    ti, dloga, span = SC.timeline_samples()
    assert np.array_equal(np.array([O.loga_from_ti(int(t)) for t in ti]), GOLD["tl/loga"])
    assert np.array_equal(np.array([O.dti_from_dloga(float(d), int(t)) for t, d in zip(ti, dloga)], np.int64), GOLD["tl/dti"])
    assert np.array_equal(np.array([O.dloga_from_dti(12345, int(t)) for t in ti]), GOLD["tl/dloga"])
    f = np.array([[O.factor(k, int(t), int(t + s)) for k in range(3)] for t, s in zip(ti, span)])
    assert np.abs(f - GOLD["tl/factor"]).max() <= 1e-14 * np.abs(GOLD["tl/factor"]).max()
    assert (f[span == 0] == 0).all()


def test_oracle_primitives_equal_reference():
    d = SC.primitives_inputs()
    out = SC.run_primitives(make_oracle(), d)
    check_primitives(out)
    # the scenario exercises what it claims to
    assert (GOLD["prim/drift_pos"] > 0).all() and (GOLD["prim/drift_pos"] <= d["box"]).all()
    gas = d["type"] == 0
    assert np.isclose(GOLD["prim/drift_hsml"][gas].max(), d["box"] / 2)                       # the Hsml cap
    cap = float(GOLD["tspar/MaxGasVel"]) * np.exp(make_oracle().loga_from_ti(int(SC.primitives_times()[0][3])))
    assert np.isclose(np.linalg.norm(GOLD["prim/halfkick_vel"][gas], axis=1), cap, rtol=1e-12).any()          # the velocity cap acted
    hb = GOLD["prim/hydro_bins"]; act = GOLD["prim/active"]
    changed = hb[act] != d["bin_hydro"][act]
    assert changed.sum() > 50 and len(np.unique(hb[act][gas[act]])) >= 4                       # hydro bins really re-assigned
    assert 0 < len(GOLD["prim/sublist36"]) < len(GOLD["prim/sublist37"]) < len(GOLD["prim/active"]) < len(d["mass"])


def test_oracle_hierarchy_equals_reference():
    """Eight passes of the hierarchical KDK loop: time bins, kick times and the PM step length bit-exact,
    positions / velocities / accelerations to 1e-12."""
    rec = SC.run_hierarchy(make_oracle(), SC.hierarchy_inputs())
    check_hierarchy(rec)
    bins = rec[-1]["bin_grav"]
    assert len(np.unique(bins)) >= 4                                                          # a real hierarchy


@pytest.mark.skipif(not R.step_available(), reason="oracle/_ref/libref_step.so not built")
def test_oracle_equals_reference_live():
    """A second, differently seeded pair of scenarios against the reference run in-process."""
    S = R.RefStep(nthreads=2, arena_gib=1.0, **SC.TIMELINE)
    O = OS.StepOracle(S.sync_loga, **S.cosmo, **S.tspar)
    d = SC.primitives_inputs(seed=21, n=1200, box=2500.0)
    a, b = SC.run_primitives(S, d), SC.run_primitives(O, d)
    gold = {"prim/" + k: v for k, v in a.items()}
    check_primitives(b, gold)
    h = SC.hierarchy_inputs(seed=33, n=1024, box=9000.0)
    ra, rb = SC.run_hierarchy(S, h, steps=5), SC.run_hierarchy(O, h, steps=5)
    for s, (x, y) in enumerate(zip(ra, rb)):
        for k in ("info", "scal", "kick", "last", "bin_grav"):
            assert np.array_equal(x[k], y[k]), (s, k)
        for k in ("pos", "vel", "fullacc"):
            assert close(y[k], x[k], 1e-12), (s, k)


def test_bench_cosmology_helper_matches_oracle(ics):
    """ics.FlatLCDM (the bench / tool stand-in for the reference host's cosmology.c + timefac.c) against the oracle's."""
    c = ics.FlatLCDM()
    O = OS.StepOracle(c.sync, Omega0=c.Omega0, Hubble=c.Hubble)
    for kind, t0, t1 in [(0, 0, 1 << 40), (1, (1 << 46) + 5, (1 << 46) + (1 << 42)), (2, 3 << 44, (3 << 44) + (1 << 30)), (1, 7, 7)]:
        a, b = c.factor(kind, t0, t1), O.factor(kind, t0, t1)
        assert abs(a - b) <= 1e-12 * max(abs(b), 1e-300)
    assert abs(c.hubble(0.37) - O.hubble(0.37)) < 1e-15 and c.loga_from_ti((1 << 46) + 12345) == O.loga_from_ti((1 << 46) + 12345)
