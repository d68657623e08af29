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


def check_nonsplit(rec, gold=GOLD, rtol=1e-12):
    for s, r in enumerate(rec):
        for k in ("info", "scal", "kick", "last", "bin_grav"):
            assert np.array_equal(r[k], gold["nonsplit/%d/%s" % (s, k)]), (s, k)
        assert r["bad"] == int(gold["nonsplit/%d/bad" % s])
        if s in SC.NONSPLIT_KEEP:
            for k in ("pos", "vel", "fullacc"):
                assert close(r[k], gold["nonsplit/%d/%s" % (s, k)], rtol), (s, k)


def test_oracle_nonsplit_loop_equals_reference():
    """Six passes with SplitGravityTimestepsOn = 0: one full tree per pass, the walk for the active particles only."""
    rec = SC.run_nonsplit(make_oracle(), SC.hierarchy_inputs(seed=15, n=1536))
    check_nonsplit(rec)
    assert len({int(r["info"][0]) for r in rec}) >= 3                       # sub-steps with different active sets


def check_gas(rec, gold=GOLD, rtol=1e-12):
    for s, r in enumerate(rec):
        for k in ("info", "scal", "kick", "last", "bin_grav"):
            assert np.array_equal(r[k], gold["gas/%d/%s" % (s, k)]), (s, k)
        assert r["bad"] == int(gold["gas/%d/bad" % s])
        if s in SC.GAS_KEEP:
            for k in ("pos", "vel", "fullacc", "hsml", "entropy"):
                assert close(r[k], gold["gas/%d/%s" % (s, k)], rtol), (s, k)


def test_oracle_gas_hierarchy_equals_reference():
    """Ten passes of the hierarchical loop with gas: hydro-only sub-steps (no gravitationally active particle) and a mixed
    one, where the hierarchical drivers start from a sub-list of the active list."""
    rec = SC.run_gas_hierarchy(make_oracle(), SC.gas_hierarchy_inputs())
    check_gas(rec)
    infos = np.array([r["info"] for r in rec])
    assert (infos[:, 1] == 0).any() and ((infos[:, 1] > 0) & (infos[:, 1] < infos[:, 0])).any()


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


def test_reference_test_timebinmgr_known_answers():
    """libgadget/tests/test_timebinmgr.c:24-90 (sync points 0.1, 0.2, 0.8, 1.0) on the oracle's integer timeline."""
    outs = np.log([0.1, 0.2, 0.8, 1.0])
    O = OS.StepOracle(outs)
    TB = 1 << 46
    assert abs(O.loga_from_ti(0) - outs[0]) < 1e-6 and abs(O.loga_from_ti(TB) - outs[1]) < 1e-6
    assert abs(O.loga_from_ti(TB - 1) - (outs[0] + (outs[1] - outs[0]) * (TB - 1) / TB)) < 1e-6
    assert abs(O.loga_from_ti(TB + 1) - (outs[1] + (outs[2] - outs[1]) / TB)) < 1e-6
    assert abs(O.loga_from_ti(2 * TB) - outs[2]) < 1e-6
    ti = lambda la: int(O.L.oracle_ti_from_loga(OS.C.byref(O.tl), OS.C.c_double(la)))
    assert ti(outs[0]) == 0 and ti(outs[1]) == TB and ti(outs[2]) == 2 * TB
    mid = (outs[2] + outs[1]) / 2
    assert ti(mid) == TB + TB // 2 and abs(O.loga_from_ti(TB + TB // 2) - mid) < 1e-6
    assert ti(0.0) == 3 * TB
    assert abs(O.loga_from_ti(ti(np.log(0.1))) - np.log(0.1)) < 1e-6
    # test_dloga: get_dloga_for_bin = dloga_from_dti(dti_from_timebin(bin))
    Ti = ti(np.log(0.55))
    assert abs(O.dloga_from_dti(0, Ti)) < 1e-6
    assert abs(O.dloga_from_dti(1 << 46, Ti) - (outs[2] - outs[1])) < 1e-6
    assert abs(O.dloga_from_dti(1 << 44, Ti) - (outs[2] - outs[1]) / 4) < 1e-6


def test_reference_test_timefac_known_answers():
    """libgadget/tests/test_timefac.c:78-106: matter-dominated closed forms of the drift / kick factors (Omega0 = 1,
    H0 = 0.1) and the identity hydrokick = drift, on the Gauss-Legendre stand-in both sides of the step fixtures use."""
    amin, amax = 0.005, 1.0
    O = OS.StepOracle(np.log([amin, amax]), Omega0=1.0, Hubble=0.1)
    logdt = (np.log(amax) - np.log(amin)) / (1 << 46)
    ti = lambda a: int((np.log(a) - np.log(amin)) / logdt)
    assert abs(O.factor(0, ti(0.8), ti(0.85)) + 2 / 0.1 * (1 / np.sqrt(0.85) - 1 / np.sqrt(0.8))) < 5e-5
    assert abs(O.factor(1, ti(0.8), ti(0.85)) - 2 / 0.1 * (np.sqrt(0.85) - np.sqrt(0.8))) < 5e-5
    assert abs(O.factor(0, ti(0.8), ti(0.8003)) + 2 / 0.1 * (1 / np.sqrt(0.8003) - 1 / np.sqrt(0.8))) < 5e-6
    assert abs(O.factor(2, ti(0.8), ti(0.85)) - O.factor(0, ti(0.8), ti(0.85))) < 5e-5
    # a more realistic cosmology against an independent quadrature of the same integrands (test_timefac.c:88-102)
    from scipy.integrate import quad
    O2 = OS.StepOracle(np.log([amin, amax]), Omega0=0.25, Hubble=0.1)
    H = lambda a: 0.1 * np.sqrt(0.25 / a ** 3 + 0.75)
    for kind, p, lo, hi in ((0, 3, 0.95, 0.98), (0, 3, 0.05, 0.06), (1, 2, 0.8, 0.85), (1, 2, 0.05, 0.06)):
        want = quad(lambda a: 1 / (H(a) * a ** p), lo, hi, epsrel=1e-10)[0]
        assert abs(O2.factor(kind, ti(lo), ti(hi)) - want) < 5e-5 * max(1.0, abs(want))
