"""PM long-range force pinned against the reference's OWN petapm.c / gravpm.c / powerspectrum.c
(tests/golden/ref_pm.npz, generator make_golden_pm.py: the whole gravpm_force on one rank, PFFT
replaced by plain DFTs in its single-rank layout).  CPU: the oracle restatement; GPU: the CUDA path."""
import os
import numpy as np
import pytest

import oracle
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_pm.npz"))
CASES = ["zeldovich16_n32", "clustered16_n48", "edges_n24"]
G = 43.0071
CM_PER_MPC = 3.085678e24        # libgadget/physconst.h


def _case(name):
    g = lambda k: GOLD[name + "/" + k]
    return g("pos"), g("mass"), float(g("box")), int(g("nmesh")), float(g("asmth"))


def _spectrum(pw, kk, nm, norm, box):
    """powerspectrum_sum (powerspectrum.c:56-92) on the raw bin sums -> (k, P, Nmodes) of the non-empty bins."""
    box_mpc = box * float(GOLD["UnitLength_in_cm"]) / CM_PER_MPC
    sel = nm > 0
    P = pw[sel] / nm[sel] / norm * box_mpc ** 3
    k = kk[sel] / nm[sel] * 2 * np.pi / box_mpc
    return k, P, nm[sel]


@pytest.mark.parametrize("name", CASES)
def test_oracle_pm_equals_reference(name):
    pos, mass, box, nmesh, asmth = _case(name)
    g, p, _ = oracle.pm_force(pos, mass, box, nmesh, asmth, G)
    rg, rp = GOLD[name + "/gravpm"], GOLD[name + "/potential"]
    assert np.abs(g - rg).max() <= 1e-12 * np.abs(rg).max()
    assert np.abs(p - rp).max() <= 1e-12 * np.abs(rp).max()
    k, P, N = _spectrum(*oracle.pm_power(pos, mass, box, nmesh), box)
    assert np.array_equal(N, GOLD[name + "/ps_N"])
    assert np.abs(k - GOLD[name + "/ps_k"]).max() <= 2e-5 * GOLD[name + "/ps_k"].max()          # the file holds 6 digits
    assert np.all(np.abs(P - GOLD[name + "/ps_P"]) <= 2e-5 * np.abs(GOLD[name + "/ps_P"]))


@pytest.mark.skipif(not os.path.exists(R.SO_PM), reason="oracle/_ref/libref_pm.so not built")
def test_oracle_pm_equals_live_reference(tmp_path):
    """A fresh random fixture through the compiled reference PM (when it is available)."""
    rng = np.random.default_rng(31)
    pos = rng.random((3000, 3)) * 10.0
    mass = (0.5 + rng.random(3000)).astype(np.float32)
    r = R.Ref(arena_gib=2.0, nthreads=2, so=R.SO_PM)
    rg, rp = r.gravpm_force(pos, mass, 10.0, 40, 1.5, G, str(tmp_path))
    g, p, _ = oracle.pm_force(pos, mass, 10.0, 40, 1.5, G)
    assert np.abs(g - rg).max() <= 1e-12 * np.abs(rg).max()
    assert np.abs(p - rp).max() <= 1e-12 * np.abs(rp).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_pm_equals_reference(engine, name):
    """CUDA gravpm_force (deposit, cuFFT, Green's function, real-space 4-point difference fused with
    the readout) and its power-spectrum side effect against the reference's own PM."""
    pos, mass, box, nmesh, asmth = _case(name)
    engine.gravpm_init_periodic(box, asmth, nmesh, G)
    engine.set_particles(pos, mass)
    engine.pm_set_power(True)
    g, p = engine.gravpm_force()
    pw, kk, nm, norm = engine.pm_power()
    engine.pm_set_power(False)
    rg, rp = GOLD[name + "/gravpm"], GOLD[name + "/potential"]
    # the difference stencil replaces three inverse FFTs by the identical real-space operator: rounding only
    assert np.abs(g - rg).max() <= 1e-9 * np.abs(rg).max()
    assert np.abs(p - rp).max() <= 1e-10 * np.abs(rp).max()
    k, P, N = _spectrum(pw, kk, nm, norm, box)
    assert np.array_equal(N, GOLD[name + "/ps_N"])
    assert np.all(np.abs(P - GOLD[name + "/ps_P"]) <= 2e-5 * np.abs(GOLD[name + "/ps_P"]))
